#!/usr/bin/env python
"""Generates tests/golden/*.npz from the oracle (run on CPU: `python tests/golden/make_golden.py`).

The reference itself (Julia + Gridap) cannot run in the build image, so these fixtures are outputs of the CPU restatement
`oracle/`, which is pinned against the reference's own tests (tests/test_oracle_*.py) -- including the two published regression
pairs, which are stored here as the reference states them.  The GPU parity tests compare the CUDA path against these files
(tests/test_gpu_golden.py), so a later change of the oracle cannot silently move the target.
tools/dump_reference.jl produces the same arrays from the real reference wherever Julia + Gridap are available.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import advection, forms, shell, swe, tswe  # noqa: E402
from oracle import surface_stokes, vector_h1  # noqa: E402
from oracle import rk as ork  # noqa: E402
from oracle.mesh import CubedSphereMesh2D  # noqa: E402
from oracle.spaces import CGSpace, DGSpace, RTSpace  # noqa: E402


def main():
    # 1. Laplace-Beltrami, config 1 flavour: n = 4, Q1 / Q2, degrees 12 / 18
    out = {}
    mesh = CubedSphereMesh2D(4, 1.0)
    out["cell_nodes"] = mesh.cell_nodes
    out["cell_chart_coords"] = mesh.cell_chart_coords
    for p, deg in ((1, 12), (2, 18)):
        V = CGSpace(mesh, p)
        q = forms.CellQuadrature(mesh, deg)
        Ke = forms.lb_element_matrices(V, q)
        A = forms.assemble_matrix(V, V, Ke)
        out[f"q{p}_cell_dofs"] = V.cell_dofs
        out[f"q{p}_Ke_first8"] = Ke[:8]
        out[f"q{p}_indptr"], out[f"q{p}_indices"], out[f"q{p}_data"] = A.indptr, A.indices, A.data
    np.savez_compressed(os.path.join(HERE, "laplace_beltrami_n4.npz"), **out)

    # 2. wave (RT1 x DG1) and shallow water (RT1 x DG1 + CG2): numbering, one residual, three steps, regression pairs
    out = {}
    mesh = CubedSphereMesh2D(3, 1.0)
    W = ork.WaveProblem(mesh, 1)
    rng = np.random.default_rng(20261019)
    x = rng.standard_normal(W.nu + W.np_)
    out["rt1_cell_dofs"], out["rt1_signs"] = RTSpace(mesh, 1).cell_dofs, RTSpace(mesh, 1).signs
    out["wave_x"], out["wave_res"] = x, W.residual(x)
    xs = x.copy()
    for _ in range(3):
        xs = W.step(xs, 0.01)
    out["wave_x_after_3_steps_dt0.01_ssp33"] = xs
    P = swe.ShallowWaterProblem(mesh, 1)
    dt = swe.reference_dt(mesh, 1)
    x0 = P.initial_state()
    y0 = P.diagnostics(x0)
    P.dt, P.tau = dt, dt / 2
    out["swe_x0"], out["swe_y0"], out["swe_res0"], out["swe_dt"] = x0, y0, P.residual_x(x0, y0, y0), dt
    xs = x0.copy()
    for _ in range(2):
        xs, ys = P.step(xs, dt)
    out["swe_x_after_2_steps"], out["swe_y_after_2_steps"] = xs, ys
    out["reference_wave_eu_ep"] = np.array([0.001602838116424487, 0.010030006030668944])      # TransientWaveEquation.jl:128-129
    out["reference_swe_eu_ep"] = np.array([0.0035085422284689785, 3.5206285360509003e-6])     # TransientShallowWater.jl:208-209
    np.savez_compressed(os.path.join(HERE, "transient_n3.npz"), **out)

    # 3. 3-D shell (hex Q2) and DG advection
    out = {}
    m3 = shell.CubedSphereShellMesh3D(2, 1.0, 0.19)
    V3 = shell.HexCGSpace(m3, 2)
    Ke = shell.lb_element_matrices(V3, shell.ShellQuadrature(m3, 8))
    A = forms.assemble_matrix(V3, V3, Ke)
    out["shell_cell_nodes"], out["shell_cell_dofs"], out["shell_Ke_first4"] = m3.cell_nodes, V3.cell_dofs, Ke[:4]
    out["shell_indptr"], out["shell_indices"], out["shell_data"] = A.indptr, A.indices, A.data
    mesh = CubedSphereMesh2D(4, 1.0)
    beta = lambda X: np.stack([-X[..., 1], X[..., 0], 0 * X[..., 0]], -1)   # noqa: E731
    Pa = advection.AdvectionProblem(mesh, 2, beta)
    u = rng.standard_normal(Pa.Q.num_free)
    out["adv_u"], out["adv_res"] = u, Pa.residual(u)
    np.savez_compressed(os.path.join(HERE, "shell_advection.npz"), **out)
    # 4. next-row callers: thermal shallow water, vector H1 space, closed-form surface operators, assembled advection operator
    out = {}
    mesh = CubedSphereMesh2D(3, 1.0)
    T = tswe.ThermalShallowWaterProblem(mesh, 1)
    x0 = T.initial_state() * (1 + 1e-2 * np.cos(np.arange(T.nu + 2 * T.np_)))
    dt = 0.01
    T.dt, T.tau = dt, dt / 2
    y0 = T.diagnostics(x0)
    out["tswe_x0"], out["tswe_y0"], out["tswe_res0"], out["tswe_dt"] = x0, y0, T.residual_x(x0, y0, y0), dt
    out["tswe_x_after_1_step"] = T.step(x0, dt)[0]
    wX = lambda X: np.stack([X[..., 2] * X[..., 1], -X[..., 0] ** 2, X[..., 0] + 0.3], -1)   # noqa: E731
    Vv = vector_h1.VectorCGSpace(mesh, 2)
    q = forms.CellQuadrature(mesh, 12)
    out["vh1_cell_dofs"], out["vh1_cob"] = Vv.cell_dofs, Vv.cob
    Av = forms.assemble_matrix(Vv, Vv, vector_h1.mass_element_matrices(Vv, q))
    out["vh1_mass_indptr"], out["vh1_mass_indices"], out["vh1_mass_data"] = Av.indptr, Av.indices, Av.data
    out["vh1_interp"] = vector_h1.interpolate(Vv, wX)
    fX = lambda X: X[..., 0] ** 2 + X[..., 1] ** 2 * X[..., 2]                                 # noqa: E731
    gfX = lambda X: np.stack([2 * X[..., 0], 2 * X[..., 1] * X[..., 2], X[..., 1] ** 2], -1)    # noqa: E731

    def hfX(X):
        H = np.zeros(X.shape + (3,))
        H[..., 0, 0] = 2.0
        H[..., 1, 1] = 2 * X[..., 2]
        H[..., 1, 2] = H[..., 2, 1] = 2 * X[..., 1]
        return H
    q6 = forms.CellQuadrature(mesh, 6)
    out["surf_lap"] = forms.surface_laplacian(fX, gfX, hfX, mesh, q6.x, mesh.cell_to_chart)
    out["surf_grad"] = forms.surface_gradient(gfX, mesh, q6.x, mesh.cell_to_chart)
    out["surf_div_of_grad"] = forms.surface_divergence(gfX, hfX, mesh, q6.x, mesh.cell_to_chart)
    Pa = advection.AdvectionProblem(mesh, 1, beta, degree=4)
    Aa = Pa.A.tocsr()
    Aa.sort_indices()
    out["adv_matrix_dense_n3_p1"] = Aa.toarray()
    np.savez_compressed(os.path.join(HERE, "next_rows_n3.npz"), **out)
    # 5. vector Laplacian / Stokes blocks on the vector H1 space (n = 3, Q2 x Q1): element matrices of the first cells (cell 0
    #    touches a panel corner), assembled matrices, the H1 seminorm of an interpolant
    out = {}
    mesh = CubedSphereMesh2D(3, 1.0)
    Vv, Qs = vector_h1.VectorCGSpace(mesh, 2), CGSpace(mesh, 1, zeromean=True)
    q = forms.CellQuadrature(mesh, 18)
    K = surface_stokes.vector_laplacian_element_matrices(Vv, q)
    B = surface_stokes.div_element_matrices(Qs, Vv, q)
    out["K_first4"], out["B_first4"] = K[:4], B[:4]
    A = forms.assemble_matrix(Vv, Vv, K)
    out["A_indptr"], out["A_indices"], out["A_data"] = A.indptr, A.indices, A.data
    Bm = forms.assemble_matrix(Qs, Vv, B)
    out["B_indptr"], out["B_indices"], out["B_data"] = Bm.indptr, Bm.indices, Bm.data
    wX = lambda X: np.stack([X[..., 2] * X[..., 1], -X[..., 0] ** 2, X[..., 0] + 0.3], -1)   # noqa: E731
    U = vector_h1.interpolate(Vv, wX)
    out["h1_of_interp"] = surface_stokes.h1_error(Vv, q, U, 0 * U) ** 2 - surface_stokes.fe_l2_error(Vv, q, U, 0 * U) ** 2
    np.savez_compressed(os.path.join(HERE, "surface_stokes_n3.npz"), **out)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
