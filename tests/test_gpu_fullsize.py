"""Parity at BASELINE.json's full sizes, where the numpy oracle is too slow: size-independent properties of the forms.

* Laplace-Beltrami (target config: 256x256 per panel, CG-Q2, degree 18): constants are in the kernel (A 1 = 0), A is symmetric,
  the fused residual equals A u_h, the stiffness "energy" of the interpolant of xyz equals int |grad_s f|^2 = 12 int f^2
  (f = xyz is a degree-3 spherical harmonic), a sample of element matrices equals the C restatement of the oracle.
* scalar mass: the sum of all entries is the area 4 pi r^2.
* shallow water (config 4: 256x256, RT2 x DG2 + CG3): int h sqrtg is conserved to round-off over the steps.
* DG advection (config 3: 128x128, DG-Q2): int c sqrtg is conserved to round-off.
* 3-D shell (config 5 size used by the bench: 48^3 per panel, hex-Q2, degree 18): A 1 = 0, residual = A u_h, volume.
"""
import math

import numpy as np
import pytest

from __graft_entry__ import load_package

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gg():
    return load_package()


def fX(X):
    return X[..., 0] * X[..., 1] * X[..., 2]


def test_laplace_beltrami_target_config_properties(gg):
    n, p, degree = 256, 2, 18
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1")
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    assert V.num_free_dofs == 1572866
    u = gg.interpolate(fX, V)
    assem = gg.SparseMatrixAssembler(V, V)
    A, b = gg.assemble_matrix_and_vector(gg.LaplaceBeltrami(dΩ, u=u), assem)
    assert A.nnz == 25165826
    S = A.to_scipy()
    scale = np.abs(S.data).max()
    # constants are in the kernel; symmetric
    assert np.abs(S @ np.ones(S.shape[0])).max() < 1e-12 * scale
    D = (S - S.T).tocsr()
    assert (np.abs(D.data).max() if D.nnz else 0.0) < 1e-13 * scale
    # the fused residual is A u_h (same element matrices, different summation order)
    uh = u.get()
    r = b.get()
    tol = 1e-13 * scale * np.abs(uh).max()          # r is small by cancellation: round-off scales with |A| |u|
    assert np.abs(r - S @ uh).max() < tol
    # device SpMV agrees with the host product
    y = gg.DeviceVector(model.ctx, V.num_free_dofs)
    A.spmv(u, y)
    assert np.abs(y.get() - S @ uh).max() < tol
    # energy of the interpolant of a degree-3 harmonic: int |grad_s f|^2 = l(l+1) int f^2, int (xyz)^2 = 4 pi / 105
    energy = float(uh @ r)
    assert abs(energy - 12 * 4 * math.pi / 105) < 1e-7
    # a sample of element matrices against the C restatement of the oracle (first, middle, last cells)
    from oracle import c_oracle
    cc = model.get_cell_coordinates()
    for first in (0, 3 * n * n + 17, 6 * n * n - 8):
        Ke = gg.element_matrices(gg.LaplaceBeltrami(dΩ), V, V, first=first, n=8)
        Ko = c_oracle.lb_element_matrices(cc[first:first + 8], 1.0, p, degree, 1)
        assert np.abs(Ke - Ko).max() <= 1e-12 * np.abs(Ko).max()


def test_scalar_mass_sums_to_the_area(gg):
    n, r = 256, 1.7
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(r), n=n)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, 2), conformity="H1")
    dΩ = gg.Measure(gg.Triangulation(model), 12)
    M = gg.assemble_matrix(gg.ScalarMass(dΩ), V, V)
    assert abs(M.values().sum() - 4 * math.pi * r * r) < 1e-10 * 4 * math.pi * r * r
    assert abs(dΩ.integrate() - 4 * math.pi * r * r) < 1e-10 * 4 * math.pi * r * r


def test_shallow_water_config4_conserves_mass(gg):
    ic = gg.initial_conditions
    n, p = 256, 2
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.raviart_thomas, float, p), conformity="HDiv")
    Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="L2")
    assert V.num_free_dofs == 7077888 and Q.num_free_dofs == 3538944
    dΩ = gg.Measure(gg.Triangulation(model), 4 * (p + 1))
    one = gg.assemble_vector(gg.Source(dΩ, np.ones(model.num_cells * dΩ.num_points)), Q).get()     # int phi_i sqrtg
    x0 = np.concatenate([gg.interpolate(ic.galewsky_velocity, V).get(), gg.interpolate(ic.galewsky_depth, Q).get()])
    dt = gg.dx(model) * 0.1 / (p * math.sqrt(ic.GRAVITY * 1.0e4 / ic.A_E))
    st = gg.RKStepper(gg.RungeKutta(gg.CGSolver(rtol=1e-12), None, dt, "EXRK_SSP_3_3"),
                      gg.ShallowWaterOperator(model, p, ic.coriolis, gravity=ic.GRAVITY))
    st.set_state(x0)
    st.step(4)
    x = st.get_state()
    assert np.isfinite(x).all()
    m0, m1 = float(one @ x0[st.nu:]), float(one @ x[st.nu:])
    assert abs(m1 - m0) < 1e-12 * abs(m0)
    assert 0 < np.abs(x - x0).max() < 1e-2 * np.abs(x0).max()


def test_advection_config3_conserves_the_tracer(gg):
    n, p = 128, 2
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="L2")
    assert Q.num_free_dofs == 884736
    dΩ = gg.Measure(gg.Triangulation(model), 4 * p)
    one = gg.assemble_vector(gg.Source(dΩ, np.ones(model.num_cells * dΩ.num_points)), Q).get()
    wind = lambda X: np.stack([-X[..., 1], X[..., 0], 0 * X[..., 0]], -1)   # noqa: E731
    st = gg.RKStepper(gg.RungeKutta(None, None, 2 * math.pi / 20000, "EXRK_SSP_3_3"), gg.AdvectionOperator(model, p, wind))
    c0 = gg.interpolate(lambda X: np.exp(-(X[..., 1] ** 2 + X[..., 2] ** 2)), Q).get()
    st.set_state(c0)
    st.step(50)
    c = st.get_state()
    assert abs(one @ c - one @ c0) < 1e-12 * abs(one @ c0)
    assert 0 < np.abs(c - c0).max() < 0.05


def test_shell_config5_properties(gg):
    n, t = 48, 0.19
    model = gg.AtlasDiscreteModel(gg.CubedSphereWithThicknessMesh(1.0, t), n=n)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, 2), conformity="H1")
    dΩ = gg.Measure(gg.Triangulation(model), 18)
    rng = np.random.default_rng(20261019)
    u = gg.DeviceVector.from_numpy(model.ctx, rng.standard_normal(V.num_free_dofs))
    assem = gg.SparseMatrixAssembler(V, V)
    A, b = gg.assemble_matrix_and_vector(gg.LaplaceBeltrami(dΩ, u=u), assem)
    vals_max = np.abs(A.values()).max()
    one = gg.DeviceVector(model.ctx, V.num_free_dofs)
    one.fill(1.0)
    y = gg.DeviceVector(model.ctx, V.num_free_dofs)
    A.spmv(one, y)
    assert np.abs(y.get()).max() < 1e-11 * vals_max
    A.spmv(u, y)
    r = b.get()
    assert np.abs(y.get() - r).max() < 1e-11 * np.abs(r).max()
    vol = 4 * math.pi / 3 * ((1 + t) ** 3 - 1)
    assert abs(dΩ.integrate() - vol) < 1e-9 * vol
