"""GPU parity at the sizes BASELINE.json names, as VALUE comparisons against the oracle (not only through size-independent
properties, tests/test_gpu_fullsize.py):

  cfg 1   Laplace-Beltrami, 32x32 per panel, CG-Q1, degree 12 (test/Laplacian/LaplaceBeltrami.jl:29-53 at nref 5)
  cfg 2   linear wave, 48x48 per panel, RT1 x DG1, degree 10, SSPRK(2,2) and SSPRK(3,3) (test/Transient/TransientWaveEquation.jl)
  target  Laplace-Beltrami, 256x256 per panel, CG-Q2, degree 18: the whole CSR matrix (25 165 826 values) and the residual
          against the C restatement of the reference's per-cell loop
  boundary  CSC export (SparseMatrixCSC) of rectangular / non-symmetric operators

Tolerances: pattern and numbering bit-exact; values 1e-12 relative to the largest entry (north_star)."""
import numpy as np
import pytest

from __graft_entry__ import load_package
from oracle import c_oracle, forms
from oracle import rk as ork
from oracle.mesh import CubedSphereMesh2D
from oracle.spaces import CGSpace, DGSpace, RTSpace

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.fixture(scope="module")
def gg():
    return load_package()


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def test_cfg1_laplace_beltrami_n32_q1(gg):
    n, p, degree = 32, 1, 12
    mesh = CubedSphereMesh2D(n, 1.0)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1", constraint="zeromean")
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    Vo = CGSpace(mesh, p, zeromean=True)
    assert (V.get_cell_dof_ids() == np.where(Vo.cell_dofs >= 0, Vo.cell_dofs + 1, -1)).all()        # numbering, bit-exact
    q = forms.CellQuadrature(mesh, degree)
    Ao = forms.assemble_matrix(Vo, Vo, forms.lb_element_matrices(Vo, q))
    u = np.random.default_rng(20261019).standard_normal(V.num_free_dofs)
    assem = gg.SparseMatrixAssembler(V, V)
    A, b = gg.assemble_matrix_and_vector(gg.LaplaceBeltrami(dΩ, u=u), assem)
    A = A.to_scipy()
    assert A.shape == Ao.shape == (6145, 6145)
    assert (A.indptr == Ao.indptr).all() and (A.indices == Ao.indices).all()
    assert rel(A.data, Ao.data) < RTOL
    # residual with the Dirichlet (zeromean) DOF at 0: r = A u on the free rows
    assert rel(b.get(), Ao @ u) < RTOL
    # the manufactured right-hand side of the driver (u = xyz, closed-form surface Laplacian) on the device
    xyz = dΩ.quadrature_points(chart=False)[1]
    f = -12.0 * xyz[..., 0] * xyz[..., 1] * xyz[..., 2]             # Delta_s(xyz) on the unit sphere: -l(l+1) xyz, l = 3
    rhs = gg.assemble_vector(gg.Source(dΩ, f), V).get()
    assert rel(rhs, forms.assemble_vector(Vo, forms.source_element_vectors(Vo, q, f))) < RTOL


@pytest.mark.parametrize("tableau", ["EXRK_SSP_2_2", "EXRK_SSP_3_3"])
def test_cfg2_wave_n48(gg, tableau):
    n, p = 48, 1
    mesh = CubedSphereMesh2D(n)
    W = ork.WaveProblem(mesh, p)                                   # degree 10 = 5 (p + 1), TransientWaveEquation.jl:36
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    dt = ork.dx(mesh) * 0.1 / p
    st = gg.WaveStepper(gg.RungeKutta(gg.CGSolver(rtol=1e-14), None, dt, tableau), gg.WaveOperator(model, p))
    assert (st.nu, st.np_) == (W.nu, W.np_) == (110592, 55296)

    def depth(X):  # TransientWaveEquation.jl:23-25
        return 1.0 + 0.01 * np.exp(-5 * ((1 - X[..., 0]) ** 2 + X[..., 1] ** 2 + X[..., 2] ** 2))

    x0 = np.concatenate([np.zeros(W.nu), ork.interpolate_scalar(W.Q, mesh, depth)])
    Qg = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="L2")
    assert rel(gg.interpolate(depth, Qg).get(), x0[W.nu:]) < 1e-14
    xr = x0 + 1e-3 * np.random.default_rng(3).standard_normal(x0.size)
    r, ro = st.residual(xr), W.residual(xr)
    assert rel(r, ro) < RTOL
    st.set_state(x0)
    st.step(3)
    x = x0.copy()
    for _ in range(3):
        x = W.step(x, dt, tableau)
    assert rel(st.get_state(), x) < 1e-11


def test_target_full_matrix(gg):
    """256x256 per panel, CG-Q2, degree 18: every CSR value and residual entry against the C oracle (all host cores)."""
    n, p, degree = 256, 2, 18
    mesh = CubedSphereMesh2D(n, 1.0)
    Vo = CGSpace(mesh, p)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1")
    assert V.num_free_dofs == Vo.num_free == 1572866
    assert (V.get_cell_dof_ids() - 1 == Vo.cell_dofs).all()                       # DOF numbering, bit-exact
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    assem = gg.SparseMatrixAssembler(V, V)
    u = np.random.default_rng(20261019).standard_normal(V.num_free_dofs)
    A, b = gg.assemble_matrix_and_vector(gg.LaplaceBeltrami(dΩ, u=u), assem)
    rp, ci = A.pattern()
    assert A.nnz == 25165826
    rpo, cio = forms.csr_pattern(Vo, Vo)
    assert (rp == rpo).all() and (ci == cio).all()                                 # CSR pattern, bit-exact
    Ke = c_oracle.lb_element_matrices(mesh.cell_chart_coords, 1.0, p, degree, 0)
    vo = c_oracle.scatter_csr(Vo.cell_dofs, Ke, rpo, cio, 0)
    del Ke
    vals = A.values()
    assert rel(vals, vo) < RTOL
    fe = c_oracle.lb_element_vectors(mesh.cell_chart_coords, 1.0, p, degree, forms.gather(Vo, u), 0)
    ro = c_oracle.scatter_vector(Vo.cell_dofs, fe, Vo.num_free, 0)
    assert rel(b.get(), ro) < RTOL
    # the geometry-class path gives the same bits
    model.ctx.set_geometry_classes(True)
    try:
        assem_c = gg.SparseMatrixAssembler(V, V)
        A2, b2 = gg.assemble_matrix_and_vector(gg.LaplaceBeltrami(dΩ, u=u), assem_c)
        assert (A2.values() == vals).all()
    finally:
        model.ctx.set_geometry_classes(False)


@pytest.mark.parametrize("k", [0, 1, 2])
def test_csc_export_of_rectangular_blocks(gg, k):
    """SparseMatrixCSC view (gg_csr_export_csc) of the DG x RT divergence block -- rectangular, not symmetric -- against
    scipy's CSC of the oracle matrix: colptr / rowval bit-exact (0- and 1-based), nzval 1e-12."""
    n, degree = 5, 5 * (k + 1)
    mesh = CubedSphereMesh2D(n, 1.0)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    q = forms.CellQuadrature(mesh, degree)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.raviart_thomas, float, k), conformity="HDiv")
    Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, k), conformity="L2")
    Vo, Qo = RTSpace(mesh, k), DGSpace(mesh, k)
    Bo = forms.assemble_matrix(Qo, Vo, forms.div_element_matrices(Qo, Vo, q)).tocsc()
    Bo.sort_indices()
    B = gg.assemble_matrix(gg.Div(dΩ), V, Q)                      # rows: test Q, columns: trial V
    assert B.shape == Bo.shape and B.shape[0] != B.shape[1]
    Bc = B.to_scipy_csc()
    assert (Bc.indptr == Bo.indptr).all() and (Bc.indices == Bo.indices).all()
    assert rel(Bc.data, Bo.data) < RTOL
    cp, rv, nz = B.to_scipy_csc(index_base=1)
    assert (cp == Bo.indptr + 1).all() and (rv == Bo.indices + 1).all() and (nz == Bc.data).all()
    assert abs(Bc - B.to_scipy()).max() == 0.0                    # CSC and CSR views hold the same matrix
    # values-only refresh after a new numeric phase
    B2 = gg.assemble_matrix(gg.Div(dΩ, scale=2.0), V, Q)
    B2.to_scipy_csc()
    assert (B2.csc_values() == 2.0 * nz).all()


def test_csc_export_of_the_advection_operator(gg):
    """Facet-coupled, non-symmetric DG operator (jac = a + b + s of AdvectionUpwinding.jl:124-131): CSC = transpose-free view."""
    n, p = 4, 1
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    wind = lambda X: np.stack([-X[..., 1], X[..., 0], 0.3 * X[..., 2]], -1)   # noqa: E731
    Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="L2")
    A = gg.assemble_advection_matrix(gg.AdvectionOperator(model, p, wind, degree=4 * p), Q)
    Ar, Ac = A.to_scipy(), A.to_scipy_csc()
    assert abs(Ar - Ar.T).max() > 1e-3                              # really non-symmetric
    assert abs(Ac - Ar).max() == 0.0
    ref = Ar.tocsc()
    ref.sort_indices()
    assert (Ac.indptr == ref.indptr).all() and (Ac.indices == ref.indices).all() and (Ac.data == ref.data).all()
