"""GPU parity: DG upwind advection (volume + skeleton terms, test/Tutorials/AdvectionUpwinding.jl:107-137) with explicit RK
against the oracle (oracle/advection.py), plus the structural properties the scheme must have (exact conservation of
int u sqrtg; constants are steady for the solenoidal solid-body wind up to quadrature error)."""
import numpy as np
import pytest

from __graft_entry__ import load_package
from oracle import advection
from oracle import rk as ork
from oracle.mesh import CubedSphereMesh2D

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gg():
    return load_package()


def beta(X):          # AdvectionUpwinding.jl:100-102
    return np.stack([-X[..., 1], X[..., 0], 0 * X[..., 0]], -1)


def u0(X):            # AdvectionUpwinding.jl:96-98
    return np.exp(-(X[..., 1] ** 2 + X[..., 2] ** 2))


@pytest.mark.parametrize("n,p", [(4, 1), (5, 2), (6, 0)])
def test_residual_and_steps_match_oracle(gg, n, p):
    mesh = CubedSphereMesh2D(n)
    degree = max(4 * p, 2)
    P = advection.AdvectionProblem(mesh, p, beta, degree=degree)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    dt = 2 * np.pi / (40 * n * (2 * p + 1))
    st = gg.RKStepper(gg.RungeKutta(None, None, dt, "EXRK_SSP_3_3"), gg.AdvectionOperator(model, p, beta, degree=degree))
    assert (st.nu, st.np_) == (0, P.Q.num_free)
    x = np.random.default_rng(3).standard_normal(P.Q.num_free)
    r, ro = st.residual(x), P.residual(x)
    assert np.abs(r - ro).max() < 1e-12 * np.abs(ro).max()
    Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="L2")
    x0 = gg.interpolate(u0, Q).get()
    assert np.abs(x0 - ork.interpolate_scalar(P.Q, mesh, u0)).max() < 1e-14
    st.set_state(x0)
    st.step(5)
    xo = x0.copy()
    for _ in range(5):
        xo = P.step(xo, dt)
    xg = st.get_state()
    assert np.abs(xg - xo).max() < 1e-11 * np.abs(xo).max()
    # conservation: the discrete mass does not change (to round-off); constants: res(1) is quadrature error only
    assert abs(P.total_mass(xg) - P.total_mass(x0)) < 1e-13 * P.total_mass(x0)
    one = st.residual(np.ones(P.Q.num_free))
    assert np.abs(one).max() < 1e-5 * np.abs(ro).max()


def test_solid_body_rotation_returns_the_initial_bump(gg):
    # one full revolution (t = 2 pi) of u0 = exp(-(y^2 + z^2)) about the z axis: the exact solution is u0 itself
    n, p = 12, 2
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="L2")
    nsteps = 600
    st = gg.RKStepper(gg.RungeKutta(None, None, 2 * np.pi / nsteps, ":EXRK_SSP_3_3"), gg.AdvectionOperator(model, p, beta))
    x0 = gg.interpolate(u0, Q).get()
    st.set_state(x0)
    st.step(nsteps)
    e = gg.norm(st.get_state() - x0, Q, 8) / gg.norm(x0, Q, 8)
    assert e < 2e-3, e


@pytest.mark.parametrize("n,p", [(1, 1), (3, 2), (4, 0)])
def test_assembled_operator_with_skeleton_pattern_matches_oracle(gg, n, p):
    """jac = a + b + s (test/Tutorials/AdvectionUpwinding.jl:124-131) on the facet-coupled DG pattern; the matrix-free residual of
    the explicit stepper is the product with it."""
    from oracle import advection as oadv
    from oracle.mesh import CubedSphereMesh2D
    degree = 4 * max(p, 1)
    wind = lambda X: np.stack([-X[..., 1], X[..., 0] + 0.2 * X[..., 2], 0.1 * X[..., 0]], -1)   # noqa: E731
    mesh = CubedSphereMesh2D(n)
    P = oadv.AdvectionProblem(mesh, p, wind, degree=degree)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="L2")
    op = gg.AdvectionOperator(model, p, wind, degree=degree)
    A = gg.assemble_advection_matrix(op, Q)
    M = (p + 1) ** 2
    assert A.nnz == model.num_cells * M * 5 * M
    S = A.to_scipy()
    S.sort_indices()
    Ao = P.A.tocsr()
    # the oracle's pattern is the set of structurally coupled pairs too; compare as dense-free sparse difference
    D = (S - Ao).tocsr()
    assert (np.abs(D.data).max() if D.nnz else 0.0) <= 1e-12 * np.abs(Ao.data).max()
    rp, ci = A.pattern()
    assert (np.diff(rp) == 5 * M).all() and all((np.diff(ci[rp[r]:rp[r + 1]]) > 0).all() for r in range(0, len(rp) - 1, 7))
    u = np.random.default_rng(4).standard_normal(Q.num_free_dofs)
    st = gg.RKStepper(gg.RungeKutta(None, None, 1e-3, "EXRK_SSP_3_3"), op)
    assert np.abs(st.residual(u) - S @ u).max() <= 1e-12 * np.abs(S @ u).max()
