"""GPU parity: linearised Boussinesq equations on the 3-D shell (test/Tutorials/LinearBoussinesq.jl:112-134), hexahedral
RT_k x DG_k x DG_k, against the oracle (oracle/boussinesq.py) through the C ABI: numbering bit-exact, mass matrix, interpolated
initial state, stage residual 1e-12, explicit RK steps 1e-10, and the exact energy balance of the semi-discrete system."""
import numpy as np
import pytest

from __graft_entry__ import load_package
from oracle import boussinesq as obq
from oracle.shell import CubedSphereShellMesh3D

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gg():
    return load_package()


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.mark.parametrize("k,n,L", [(0, 2, 2), (1, 2, 2), (1, 3, 1)])
def test_boussinesq_matches_oracle(gg, k, n, L):
    r, t = 1.0, 0.19
    mesh = CubedSphereShellMesh3D(n, r, t, n_layers=L)
    P = obq.BoussinesqProblem(mesh, k)
    model = gg.AtlasDiscreteModel(gg.CubedSphereWithThicknessMesh(r, t), n=n, n_layers=L)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.raviart_thomas, float, k), conformity="HDiv",
                       dirichlet_tags=["bottom_boundary", "top_boundary"])
    Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, k), conformity="L2")
    # numbering and signs, bit-exact
    assert V.num_free_dofs == P.V.num_free and Q.num_free_dofs == P.Q.num_free
    ids = V.get_cell_dof_ids()
    assert (np.where(ids < 0, -1, ids - 1) == P.V.cell_dofs).all()
    assert (V.get_cell_dof_signs() == P.V.signs).all()
    # interpolated initial state (u_cf = meas * pinvJ . u0, moment functionals of the hexahedral element)
    u0 = gg.interpolate(obq.u0, V).get()
    assert rel(u0, obq.interpolate_rt(P.V, obq.u0)) < 1e-12
    b0 = gg.interpolate(obq.b0(r), Q).get()
    assert np.abs(b0 - obq.interpolate_scalar(P.Q, obq.b0(r))).max() < 1e-14        # |b0| <= 1 (zero at the nodes when L = 1)
    dt = 0.01
    op = gg.BoussinesqOperator(model, k, obq.omega, c=1.0, N=1.48)
    for tableau in ("EXRK_SSP_3_3", "EXRK_SSP_2_2"):
        st = gg.RKStepper(gg.RungeKutta(gg.CGSolver(rtol=1e-14), None, dt, tableau), op)
        assert (st.nu, st.np_) == (P.nu, 2 * P.np_)
        x = np.random.default_rng(4).standard_normal(P.nu + 2 * P.np_)
        rr, ro = st.residual(x), P.residual(x)
        for a, b in zip(P.split(rr), P.split(ro)):
            assert rel(a, b) < 1e-12
        # exact energy balance of the semi-discrete system on the device result
        u, p, b = P.split(x)
        ru, rp, rb = P.split(rr)
        assert abs(u @ ru + p @ rp / P.c2 + b @ rb / P.N2) < 1e-11 * (abs(u @ ru) + abs(p @ rp) + abs(b @ rb))
        x0 = np.concatenate([u0, np.zeros(P.np_), b0])
        st.set_state(x0)
        st.step(3)
        xo = x0.copy()
        for _ in range(3):
            xo = P.step(xo, dt, tableau)
        assert rel(st.get_state(), xo) < 1e-10
        iters, solves = st.stats()
        assert solves == 3 * len(gg.TABLEAUS[tableau][1]) and iters > 0


def test_boussinesq_requests_outside_the_implemented_range_fail_loudly(gg):
    model = gg.AtlasDiscreteModel(gg.CubedSphereWithThicknessMesh(1.0, 0.19), n=2)
    with pytest.raises(gg.GGError):
        gg.TestFESpace(model, gg.ReferenceFE(gg.raviart_thomas, float, 2), conformity="HDiv")
    # the stepper needs the 3-D shell
    surf = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=2)
    op = gg.BoussinesqOperator.__new__(gg.BoussinesqOperator)
    op.model, op.p_fe, op.degree, op.c, op.N = surf, 0, 4, 1.0, 1.0
    op.omega_q = np.zeros((surf.num_cells, 27))
    with pytest.raises(gg.GGError):
        gg.RKStepper(gg.RungeKutta(gg.CGSolver(rtol=1e-12), None, 0.01, "EXRK_SSP_3_3"), op)
