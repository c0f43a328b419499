"""GPU parity: thermal shallow water (test/Tutorials/ThermalShallowWater.jl) -- the shallow-water DAE stepper plus the buoyancy
cell and skeleton terms -- against oracle/tswe.py: diagnostics, stage residual, whole steps, conservation."""
import numpy as np
import pytest

from __graft_entry__ import load_package
from oracle import rk as ork
from oracle import swe as oswe
from oracle import tswe as otswe
from oracle.mesh import CubedSphereMesh2D

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gg():
    return load_package()


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def _setup(gg, n, p, q0_mode="alias"):
    mesh = CubedSphereMesh2D(n)
    P = otswe.ThermalShallowWaterProblem(mesh, p, q0_mode=q0_mode)
    dt = ork.dx(mesh) * 0.05 / np.sqrt(oswe._g * oswe._H_0)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n)
    op = gg.ThermalShallowWaterOperator(model, p, oswe.coriolis(), q0_mode=q0_mode)
    st = gg.RKStepper(gg.RungeKutta(gg.CGSolver(rtol=1e-14), None, dt, "EXRK_SSP_3_3"), op)
    assert st.nu == P.nu and st.np_ == 2 * P.np_ and st.ny == P.nq + P.nu + 2 * P.np_
    x0 = P.initial_state()
    rng = np.random.default_rng(11)
    x0 = x0 * (1 + 1e-2 * rng.standard_normal(x0.shape))       # off balance: buoyancy jumps on every facet
    return mesh, P, st, x0, dt


@pytest.mark.parametrize("n,p", [(4, 1), (3, 2), (4, 0)])
def test_diagnostics_and_stage_residual_match_oracle(gg, n, p):
    mesh, P, st, x0, dt = _setup(gg, n, p)
    P.dt, P.tau = dt, dt / 2
    y = P.diagnostics(x0)
    yg = st.diagnostics(x0)
    for a, b, name in zip(P.split_y(yg), P.split_y(y), "qFPb"):
        assert _rel(a, b) < 1e-10, name
    r = P.residual_x(x0, y, y)
    rg = st.stage_residual(x0, y)
    a, b = P.nu, P.nu + P.np_
    for sl, name in ((slice(0, a), "r_u"), (slice(a, b), "r_p"), (slice(b, None), "r_B")):
        assert np.abs(rg[sl] - r[sl]).max() < 1e-11 * np.abs(r[sl]).max(), name
    # q0 != q exercises the -tau (q - q0)/dt term together with the buoyancy terms
    y0 = y * (1 + 1e-3 * np.cos(np.arange(y.size)))
    r = P.residual_x(x0, y, y0)
    rg = st.stage_residual(x0, y, q0=y0[: P.nq])
    assert np.abs(rg - r).max() < 1e-11 * np.abs(r).max()


@pytest.mark.parametrize("n,p,mode", [(4, 1, "alias"), (3, 2, "alias"), (4, 1, "start_of_step")])
def test_steps_match_oracle_and_conserve(gg, n, p, mode):
    mesh, P, st, x0, dt = _setup(gg, n, p, q0_mode=mode)
    st.set_state(x0)
    st.step(2)
    x = x0.copy()
    for _ in range(2):
        x, y = P.step(x, dt)
    xg = st.get_state()
    for a, b, name in zip(P.split_x(xg), P.split_x(x), "upB"):
        assert _rel(a, b) < 1e-9, name
    assert _rel(st.get_diagnostics(), y) < 1e-8
    _, p0, B0 = P.split_x(x0)
    _, p1, B1 = P.split_x(xg)
    assert abs(P.total(p1) - P.total(p0)) < 1e-12 * abs(P.total(p0))
    assert abs(P.total(B1) - P.total(B0)) < 1e-12 * abs(P.total(B0))
