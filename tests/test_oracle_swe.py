"""Pins the shallow-water oracle (oracle/swe.py) against the reference's own regression numbers.

/root/reference/test/Transient/TransientShallowWater.jl:208-209 (Williamson test case 2, nref 3, p_fe 1, CFL 0.1, one
day, SSPRK(3,3)):  eu ~ 0.0035085422284689785, ep ~ 3.5206285360509003e-6 with rtol 1e-2 / atol 1e-3."""
import numpy as np

from oracle import swe
from oracle.mesh import CubedSphereMesh2D
from oracle.rk import gridap_num_steps

EU_REF, EP_REF = 0.0035085422284689785, 3.5206285360509003e-6


def test_reference_time_step():
    mesh = CubedSphereMesh2D(8, 1.0)
    dt = swe.reference_dt(mesh, 1)
    assert gridap_num_steps(0.0, swe._tF, dt) == 128
    assert abs(dt * 128 - swe._tF) < 1e-13


def test_williamson2_regression_numbers():
    mesh = CubedSphereMesh2D(8, 1.0)
    P = swe.ShallowWaterProblem(mesh, 1)              # q0 aliases q, as in the reference stepper
    dt = swe.reference_dt(mesh, 1)
    x0 = P.initial_state()
    x = x0.copy()
    for _ in range(gridap_num_steps(0.0, swe._tF, dt)):
        x, _y = P.step(x, dt)
    eu, ep = P.errors(x0, x)
    # the reference's own tolerance
    assert abs(eu - EU_REF) <= 1e-2 * EU_REF + 1e-3 and abs(ep - EP_REF) <= 1e-2 * EP_REF + 1e-3
    # what the restatement actually reaches: 8e-5 / 1.5e-4 relative (the atol of the reference test makes its ep check
    # vacuous; this one is not)
    assert abs(eu - EU_REF) < 2e-4 * EU_REF
    assert abs(ep - EP_REF) < 3e-4 * EP_REF


def test_alias_mode_is_what_the_reference_does():
    """40 steps are enough to tell the two q0 conventions apart: `alias` tracks the reference numbers' trend."""
    mesh = CubedSphereMesh2D(4, 1.0)
    dt = swe.reference_dt(mesh, 1)
    out = {}
    for mode in ("alias", "start_of_step"):
        P = swe.ShallowWaterProblem(mesh, 1, q0_mode=mode)
        x0 = P.initial_state()
        x = x0.copy()
        for _ in range(5):
            x, _y = P.step(x, dt)
        out[mode] = x
    d = np.abs(out["alias"] - out["start_of_step"]).max()
    assert 0 < d < 1e-2 * np.abs(out["alias"]).max()   # the (q - q0)/dt term is active, but a small correction


def test_steady_state_is_nearly_preserved_by_one_step():
    # Williamson 2 is a steady solution: one step changes the state by O(h^2 dt)
    mesh = CubedSphereMesh2D(6, 1.0)
    P = swe.ShallowWaterProblem(mesh, 1)
    x0 = P.initial_state()
    x, y = P.step(x0, swe.reference_dt(mesh, 1))
    nu = P.nu
    assert np.abs(x - x0)[:nu].max() < 3e-2 * np.abs(x0[:nu]).max()   # coarse mesh (6 x 6 per panel): 1e-2 drift
    assert np.abs(x - x0)[nu:].max() < 3e-3 * np.abs(x0[nu:]).max()
    q, F, Phi = P.split_y(y)
    assert np.isfinite(y).all() and np.abs(Phi).max() > 0
