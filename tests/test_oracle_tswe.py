"""Oracle of the thermal shallow-water tutorial (test/Tutorials/ThermalShallowWater.jl): no reference numbers exist for it, so
the restatement is checked through the properties the tutorial's own set-up implies."""
import numpy as np
import pytest

from oracle import rk, swe, tswe
from oracle.mesh import CubedSphereMesh2D


@pytest.fixture(scope="module")
def prob():
    mesh = CubedSphereMesh2D(4)
    P = tswe.ThermalShallowWaterProblem(mesh, 1)
    return mesh, P, P.initial_state()


def test_diagnostics_reproduce_the_analytic_fields(prob):
    mesh, P, x0 = prob
    y = P.diagnostics(x0)
    qv, Fv, Pv, bv = P.split_y(y)
    # b = B / p: the projection of the analytic buoyancy up to discretisation error
    from oracle.rk import interpolate_scalar
    b_exact = interpolate_scalar(P.Q, mesh, tswe.buoyancy())
    assert np.abs(bv - b_exact).max() < 2e-2 * np.abs(b_exact).max()
    assert np.isfinite(y).all()


def test_thermogeostrophic_balance_is_steady_and_conservative(prob):
    mesh, P, x0 = prob
    dt = rk.dx(mesh) * 0.05 / np.sqrt(swe._g * swe._H_0)
    x = x0.copy()
    for _ in range(3):
        x, _ = P.step(x, dt)
    u0, p0, B0 = P.split_x(x0)
    u, p, B = P.split_x(x)
    # discrete mass and weighted buoyancy are conserved to round-off (closed surface; the skeleton terms telescope)
    assert abs(P.total(p) - P.total(p0)) < 1e-12 * abs(P.total(p0))
    assert abs(P.total(B) - P.total(B0)) < 1e-12 * abs(P.total(B0))
    # the balanced state only drifts by discretisation error on this coarse mesh
    assert np.abs(p - p0).max() < 1e-2 * np.abs(p0).max()
    assert np.abs(B - B0).max() < 1e-2 * np.abs(B0).max()
    assert np.abs(u - u0).max() < 5e-2 * np.abs(u0).max()


def test_constant_buoyancy_reduces_to_shallow_water():
    """With b = g_r constant (B = g_r p) the thermal terms collapse: Phi + b th = u.u/2 + g_r p, i.e. rotating shallow water."""
    mesh = CubedSphereMesh2D(3)
    T = tswe.ThermalShallowWaterProblem(mesh, 1, degree=8)
    S = swe.ShallowWaterProblem(mesh, 1, degree=8)
    xs = S.initial_state()
    rng = np.random.default_rng(0)
    xs = xs * (1 + 1e-3 * rng.standard_normal(xs.shape))
    xt = np.concatenate([xs, swe._g * xs[S.nu:]])
    dt = rk.dx(mesh) * 0.05 / np.sqrt(swe._g * swe._H_0)
    xs1, _ = S.step(xs, dt)
    xt1, _ = T.step(xt, dt)
    u, p, B = T.split_x(xt1)
    assert np.abs(u - xs1[: S.nu]).max() < 1e-10 * np.abs(xs1[: S.nu]).max()
    assert np.abs(p - xs1[S.nu:]).max() < 1e-10 * np.abs(xs1[S.nu:]).max()
    assert np.abs(B - swe._g * p).max() < 1e-10 * np.abs(B).max()
