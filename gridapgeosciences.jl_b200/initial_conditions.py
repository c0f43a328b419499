"""Initial states of the shallow-water drivers, as functions of the ambient coordinates (host side, numpy).

williamson2_*: test case 2 of Williamson et al. (1992) exactly as the reference states it
  (/root/reference/test/Geophysical/Williamson_functions.jl:1-37, non-dimensionalisation of
  test/Transient/TransientShallowWater.jl:27-46: L = a_e, time scale 1/omega).
galewsky_*: the barotropically unstable jet of Galewsky, Scott & Polvani (Tellus A 56, 2004), which BASELINE config 4 names.
  It is NOT in the reference (only a README gif caption), so it is restated from the paper and its results are
  parity unpinned; the same non-dimensionalisation is applied so that it can be fed to the same operator.
"""
import numpy as np

A_E, G_DIM, OMEGA_DIM = 6.37e6, 9.8, 7.29e-5
T_SCALE = 1.0 / OMEGA_DIM
GRAVITY = G_DIM * T_SCALE ** 2 / A_E          # _g
OMEGA = OMEGA_DIM * T_SCALE                   # _omega = 1
H0_W2 = 2.94e4 / G_DIM / A_E                  # _H_0
U0_W2 = 2 * np.pi * A_E / (12 * 24 * 3600) / A_E * T_SCALE


def _lonlat(X):
    r = np.sqrt((X ** 2).sum(-1))
    return np.arctan2(X[..., 1], X[..., 0]), np.arcsin(X[..., 2] / r)


def coriolis(X):
    """f = 2 omega sin(phi)."""
    _, ph = _lonlat(X)
    return 2.0 * OMEGA * np.sin(ph)


def _east(X):
    th, _ = _lonlat(X)
    return np.stack([-np.sin(th), np.cos(th), 0 * th], -1)


def williamson2_velocity(X):
    _, ph = _lonlat(X)
    return (U0_W2 * np.cos(ph))[..., None] * _east(X)


def williamson2_depth(X):
    _, ph = _lonlat(X)
    return H0_W2 - (OMEGA * U0_W2 + 0.5 * U0_W2 ** 2) * np.sin(ph) ** 2 / GRAVITY


# ---- Galewsky et al. (2004), eqs. (2)-(4) -------------------------------------------------------------------------------------
def thermogeostrophic_buoyancy(X, c=0.05):
    """b0 of test/Tutorials/ThermalShallowWater.jl:121-125: g (1 + c H0^2 / h^2) on the Williamson-2 depth."""
    h = williamson2_depth(X)
    return GRAVITY * (1.0 + c * H0_W2 ** 2 / h ** 2)


def thermogeostrophic_weighted_buoyancy(X, c=0.05):
    """B0 = b0 h0 (test/Tutorials/ThermalShallowWater.jl:127-131)."""
    return thermogeostrophic_buoyancy(X, c) * williamson2_depth(X)


_UMAX = 80.0 / A_E * T_SCALE
_PHI0, _PHI1 = np.pi / 7, np.pi / 2 - np.pi / 7
_EN = np.exp(-4.0 / (_PHI1 - _PHI0) ** 2)
_H_MEAN = 1.0e4 / A_E
_HHAT, _ALPHA, _BETA, _PHI2 = 120.0 / A_E, 1.0 / 3.0, 1.0 / 15.0, np.pi / 4


def _jet(ph):
    ph = np.asarray(ph, dtype=np.float64)
    inside = (ph > _PHI0) & (ph < _PHI1)
    d = np.where(inside, (ph - _PHI0) * (ph - _PHI1), -1.0)
    return np.where(inside, _UMAX / _EN * np.exp(1.0 / d), 0.0)


def galewsky_velocity(X):
    _, ph = _lonlat(X)
    return _jet(ph)[..., None] * _east(X)


_TABLE = None


def _balanced_depth_table(n=20001):
    """g h(phi) = g h0 - int_{-pi/2}^{phi} a u (f + tan(phi') u / a) dphi' (a = 1 here), h0 fixed by the global mean depth."""
    global _TABLE
    if _TABLE is None:
        ph = np.linspace(-np.pi / 2, np.pi / 2, n)
        u = _jet(ph)
        integrand = u * (2.0 * OMEGA * np.sin(ph) + np.tan(np.clip(ph, -1.5607, 1.5607)) * u)
        dp = ph[1] - ph[0]
        integ = np.concatenate([[0.0], np.cumsum(0.5 * (integrand[1:] + integrand[:-1]) * dp)])
        h = -integ / GRAVITY
        mean = np.trapezoid(h * np.cos(ph), ph) / 2.0
        _TABLE = (ph, h + (_H_MEAN - mean))
    return _TABLE


def galewsky_depth(X, perturbed=True):
    th, ph = _lonlat(X)
    tp, th_ = _balanced_depth_table()
    h = np.interp(ph, tp, th_)
    if perturbed:
        lam = np.where(th > np.pi, th - 2 * np.pi, th)
        h = h + _HHAT * np.cos(ph) * np.exp(-(lam / _ALPHA) ** 2) * np.exp(-((_PHI2 - ph) / _BETA) ** 2)
    return h
