"""Oracle (test infrastructure): analytic cubed-sphere geometry, vectorised numpy.

Restates, formula by formula,
  /root/reference/src/Fields/CubedSphereMap.jl:19-49,97-123     (map, Jacobian, Hessian)
  /root/reference/src/Fields/CubedSphereMetric.jl:16-22,52-71   (g, dg)
  /root/reference/src/Fields/CubedSphereInvMetric.jl:1-7,37-56  (g^-1, dg^-1)
  /root/reference/src/Fields/CubedSphereInvMap.jl:12-44         (inverse map and its Jacobian)
  /root/reference/src/Fields/CubedSphereWithThickness{Map,Metric,InvMetric}.jl (3-D shell)
  /root/reference/src/Helpers/Operators.jl:16-24,62-74          (pinvJ, perp)
All arrays have the point index leading: x[..., 2] -> out[..., k].  Jacobians use Gridap's convention
J[..., i, k] = d phi_k / d x_i, so that g = J J^T (src/Geometry/AtlasGrids.jl:78-79).
Panels are numbered 1..6 as in the reference.
"""
import numpy as np

# (j1,s1, j2,s2, j3,s3): output component k = s_k * h[j_k]  (CubedSphereMap.jl:19-26)
CSPHERE_PERM = (
    (1, 1, 2, 1, 3, 1),
    (3, -1, 2, 1, 1, 1),
    (2, -1, 1, 1, 3, 1),
    (1, -1, 3, 1, 2, 1),
    (2, -1, 3, 1, 1, -1),
    (3, -1, 1, -1, 2, 1),
)
# (i1,sg1,i2,sg2,i3,sg3): h_m = sg_m * xyz[i_m]  (CubedSphereInvMap.jl:12-19)
CSPHERE_INV_PERM = (
    (1, 1, 2, 1, 3, 1),
    (3, 1, 2, 1, 1, -1),
    (2, 1, 1, -1, 3, 1),
    (1, -1, 3, 1, 2, 1),
    (3, -1, 1, -1, 2, 1),
    (2, -1, 3, 1, 1, -1),
)
CUBE_HALF_EDGE = np.pi / 4  # src/Geometry/CubedSphereMesh.jl:23


def _perm(panel):
    j1, s1, j2, s2, j3, s3 = CSPHERE_PERM[panel - 1]
    return (j1 - 1, j2 - 1, j3 - 1), (s1, s2, s3)


def csphere_eval(panel, r, x):
    """phi_p(alpha,beta) -- CubedSphereMap.jl:28-34."""
    x = np.asarray(x, dtype=np.float64)
    a, b = np.tan(x[..., 0]), np.tan(x[..., 1])
    rho = np.sqrt(1 + a**2 + b**2)
    h = (1 / rho, a / rho, b / rho)
    j, s = _perm(panel)
    return r * np.stack([s[0] * h[j[0]], s[1] * h[j[1]], s[2] * h[j[2]]], axis=-1)


def csphere_jac(panel, r, x):
    """J[..., i, k] = d phi_k / d x_i -- CubedSphereMap.jl:36-49."""
    x = np.asarray(x, dtype=np.float64)
    a, b = np.tan(x[..., 0]), np.tan(x[..., 1])
    rho2 = 1 + a**2 + b**2
    rho3 = rho2 * np.sqrt(rho2)
    sa, sb = 1 + a**2, 1 + b**2
    c = ((-a * sa, -b * sb), (sa * sb, -a * b * sb), (-a * b * sa, sa * sb))
    j, s = _perm(panel)
    J = np.empty(x.shape[:-1] + (2, 3))
    for k in range(3):
        for i in range(2):
            J[..., i, k] = (r / rho3) * s[k] * c[j[k]][i]
    return J


def csphere_hess(panel, r, x):
    """T[..., l, i, k] = d^2 phi_k / (d x_i d x_l) -- CubedSphereMap.jl:97-123."""
    x = np.asarray(x, dtype=np.float64)
    a, b = np.tan(x[..., 0]), np.tan(x[..., 1])
    sa, sb = 1 + a**2, 1 + b**2
    rho2 = 1 + a**2 + b**2
    rho5 = rho2**2 * np.sqrt(rho2)
    D = rho2 + 3 * a**2 * b**2
    E = 3 * a * b * sa * sb
    F = a * sa * sb * (2 * b**2 - sa)
    G = b * sa * sb * (2 * a**2 - sb)
    n = ((-sa * D, E, E, -sb * D), (F, G, G, -a * sb * D), (-b * sa * D, F, F, G))
    j, s = _perm(panel)
    fac = r / rho5
    T = np.empty(x.shape[:-1] + (2, 2, 3))
    for k in range(3):
        nj = n[j[k]]
        # column-major [l,i] with l fastest: entries (N[1,1], N[2,1], N[1,2], N[2,2])
        T[..., 0, 0, k] = fac * s[k] * nj[0]
        T[..., 1, 0, k] = fac * s[k] * nj[1]
        T[..., 0, 1, k] = fac * s[k] * nj[2]
        T[..., 1, 1, k] = fac * s[k] * nj[3]
    return T


def _sym2(g11, g12, g22):
    out = np.empty(np.shape(g11) + (2, 2))
    out[..., 0, 0] = g11
    out[..., 0, 1] = g12
    out[..., 1, 0] = g12
    out[..., 1, 1] = g22
    return out


def csphere_metric(r, x):
    """g (2x2) -- CubedSphereMetric.jl:16-22."""
    x = np.asarray(x, dtype=np.float64)
    a, b = np.tan(x[..., 0]), np.tan(x[..., 1])
    sa, sb = 1 + a**2, 1 + b**2
    rho4 = (1 + a**2 + b**2) ** 2
    c = r**2 * sa * sb / rho4
    return _sym2(c * sa, -c * a * b, c * sb)


def csphere_inv_metric(r, x):
    """g^-1 (2x2) -- CubedSphereInvMetric.jl:1-7."""
    x = np.asarray(x, dtype=np.float64)
    a, b = np.tan(x[..., 0]), np.tan(x[..., 1])
    sa, sb = 1 + a**2, 1 + b**2
    rho2 = 1 + a**2 + b**2
    c = rho2 / (r**2 * sa * sb)
    return _sym2(c * sb, c * a * b, c * sa)


def csphere_sqrtg(r, x):
    """MeasureCellField = sqrt(det(g)) with the generic det of the 2x2 metric (CellFields.jl:80-82)."""
    g = csphere_metric(r, x)
    return np.sqrt(g[..., 0, 0] * g[..., 1, 1] - g[..., 0, 1] * g[..., 1, 0])


def csphere_metric_grad(r, x):
    """T[..., k, i, j] = d g_ij / d x_k -- CubedSphereMetric.jl:52-71."""
    x = np.asarray(x, dtype=np.float64)
    a, b = np.tan(x[..., 0]), np.tan(x[..., 1])
    sa, sb = 1 + a**2, 1 + b**2
    rho2 = 1 + a**2 + b**2
    c6 = r**2 / rho2**3
    d11a = 4 * c6 * a * sa**2 * sb * b**2
    d11b = 2 * c6 * b * sa**2 * sb * (a**2 - sb)
    d12a = -c6 * b * sa * sb * (rho2 + a**2 * (3 * b**2 - sa))
    d12b = -c6 * a * sa * sb * (rho2 + b**2 * (3 * a**2 - sb))
    d22a = 2 * c6 * a * sa * sb**2 * (b**2 - sa)
    d22b = 4 * c6 * b * sa * sb**2 * a**2
    T = np.empty(x.shape[:-1] + (2, 2, 2))
    T[..., 0, :, :] = _sym2(d11a, d12a, d22a)
    T[..., 1, :, :] = _sym2(d11b, d12b, d22b)
    return T


def csphere_inv_metric_grad(r, x):
    """T[..., k, i, j] = d (g^-1)_ij / d x_k -- CubedSphereInvMetric.jl:37-56."""
    x = np.asarray(x, dtype=np.float64)
    a, b = np.tan(x[..., 0]), np.tan(x[..., 1])
    sa, sb = 1 + a**2, 1 + b**2
    rho2 = 1 + a**2 + b**2
    ci = 1 / (r**2 * sa * sb)
    d11a = -2 * ci * a * b**2 * sb
    d11b = 2 * ci * b * sb**2
    d12a = ci * b * (sa * rho2 - 2 * a**2 * b**2)
    d12b = ci * a * (sb * rho2 - 2 * a**2 * b**2)
    d22a = 2 * ci * a * sa**2
    d22b = -2 * ci * b * a**2 * sa
    T = np.empty(x.shape[:-1] + (2, 2, 2))
    T[..., 0, :, :] = _sym2(d11a, d12a, d22a)
    T[..., 1, :, :] = _sym2(d11b, d12b, d22b)
    return T


def csphere_inv_eval(panel, xyz):
    """(alpha,beta) from a sphere point -- CubedSphereInvMap.jl:21-27."""
    xyz = np.asarray(xyz, dtype=np.float64)
    i1, g1, i2, g2, i3, g3 = CSPHERE_INV_PERM[panel - 1]
    h1 = g1 * xyz[..., i1 - 1]
    h2 = g2 * xyz[..., i2 - 1]
    h3 = g3 * xyz[..., i3 - 1]
    return np.stack([np.arctan2(h2, h1), np.arctan2(h3, h1)], axis=-1)


def csphere_inv_jac(panel, xyz):
    """J[..., i, k] = d (alpha,beta)_k / d xyz_i -- CubedSphereInvMap.jl:32-44."""
    xyz = np.asarray(xyz, dtype=np.float64)
    i1, g1, i2, g2, i3, g3 = CSPHERE_INV_PERM[panel - 1]
    h1 = g1 * xyz[..., i1 - 1]
    h2 = g2 * xyz[..., i2 - 1]
    h3 = g3 * xyz[..., i3 - 1]
    r12 = h1**2 + h2**2
    r13 = h1**2 + h3**2
    J = np.zeros(xyz.shape[:-1] + (3, 2))
    J[..., i1 - 1, 0] = -h2 * g1 / r12
    J[..., i2 - 1, 0] = h1 * g2 / r12
    J[..., i1 - 1, 1] = -h3 * g1 / r13
    J[..., i3 - 1, 1] = h1 * g3 / r13
    return J


# ---- 3-D shell (chart coordinate order (gamma, alpha, beta)) -------------------------------------

def csphere3d_eval(panel, r, t, x):
    """CubedSphereWithThicknessMap.jl:25-29."""
    x = np.asarray(x, dtype=np.float64)
    surf = csphere_eval(panel, r, x[..., 1:3])
    return (1 + t * x[..., 0:1] / r) * surf


def csphere3d_jac(panel, r, t, x):
    """J[..., i, k] (3x3) -- CubedSphereWithThicknessMap.jl:31-44."""
    x = np.asarray(x, dtype=np.float64)
    surf = csphere_eval(panel, r, x[..., 1:3])
    Js = csphere_jac(panel, r, x[..., 1:3])
    scale = 1 + (t / r) * x[..., 0]
    J = np.empty(x.shape[:-1] + (3, 3))
    J[..., 0, :] = (t / r) * surf
    J[..., 1, :] = scale[..., None] * Js[..., 0, :]
    J[..., 2, :] = scale[..., None] * Js[..., 1, :]
    return J


def csphere3d_metric(r, t, x):
    """Block-diagonal g -- CubedSphereWithThicknessMetric.jl:24-35."""
    x = np.asarray(x, dtype=np.float64)
    gs = csphere_metric(r, x[..., 1:3])
    s2 = (1 + t * x[..., 0] / r) ** 2
    g = np.zeros(x.shape[:-1] + (3, 3))
    g[..., 0, 0] = t**2
    g[..., 1:, 1:] = s2[..., None, None] * gs
    return g


def csphere3d_inv_metric(r, t, x):
    """CubedSphereWithThicknessInvMetric.jl:1-12."""
    x = np.asarray(x, dtype=np.float64)
    gi = csphere_inv_metric(r, x[..., 1:3])
    s2 = (1 + t * x[..., 0] / r) ** 2
    g = np.zeros(x.shape[:-1] + (3, 3))
    g[..., 0, 0] = 1 / t**2
    g[..., 1:, 1:] = gi / s2[..., None, None]
    return g


def csphere3d_sqrtg(r, t, x):
    return np.sqrt(np.linalg.det(csphere3d_metric(r, t, x)))


# ---- helpers -------------------------------------------------------------------------------------

def pinvJ(Jcov):
    """pinvJ(J) with J = transpose(grad(map)) i.e. J[..., k, i] (Da x Dc): inv(J^T J) J^T -- Operators.jl:16-24.

    Takes the Gridap-convention gradient Jg[..., i, k] and returns P[..., i, k] (Dc x Da)."""
    Jg = np.asarray(Jcov)
    J = np.swapaxes(Jg, -1, -2)  # (Da, Dc)
    if J.shape[-1] == J.shape[-2]:
        return np.linalg.inv(J)
    Jt = Jg
    return np.linalg.solve(Jt @ J, Jt)


def perp(v):
    """R v = (-v2, v1) -- Operators.jl:62-64."""
    v = np.asarray(v)
    return np.stack([-v[..., 1], v[..., 0]], axis=-1)


def xyz2thetaphir(x):
    """src/Helpers/CoordinateMappings.jl:4-9."""
    x = np.asarray(x)
    r = np.sqrt(x[..., 0] ** 2 + x[..., 1] ** 2 + x[..., 2] ** 2)
    return np.arctan2(x[..., 1], x[..., 0]), np.arcsin(x[..., 2] / r), r
