"""Oracle (test infrastructure): vector Laplacian and surface Stokes on the vector-valued H1 space of the intrinsic cubed sphere.

Restates
  /root/reference/test/SurfaceStokes/VectorLaplacian2D.jl:38-80
        D1(u) = div(u) sqrtg + u.grad(sqrtg)                         ("divg_product_rule": div(sqrtg u))
        D2(u) = -grad2curl(grad(g.u)),  grad(g.u)[i,j] = sum_k u^k d_i g_jk + d_i u^k g_kj   ("divergenceRgu": div(R g u))
        a(u,v) = int (D1(u) D1(v) - D2(u) D2(v)) / sqrtg,   l(v) = int (rhs.(g v)) sqrtg,  rhs = -vecΔs_2D(uX)
        errors with Measure(Omega, 2 degree), degree = 6 (p + 1)
  /root/reference/test/SurfaceStokes/SurfaceStokes_TaylorHood.jl:40-123
        a((u,p),(v,q)) = nu a_lap(u,v) + alpha int (u.(g v)) sqrtg - int p D1(v) + int q D1(u)
        l((v,q)) = int (rhs.(g v)) sqrtg + int sigma q sqrtg,   rhs = -nu vecΔs_2D(uX) + alpha u + grad_s(pX),  sigma = divs(uX)
        Q2 vector x Q1 scalar zero-mean (Taylor-Hood), errors with Measure(Omega, 3 degree)
  /root/reference/src/CellData/SurfaceDiffOps.jl:385-392,455-466  vec_laps_2D = g^-1 grad(divs u) - (1/sqrtg) perp(grad(skew_surfdiv(u)))
  /root/reference/src/Helpers/Operators.jl:36-46                  surfdiv, skew_surfdiv
  /root/reference/src/CellData/CellFields.jl (deriv_det, deriv_sqrt, cpAB): grad(sqrtg)_k = 1/(2 sqrtg) sum_ij cof(g)_ij d_k g_ij
The reference evaluates vec_laps_2D by forward-mode AD of the composed closures; here the same chain of definitions is
differentiated symbolically (sympy) -- the manufactured field must be written with plain arithmetic.
Pins: test/SurfaceStokes/seq/VectorLaplacianTests_2D.jl and seq/SurfaceStokesTests.jl (p = 2, nref 1..4, slope >= p); for
uX = (xz, yz, z^2 - 1) = -grad_s z on the unit sphere, vec_laps_2D(uX) = -2 uX in closed form (checked in tests/).
"""
import functools

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import forms
from . import geometry as geo
from . import vector_h1 as vh
from .spaces import CGSpace


# ---- manufactured right-hand side: the reference's chain of definitions, symbolically -----------------------------------

def _symbolic_panel(panel, r):
    """Symbols a = tan(alpha), b = tan(beta): d/dalpha = (1 + a^2) d/da, so every quantity is a rational function of (a, b, rho)."""
    import sympy as sy
    a, b = sy.symbols("a b", real=True)
    sa, sb, rho2 = 1 + a**2, 1 + b**2, 1 + a**2 + b**2
    rho = sy.sqrt(rho2)
    h = (1 / rho, a / rho, b / rho)
    j, s = geo._perm(panel)
    phi = sy.Matrix([r * s[k] * h[j[k]] for k in range(3)])
    d = (lambda f: sa * sy.diff(f, a), lambda f: sb * sy.diff(f, b))
    J = sy.Matrix(3, 2, lambda k, i: d[i](phi[k]))           # J[k,i] = d phi_k / d x_i
    c = r**2 * sa * sb / rho2**2                             # CubedSphereMetric.jl:16-22
    g = sy.Matrix([[c * sa, -c * a * b], [-c * a * b, c * sb]])
    ci = rho2 / (r**2 * sa * sb)                             # CubedSphereInvMetric.jl:1-7
    ginv = sy.Matrix([[ci * sb, ci * a * b], [ci * a * b, ci * sa]])
    sqrtg = r**2 * sa * sb / rho**3                          # sqrt(det g)
    return sy, (a, b), d, phi, J, g, ginv, sqrtg


@functools.lru_cache(maxsize=None)
def _vec_laps_2d_callable(panel, r, uX_sym):
    sy, ab, d, phi, J, g, ginv, sqrtg = _symbolic_panel(panel, r)
    detg = sqrtg**2
    u_amb = sy.Matrix(uX_sym(phi))
    contra = ginv * (J.T * u_amb)                            # pinvJ(J) . u(phi(x))   (Operators.jl:16-24)
    div = lambda v: d[0](v[0]) + d[1](v[1])
    grad = lambda f: sy.Matrix([d[0](f), d[1](f)])
    perp = lambda v: sy.Matrix([-v[1], v[0]])
    divs = div(sqrtg * contra) / sqrtg                       # SurfaceDiffOps.jl:385-386
    contra_grads_divs = ginv * grad(divs)                    # :455-456
    skew = -div(detg * (ginv * perp(contra))) / sqrtg        # Operators.jl:44-46
    curly_curl = perp(grad(skew)) / sqrtg                    # SurfaceDiffOps.jl:458
    out = contra_grads_divs - curly_curl                     # :459
    return sy.lambdify(ab, [out[0], out[1]], modules="numpy", cse=True)


def vec_laps_2d(mesh, quad, uX_sym):
    """contravariant components of vecΔs_2D(uX) at the quadrature points, (Nc, Q, 2)."""
    out = np.empty(quad.x.shape)
    for p in range(6):
        sel = mesh.cell_to_chart == p
        f = _vec_laps_2d_callable(p + 1, float(mesh.radius), uX_sym)
        ta, tb = np.tan(quad.x[sel][..., 0]), np.tan(quad.x[sel][..., 1])
        v0, v1 = f(ta, tb)
        out[sel] = np.stack([v0 + 0 * ta, v1 + 0 * ta], axis=-1)
    return out


# ---- forms ------------------------------------------------------------------------------------------------------------

def grad_sqrtg(quad):
    """(Nc, Q, 2): grad_meas_cf = deriv_sqrt(det g) * cpAB(deriv_det(g), grad(g))."""
    g = quad.g
    dg = geo.csphere_metric_grad(quad.mesh.radius, quad.x)                     # [k, i, j]
    cof = np.empty_like(g)
    cof[..., 0, 0], cof[..., 1, 1] = g[..., 1, 1], g[..., 0, 0]
    cof[..., 0, 1], cof[..., 1, 0] = -g[..., 1, 0], -g[..., 0, 1]
    return np.einsum("cqij,cqkij->cqk", cof, dg) / (2 * quad.sqrtg[..., None])


def reference_d1_d2(V, quad):
    """D1, D2 of the reference (untransformed) vector shape functions N_a e_i, local order a + nn i: (Nc, Q, 2 nn) each."""
    S = V.scalar
    val, grad = forms.scalar_basis(S, quad)                                    # (Q, nn), (Nc, Q, nn, 2)
    nn = V.nn
    g = quad.g
    dg = geo.csphere_metric_grad(quad.mesh.radius, quad.x)
    gs = grad_sqrtg(quad)
    Nc, Q = quad.x.shape[:2]
    d1 = np.empty((Nc, Q, 2 * nn))
    d2 = np.empty((Nc, Q, 2 * nn))
    for i in range(2):
        d1[..., i * nn:(i + 1) * nn] = grad[..., i] * quad.sqrtg[..., None] + val[None] * gs[..., i, None]
        # w = g.u, w_j = g_ji N_a ;  curl = d_1 w_2 - d_2 w_1
        c = (g[..., 1, i, None] * grad[..., 0] - g[..., 0, i, None] * grad[..., 1]
             + val[None] * (dg[..., 0, 1, i] - dg[..., 1, 0, i])[..., None])
        d2[..., i * nn:(i + 1) * nn] = -c
    return d1, d2


def vector_laplacian_element_matrices(V, quad, nu=1.0):
    d1, d2 = reference_d1_d2(V, quad)
    w = nu * quad.dV / quad.sqrtg
    K = np.einsum("cq,cqa,cqb->cab", w, d1, d1) - np.einsum("cq,cqa,cqb->cab", w, d2, d2)
    C = V.cell_matrix()
    return np.einsum("cki,ckl,clj->cij", C, K, C)


def div_element_matrices(Q, V, quad):
    """int q D1(u): rows = scalar test functions, columns = vector trial functions."""
    d1, _ = reference_d1_d2(V, quad)
    vq = Q.ref.tabulate(quad.pts)[0]
    B = np.einsum("cq,qa,cqb->cab", quad.dV, vq, d1)
    return np.einsum("cal,clj->caj", B, V.cell_matrix())


def grad_eval(V, quad, U):
    """chart gradient of the FE function, (Nc, Q, 2 [d_i], 2 [component])."""
    _, grad = forms.scalar_basis(V.scalar, quad)
    Ue = U[V.cell_dofs].reshape(-1, 2, V.nn).transpose(0, 2, 1)
    loc = np.einsum("caij,caj->cai", V.cob, Ue)
    return np.einsum("cqak,cai->cqki", grad, loc)


def h1_error(V, quad, U, Uref):
    """sqrt(int e.(g e) sqrtg + grad(e) : (ginv.grad(e)) sqrtg) between two FE functions (u_int - uh)."""
    e = V.eval(quad, U - Uref)
    G = grad_eval(V, quad, U - Uref)
    w = quad.dV * quad.sqrtg
    m = (np.einsum("cqi,cqij,cqj->cq", e, quad.g, e) * w).sum()
    s = (np.einsum("cqki,cqkl,cqli->cq", G, quad.ginv, G) * w).sum()
    return np.sqrt(m + s)


def fe_l2_error(V, quad, U, Uref):
    e = V.eval(quad, U - Uref)
    return np.sqrt((np.einsum("cqi,cqij,cqj->cq", e, quad.g, e) * quad.dV * quad.sqrtg).sum())


# ---- drivers ----------------------------------------------------------------------------------------------------------

def vector_laplacian_2d(mesh, p, uX, uX_sym=None, rhs=None, error_degree=None):
    """VectorLaplacian2D.vector_laplacian2d: returns (eu_l2, eu_h1, space, DOF vector)."""
    degree = 6 * (p + 1)
    V = vh.VectorCGSpace(mesh, p)
    q = forms.CellQuadrature(mesh, degree)
    f = -vec_laps_2d(mesh, q, uX_sym) if rhs is None else rhs(q)
    A = forms.assemble_matrix(V, V, vector_laplacian_element_matrices(V, q))
    b = forms.assemble_vector(V, vh.source_element_vectors(V, q, f))
    U = spla.spsolve(A.tocsc(), b)
    Ui = vh.interpolate(V, uX)
    qe = forms.CellQuadrature(mesh, error_degree or 2 * degree)
    return fe_l2_error(V, qe, U, Ui), h1_error(V, qe, U, Ui), V, U


def surface_stokes(mesh, p, uX, pX, uX_sym, grad_pX, grad_uX, alpha=1.0, nu=1.0, error_degree=None):
    """SurfaceStokes_TaylorHood.surface_stokes: returns (eu_h1, ep_l2, eu_l2).

    grad_pX(X): ambient gradient of pX; grad_uX(X)[l, k] = d uX_k / d X_l (for divs(uX), closed-form branch)."""
    from .rk import interpolate_scalar
    degree = 6 * (p + 1)
    V = vh.VectorCGSpace(mesh, p)
    Qs, Qf = CGSpace(mesh, p - 1, zeromean=True), CGSpace(mesh, p - 1)
    q = forms.CellQuadrature(mesh, degree)
    u_c = vh.contravariant(mesh, q, uX)
    rhs = -nu * vec_laps_2d(mesh, q, uX_sym) + alpha * u_c + forms.surface_gradient(grad_pX, mesh, q.x, mesh.cell_to_chart)
    sigma = forms.surface_divergence(uX, grad_uX, mesh, q.x, mesh.cell_to_chart)
    Kuu = nu * vector_laplacian_element_matrices(V, q) + alpha * vh.mass_element_matrices(V, q)
    Auu = forms.assemble_matrix(V, V, Kuu)
    Aqu = forms.assemble_matrix(Qs, V, div_element_matrices(Qs, V, q))
    bu = forms.assemble_vector(V, vh.source_element_vectors(V, q, rhs))
    bq = forms.assemble_vector(Qs, forms.source_element_vectors(Qs, q, sigma))
    A = sp.bmat([[Auu, -Aqu.T], [Aqu, None]], format="csc")
    x = spla.spsolve(A, np.concatenate([bu, bq]))
    U, P = x[:V.num_free], np.append(x[V.num_free:], 0.0)          # zeromean: the last DOF is the fixed one
    # ZeroMeanFESpace: both the solution and the interpolant are shifted to zero (chart-measure) mean
    vol_i = forms.assemble_vector(Qf, np.einsum("cq,qa->ca", q.dV, Qf.ref.tabulate(q.pts)[0]))
    Pi = interpolate_scalar(Qf, mesh, pX)
    P = P - (vol_i @ P) / vol_i.sum()
    Pi = Pi - (vol_i @ Pi) / vol_i.sum()
    Ui = vh.interpolate(V, uX)
    qe = forms.CellQuadrature(mesh, error_degree or 3 * degree)
    ep = np.sqrt((forms.eval_scalar(Qf, qe, Pi - P) ** 2 * qe.dV * qe.sqrtg).sum())
    return h1_error(V, qe, U, Ui), ep, fe_l2_error(V, qe, U, Ui)
