"""Vector Laplacian and surface Stokes oracle (oracle/surface_stokes.py) pinned to the reference's own tests of them:
test/SurfaceStokes/seq/VectorLaplacianTests_2D.jl and seq/SurfaceStokesTests.jl (p = 2, slope >= p over the refinements), plus the
closed form of the manufactured right-hand side: uX = (xz, yz, z^2 - 1) = -grad_s z on the unit sphere, so
vecΔs_2D(uX) = grad_s(div_s uX) - curl curl = grad_s(2 z) = -2 uX."""
import numpy as np

from oracle import forms
from oracle import surface_stokes as ss
from oracle import vector_h1 as vh
from oracle.mesh import CubedSphereMesh2D


def uX(X):
    return np.stack([X[..., 0] * X[..., 2], X[..., 1] * X[..., 2], X[..., 2] ** 2 - 1], axis=-1)


def uX_sym(x):
    return (x[0] * x[2], x[1] * x[2], x[2] ** 2 - 1)


def pX(X):
    return X[..., 2]


def grad_pX(X):
    out = np.zeros_like(X)
    out[..., 2] = 1.0
    return out


def grad_uX(X):
    """[l, k] = d uX_k / d X_l"""
    out = np.zeros(X.shape + (3,))
    out[..., 0, 0] = out[..., 1, 1] = X[..., 2]
    out[..., 2, 0], out[..., 2, 1], out[..., 2, 2] = X[..., 0], X[..., 1], 2 * X[..., 2]
    return out


def closed_form_rhs(q):
    return 2.0 * vh.contravariant(q.mesh, q, uX)


def test_symbolic_vector_laplacian_equals_the_closed_form_on_all_panels():
    mesh = CubedSphereMesh2D(2)
    q = forms.CellQuadrature(mesh, 6)
    lap = ss.vec_laps_2d(mesh, q, uX_sym)
    ref = -2.0 * vh.contravariant(mesh, q, uX)
    assert np.abs(lap - ref).max() < 1e-12 * np.abs(ref).max()


def test_bilinear_form_is_symmetric_and_consistent_with_the_strong_operator():
    """a(u_int, v) = l(v) up to the interpolation error: checks D1, D2 and the sign of the curl-curl term against the strong form."""
    res = []
    for n in (4, 8):
        mesh = CubedSphereMesh2D(n)
        V = vh.VectorCGSpace(mesh, 2)
        q = forms.CellQuadrature(mesh, 18)
        K = ss.vector_laplacian_element_matrices(V, q)
        assert np.abs(K - K.transpose(0, 2, 1)).max() < 1e-12 * np.abs(K).max()
        A = forms.assemble_matrix(V, V, K)
        b = forms.assemble_vector(V, vh.source_element_vectors(V, q, closed_form_rhs(q)))
        res.append(np.abs(A @ vh.interpolate(V, uX) - b).max() / np.abs(b).max())
    assert res[1] < 0.6 * res[0]


def test_vector_laplacian_driver_converges():
    lv = (1, 2, 3)
    errs = [ss.vector_laplacian_2d(CubedSphereMesh2D(2**lvl), 2, uX, rhs=closed_form_rhs)[0] for lvl in lv]
    assert np.polyfit(np.log10([1 / 2.0**lvl for lvl in lv]), np.log10(errs), 1)[0] >= 2


def test_surface_stokes_driver_converges():
    lv = (1, 2, 3)
    errs = [ss.surface_stokes(CubedSphereMesh2D(2**lvl), 2, uX, pX, uX_sym, grad_pX, grad_uX) for lvl in lv]
    h = np.log10([1 / 2.0**lvl for lvl in lv])
    assert np.polyfit(h, np.log10([e[0] for e in errs]), 1)[0] >= 2          # eu_h1: the slope the reference tests
    assert np.polyfit(h, np.log10([e[1] for e in errs]), 1)[0] >= 1.8        # Q1 pressure, pre-asymptotic on 2x2 panels
    assert np.polyfit(h, np.log10([e[2] for e in errs]), 1)[0] >= 3


def test_divergence_coupling_matches_the_closed_form_surface_divergence():
    """b(u, q) = int q D1(u) with D1(u) = div(sqrtg u) = sqrtg div_s(u): applied to the interpolant of uX it must approach
    int q div_s(uX) sqrtg, div_s from the closed-form branch that the reference's SurfDiffOperatorTests pin (oracle/forms.py)."""
    from oracle.spaces import CGSpace
    res = []
    for n in (4, 8):
        mesh = CubedSphereMesh2D(n)
        V, Q = vh.VectorCGSpace(mesh, 2), CGSpace(mesh, 1)
        q = forms.CellQuadrature(mesh, 18)
        B = forms.assemble_matrix(Q, V, ss.div_element_matrices(Q, V, q))
        sigma = forms.surface_divergence(uX, grad_uX, mesh, q.x, mesh.cell_to_chart)
        rhs = forms.assemble_vector(Q, forms.source_element_vectors(Q, q, sigma))
        res.append(np.abs(B @ vh.interpolate(V, uX) - rhs).max() / np.abs(rhs).max())
    assert res[0] < 2e-2 and res[1] < 0.3 * res[0]
