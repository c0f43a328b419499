"""Pins the oracle's forms/assembly/RK to the reference's own tests (CPU only)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import forms, rk
from oracle.mesh import CubedSphereMesh2D
from oracle.spaces import CGSpace, DGSpace, RTSpace


def fX(X):  # test/Laplacian/LaplaceBeltrami.jl:15-17
    return X[..., 0] * X[..., 1] * X[..., 2]


def grad_fX(X):
    return np.stack([X[..., 1] * X[..., 2], X[..., 0] * X[..., 2], X[..., 0] * X[..., 1]], -1)


def hess_fX(X):
    H = np.zeros(X.shape[:-1] + (3, 3))
    H[..., 0, 1] = H[..., 1, 0] = X[..., 2]
    H[..., 0, 2] = H[..., 2, 0] = X[..., 1]
    H[..., 1, 2] = H[..., 2, 1] = X[..., 0]
    return H


def solve_lb(mesh, p, extrinsic=False):
    """laplace_beltrami_solver of test/Laplacian/LaplaceBeltrami.jl:19-98 (and its extrinsic twin)."""
    degree = 6 * (p + 1)
    V = CGSpace(mesh, p, zeromean=True)
    q = forms.CellQuadrature(mesh, degree)
    if extrinsic:
        Ke = forms.lb_extrinsic_element_matrices(V, q)
    else:
        Ke = forms.lb_element_matrices(V, q)
    A = forms.assemble_matrix(V, V, Ke)
    lap = forms.surface_laplacian(fX, grad_fX, hess_fX, mesh, q.x, mesh.cell_to_chart)
    assert abs((fX(q.ambient()) * q.dV * q.sqrtg).sum()) < 1e-14   # zero-mean check, LaplaceBeltrami.jl:43
    b = forms.assemble_vector(V, forms.source_element_vectors(V, q, -lap))
    x = spla.spsolve(A.tocsc(), b)
    Vf = CGSpace(mesh, p)
    xf = np.append(x, 0.0)
    vol_i = forms.assemble_vector(Vf, np.einsum("cq,qa->ca", q.dV, Vf.ref.tabulate(q.pts)[0]))
    xf = xf - (vol_i @ xf) / vol_i.sum()                           # Gridap ZeroMeanFESpace post-shift
    qe = forms.CellQuadrature(mesh, 2 * degree)
    uh = forms.eval_scalar(Vf, qe, xf)
    return np.sqrt(((fX(qe.ambient()) - uh) ** 2 * qe.dV * qe.sqrtg).sum()), lap, q


def test_surface_laplacian_of_xyz():
    # xyz is a degree-3 spherical harmonic: Delta_s f = -12 f / r^2 (closed form of SurfaceDiffOps.jl:26-65)
    mesh = CubedSphereMesh2D(4, 1.3)
    q = forms.CellQuadrature(mesh, 6)
    lap = forms.surface_laplacian(fX, grad_fX, hess_fX, mesh, q.x, mesh.cell_to_chart)
    assert np.abs(lap + 12 * fX(q.ambient()) / 1.3**2).max() < 1e-13


def test_laplace_beltrami_convergence_p2():
    # test/Laplacian/LaplaceBeltrami.jl:101-105 + src/ConvergenceTools/Tools.jl:4-35: slope >= p over nref 1..4
    errs, dxs = [], []
    for lvl in (1, 2, 3, 4):
        mesh = CubedSphereMesh2D(2**lvl)
        e, _, _ = solve_lb(mesh, 2)
        errs.append(e)
        dxs.append(rk.dx(mesh))
    slope = np.polyfit(np.log10(dxs), np.log10(errs), 1)[0]
    assert slope >= 2


def test_intrinsic_equals_extrinsic():
    # test/AmbientModel/seq/AmbientLaplaceBeltramiTests.jl:28-30: |e_intrinsic - e_extrinsic| < 1e-12
    mesh = CubedSphereMesh2D(8)
    ei, _, _ = solve_lb(mesh, 2)
    ee, _, _ = solve_lb(mesh, 2, extrinsic=True)
    assert abs(ei - ee) < 1e-12
    V = CGSpace(mesh, 2)
    q = forms.CellQuadrature(mesh, 18)
    Ki, Ke = forms.lb_element_matrices(V, q), forms.lb_extrinsic_element_matrices(V, q)
    assert np.abs(Ki - Ke).max() < 1e-12 * np.abs(Ki).max()


def test_element_matrix_properties():
    mesh = CubedSphereMesh2D(4, 2.0)
    V = CGSpace(mesh, 2)
    q = forms.CellQuadrature(mesh, 18)
    Ke = forms.lb_element_matrices(V, q)
    assert np.abs(Ke - np.swapaxes(Ke, 1, 2)).max() < 1e-14          # symmetric
    assert np.abs(Ke.sum(axis=2)).max() < 1e-13                       # constants in the kernel
    A = forms.assemble_matrix(V, V, Ke)
    assert A.shape == (V.num_free, V.num_free) and V.num_free == 6 * 16 * 4 + 2
    # nnz for CG-Q2 = 16*Nv_like ... check against the closed form 6n^2*(9+2*15+25)/... via pattern
    rowptr, colind = forms.csr_pattern(V, V)
    assert rowptr[-1] == A.nnz


def test_cg_q1_counts_match_survey():
    mesh = CubedSphereMesh2D(32)
    V = CGSpace(mesh, 1)
    rowptr, colind = forms.csr_pattern(V, V)
    assert V.num_free == 6146 and rowptr[-1] == 55298


def test_wave_regression_numbers():
    # test/Transient/TransientWaveEquation.jl:106-130: nref 3, p 1, CFL 0.1, tF 2 pi, SSPRK(3,3)
    mesh = CubedSphereMesh2D(8)
    W = rk.WaveProblem(mesh, 1)
    assert (W.nu, W.np_) == (3072, 1536)
    depth = lambda X: 1.0 + 0.01 * np.exp(-5 * ((1 - X[..., 0]) ** 2 + X[..., 1] ** 2 + X[..., 2] ** 2))
    x0 = np.concatenate([np.zeros(W.nu), rk.interpolate_scalar(W.Q, mesh, depth)])
    tF = 2 * np.pi
    dt = rk.snapped_dt(rk.dx(mesh) * 0.1 / 1, tF)
    nsteps = rk.gridap_num_steps(0.0, tF, dt)
    assert nsteps == 348
    x = x0.copy()
    for _ in range(nsteps):
        x = W.step(x, dt)
    eu, ep = W.errors(x0, x)
    # the reference asserts rtol 1e-2 / atol 1e-3; the restatement reproduces its numbers to ~1e-10
    assert abs(eu - 0.001602838116424487) < 1e-9 * 0.0016
    assert abs(ep - 0.010030006030668944) < 1e-9 * 0.01


def test_wave_energy_conservation_semidiscrete():
    # res is skew: x^T res(x) = 0, so d/dt (x^T M x) = 0 for the semi-discrete system
    mesh = CubedSphereMesh2D(4)
    W = rk.WaveProblem(mesh, 1)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(W.nu + W.np_)
    assert abs(x @ W.residual(x)) < 1e-12 * np.abs(x).sum()


def test_rt_interpolation_reproduces_rt_fields():
    mesh = CubedSphereMesh2D(4)
    V = RTSpace(mesh, 1)
    ref = V.ref
    # reference element: dofs_of(phi_j) = delta_ij
    pts, Wm = ref.moment_points()
    val, _ = ref.tabulate(pts)
    assert np.abs(np.einsum("mqi,qji->mj", Wm, val) - np.eye(ref.ndofs)).max() < 1e-12
    # rigid rotation about z: exactly tangential, its flux DOFs must be single-valued across panels
    vfun = lambda X: np.stack([-X[..., 1], X[..., 0], 0 * X[..., 0]], -1)
    u = rk.interpolate_rt(V, mesh, vfun)
    q = forms.CellQuadrature(mesh, 8)
    uh, div = forms.eval_rt(V, q, u)
    # div of sqrtg * contravariant rotation = 0 up to interpolation error; small at this resolution
    assert np.abs(div).max() < 5e-2
    # exact field at quadrature points
    from oracle import geometry as geo
    err = 0.0
    for p in range(6):
        sel = mesh.cell_to_chart == p
        X = geo.csphere_eval(p + 1, 1.0, q.x[sel])
        P = geo.pinvJ(geo.csphere_jac(p + 1, 1.0, q.x[sel]))
        ue = geo.csphere_sqrtg(1.0, q.x[sel])[..., None] * np.einsum("...ik,...k->...i", P, vfun(X))
        err = max(err, np.abs(ue - uh[sel]).max())
    assert err < 2e-2


def _ambient_fX(X):           # test/Autodiff/seq/SurfDiffOperatorTests.jl:25-28
    return X[..., 0] ** 2 + X[..., 1] ** 2 * X[..., 2]


def _grad_ambient_fX(X):
    x, y, z = X[..., 0], X[..., 1], X[..., 2]
    return np.stack([2 * x, 2 * y * z, y * y], axis=-1)


def _ambient_vecX(X):         # :30-33
    x, y = X[..., 0], X[..., 1]
    return np.stack([y * y, -x * x, np.zeros_like(x)], axis=-1)


def _grad_ambient_vecX(X):    # [l,k] = d F_k / d X_l
    x, y = X[..., 0], X[..., 1]
    G = np.zeros(X.shape + (3,))
    G[..., 1, 0] = 2 * y
    G[..., 0, 1] = -2 * x
    return G


def test_surface_gradient_and_divergence_equal_their_ambient_forms():
    # test/Autodiff/seq/SurfDiffOperatorTests.jl:43-62: intrinsic == extrinsic pointwise, 1e-12
    from oracle import geometry as geo
    r = 1.0
    mesh = CubedSphereMesh2D(2, r)
    q = forms.CellQuadrature(mesh, 4)
    X = q.ambient()
    n = X / r
    P = np.eye(3) - n[..., :, None] * n[..., None, :]
    sg = forms.surface_gradient(_grad_ambient_fX, mesh, q.x, mesh.cell_to_chart)
    push = np.empty(X.shape)
    for p in range(6):
        sel = mesh.cell_to_chart == p
        J = geo.csphere_jac(p + 1, r, q.x[sel])
        push[sel] = np.einsum("...ik,...i->...k", J, sg[sel])                # covariant basis . contravariant components
    assert np.abs(push - np.einsum("...kl,...l->...k", P, _grad_ambient_fX(X))).max() < 1e-12
    sd = forms.surface_divergence(_ambient_vecX, _grad_ambient_vecX, mesh, q.x, mesh.cell_to_chart)
    F, dF = _ambient_vecX(X), _grad_ambient_vecX(X)
    amb = np.einsum("...lk,...lk->...", P, dF) - 2.0 / r * np.einsum("...k,...k->...", F, n)
    assert np.abs(sd - amb).max() < 1e-12
    # divs(grad f) = Delta_s f for any ambient f (the normal part of grad f drops out)
    lap = forms.surface_laplacian(fX, grad_fX, hess_fX, mesh, q.x, mesh.cell_to_chart)
    sd2 = forms.surface_divergence(grad_fX, hess_fX, mesh, q.x, mesh.cell_to_chart)
    assert np.abs(sd2 - lap).max() < 1e-12


def solve_helmholtz(mesh, p):
    """helmholtz_solver of test/Laplacian/Helmholtz.jl:19-86: (M - K) u = int (f + Delta_s f) v sqrtg, LU solve."""
    degree = 6 * (p + 1)
    V = CGSpace(mesh, p)
    q = forms.CellQuadrature(mesh, degree)
    Ke = forms.scalar_mass_element_matrices(V, V, q) - forms.lb_element_matrices(V, q)
    A = forms.assemble_matrix(V, V, Ke)
    rhs = fX(q.ambient()) + forms.surface_laplacian(fX, grad_fX, hess_fX, mesh, q.x, mesh.cell_to_chart)
    b = forms.assemble_vector(V, forms.source_element_vectors(V, q, rhs))
    x = spla.spsolve(A.tocsc(), b)
    qe = forms.CellQuadrature(mesh, 2 * degree)
    return np.sqrt(((fX(qe.ambient()) - forms.eval_scalar(V, qe, x)) ** 2 * qe.dV * qe.sqrtg).sum())


def test_helmholtz_convergence_p2():
    # test/Laplacian/seq/HelmholtzTests.jl + src/ConvergenceTools/Tools.jl:4-35
    errs = [solve_helmholtz(CubedSphereMesh2D(2**lvl), 2) for lvl in (1, 2, 3, 4)]
    slope = np.polyfit(np.log10([1 / 2.0**lvl for lvl in (1, 2, 3, 4)]), np.log10(errs), 1)[0]
    assert slope >= 2


def sinlat(X):                # test/Laplacian/HodgeLaplacian_scalar.jl:18-21: sin of the latitude
    return X[..., 2] / np.linalg.norm(X, axis=-1)


def grad_sinlat(X):
    r = np.linalg.norm(X, axis=-1)[..., None]
    ez = np.zeros_like(X)
    ez[..., 2] = 1.0
    return ez / r - X[..., 2:3] * X / r**3


def hess_sinlat(X):
    r = np.linalg.norm(X, axis=-1)[..., None, None]
    z = X[..., 2][..., None, None]
    ez = np.zeros_like(X)
    ez[..., 2] = 1.0
    XX = X[..., :, None] * X[..., None, :]
    eX = ez[..., :, None] * X[..., None, :]
    return -(eX + np.swapaxes(eX, -1, -2)) / r**3 - z * np.eye(3) / r**3 + 3 * z * XX / r**5


def solve_hodge_scalar(mesh, p):
    """hodge_laplacian_scalar of test/Laplacian/HodgeLaplacian_scalar.jl:23-121: u + grad_s(phi) = 0, div_s u = -Delta_s f in
    mixed form on RT_p x DG_p (zero-mean pressure), LU solve; returns (velocity error, pressure error)."""
    import scipy.sparse as sp
    degree = 4 * (p + 1) if p > 0 else 10
    V, Q = RTSpace(mesh, p), DGSpace(mesh, p)
    q = forms.CellQuadrature(mesh, degree)
    M = forms.assemble_matrix(V, V, forms.rt_mass_element_matrices(V, q))
    B = forms.assemble_matrix(Q, V, forms.div_element_matrices(Q, V, q))
    lap = forms.surface_laplacian(sinlat, grad_sinlat, hess_sinlat, mesh, q.x, mesh.cell_to_chart)
    bq = forms.assemble_vector(Q, forms.source_element_vectors(Q, q, -lap))
    nq = Q.num_free - 1                                              # zeromean: the last DOF is fixed to zero
    A = sp.bmat([[M, -B.T[:, :nq]], [B[:nq], None]]).tocsc()
    x = spla.spsolve(A, np.concatenate([np.zeros(V.num_free), bq[:nq]]))
    u, ph = x[: V.num_free], np.append(x[V.num_free:], 0.0)
    vol_i = forms.assemble_vector(Q, np.einsum("cq,qa->ca", q.dV, Q.ref.tabulate(q.pts)[0]))
    ph = ph - (vol_i @ ph) / vol_i.sum()
    qe = forms.CellQuadrature(mesh, 2 * degree)
    e_p = np.sqrt(((sinlat(qe.ambient()) - forms.eval_scalar(Q, qe, ph)) ** 2 * qe.dV * qe.sqrtg).sum())
    sigma = forms.surface_gradient(grad_sinlat, mesh, qe.x, mesh.cell_to_chart)
    e = forms.eval_rt(V, qe, u)[0] / qe.sqrtg[..., None] + sigma
    e_u = np.sqrt((np.einsum("cqi,cqij,cqj->cq", e, qe.g, e) * qe.dV * qe.sqrtg).sum())
    return e_u, e_p


def test_hodge_laplacian_scalar_convergence_p2():
    # test/Laplacian/seq/HodgeLaplacianScalarTests.jl (nref 1..4, p = 2): slope of the velocity error >= p
    errs = [solve_hodge_scalar(CubedSphereMesh2D(2**lvl), 2) for lvl in (1, 2, 3)]
    lv = np.log10([1 / 2.0**lvl for lvl in (1, 2, 3)])
    assert np.polyfit(lv, np.log10([e[0] for e in errs]), 1)[0] >= 2
    assert np.polyfit(lv, np.log10([e[1] for e in errs]), 1)[0] >= 2


@pytest.mark.parametrize("k", [0, 1, 2])
def test_rt_basis_is_a_tensor_product_of_1d_dual_bases(k):
    """What the lane kernels of the shallow-water path rely on (gridapgeosciences.jl_b200/csrc/gg_swe_lanes.cuh): every moment
    functional of RT_k on the square is a tensor product of 1-D functionals -- Lambda_0 f = -f(0), Lambda_1 f = f(1),
    Lambda_{2+i} f = int x^i f (i < k) on P_{k+1} and mu_m f = int s^m f (m <= k) on P_k -- so the basis is
    phi^x_{alpha,m} = A_alpha(x) B_m(y), phi^y_{m,beta} = B_m(x) A_beta(y) with the 1-D dual bases A, B, div phi^x = A' B."""
    from oracle.refel import RaviartThomasQuad, gauss_legendre_01
    NA, NB = k + 2, k + 1
    LA = np.zeros((NA, NA))
    LA[0, 0] = -1.0
    LA[1, :] = 1.0
    for i in range(k):
        LA[2 + i, :] = 1.0 / (i + np.arange(NA) + 1.0)
    LB = 1.0 / (np.arange(NB)[:, None] + np.arange(NB)[None, :] + 1.0)
    CA, CB = np.linalg.inv(LA), np.linalg.inv(LB)            # [monomial][function]
    x, _ = gauss_legendre_01(k + 3)
    pw = lambda n: x[:, None] ** np.arange(n)[None, :]       # noqa: E731
    dpw = lambda n: np.arange(n)[None, :] * x[:, None] ** np.maximum(np.arange(n) - 1, 0)[None, :]   # noqa: E731
    A, dA, B = pw(NA) @ CA, dpw(NA) @ CA, pw(NB) @ CB        # [point][function]
    jx = lambda al, m: 2 * NB + m if al == 0 else 3 * NB + m if al == 1 else 4 * NB + m * k + (al - 2)                  # noqa: E731
    jy = lambda m, be: m if be == 0 else NB + m if be == 1 else 4 * NB + k * NB + (be - 2) * NB + m                      # noqa: E731
    X, Y = np.meshgrid(x, x, indexing="ij")
    val, div = RaviartThomasQuad(k).tabulate(np.stack([X.ravel(), Y.ravel()], -1))
    val, div = val.reshape(len(x), len(x), -1, 2), div.reshape(len(x), len(x), -1)
    seen = set()
    scale = max(1.0, np.abs(val).max(), np.abs(div).max())
    for al in range(NA):
        for m in range(NB):
            j = jx(al, m)
            seen.add(j)
            assert np.abs(val[:, :, j, 0] - np.outer(A[:, al], B[:, m])).max() < 1e-10 * scale
            assert np.abs(val[:, :, j, 1]).max() < 1e-10 * scale
            assert np.abs(div[:, :, j] - np.outer(dA[:, al], B[:, m])).max() < 1e-9 * scale
            j = jy(m, al)
            seen.add(j)
            assert np.abs(val[:, :, j, 1] - np.outer(B[:, m], A[:, al])).max() < 1e-10 * scale
            assert np.abs(val[:, :, j, 0]).max() < 1e-10 * scale
            assert np.abs(div[:, :, j] - np.outer(B[:, m], dA[:, al])).max() < 1e-9 * scale
    assert seen == set(range(2 * NA * NB))
