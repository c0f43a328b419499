"""Oracle (test infrastructure): cell-wise quadrature, element matrices/vectors and sparse assembly.

This is the arithmetic the reference leaves to Gridap's `Measure`/`IntegrationMap`/`SparseMatrixAssembler`
(SURVEY.md App. B.3/B.4/B.7), applied to the weak forms written in the reference's drivers:
  Laplace-Beltrami   a = int grad(v).(ginv.grad(u)) sqrtg, l = int rhs v sqrtg      test/Laplacian/LaplaceBeltrami.jl:47,53
  extrinsic LB       a = int grad_G(v).grad_G(u) dG                                test/AmbientModel/AmbientLaplaceBeltrami.jl:44
  wave               mass, res                                                     test/Transient/TransientWaveEquation.jl:63-66
  shallow water      res_y, jac_y, mass, res_x                                     test/Transient/TransientShallowWater.jl:113-139
Straightforward einsum over (cell, quadrature point, dof) -- deliberately NOT the factorised
algorithm of the CUDA kernels, so the two are independent.
"""
import numpy as np
import scipy.sparse as sp
from . import geometry as geo
from .refel import tensor_quadrature


class CellQuadrature:
    """Measure(Omega, degree): per-cell quadrature points in chart space + reference->chart Jacobians."""

    def __init__(self, mesh, degree):
        self.mesh, self.degree = mesh, degree
        self.pts, self.w = tensor_quadrature(mesh.D, degree)
        xi, eta = self.pts[:, 0], self.pts[:, 1]
        N = np.stack([(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta], axis=1)      # (Q,4)
        dN = np.stack([
            np.stack([-(1 - eta), (1 - eta), -eta, eta], axis=1),
            np.stack([-(1 - xi), -xi, (1 - xi), xi], axis=1)], axis=2)                               # (Q,4,2): d/dxi_i
        c = mesh.cell_chart_coords                                                                    # (Nc,4,2)
        self.x = np.einsum("qk,ckd->cqd", N, c)                                                       # chart points
        self.Jt = np.einsum("qki,ckd->cqid", dN, c)             # Jt[i,d] = d x_d / d xi_i (Gridap grad(cell_map))
        self.detJ = np.linalg.det(self.Jt)
        self.invJt = np.linalg.inv(self.Jt)
        self.dV = self.w[None, :] * np.abs(self.detJ)           # w_q * meas(J)
        r = mesh.radius
        self.g = geo.csphere_metric(r, self.x)
        self.ginv = geo.csphere_inv_metric(r, self.x)
        self.sqrtg = geo.csphere_sqrtg(r, self.x)

    def ambient(self):
        """xyz at the quadrature points (AmbientMapCellField)."""
        out = np.empty(self.x.shape[:2] + (3,))
        for p in range(6):
            sel = self.mesh.cell_to_chart == p
            out[sel] = geo.csphere_eval(p + 1, self.mesh.radius, self.x[sel])
        return out

    def ambient_jac(self):
        out = np.empty(self.x.shape[:2] + (2, 3))
        for p in range(6):
            sel = self.mesh.cell_to_chart == p
            out[sel] = geo.csphere_jac(p + 1, self.mesh.radius, self.x[sel])
        return out


# ---- basis push-forward ---------------------------------------------------------------------------

def scalar_basis(space, quad):
    """values (Q,m) and chart-space gradients (Nc,Q,m,2): grad = inv(Jt) . refgrad."""
    val, rg = space.ref.tabulate(quad.pts)
    grad = np.einsum("cqij,qmj->cqmi", quad.invJt, rg)
    return val, grad


def rt_basis(space, quad):
    """Contravariant Piola: u = (1/det J) uhat.Jt, div u = divhat/det J; slave-facet sign flips applied."""
    val, div = space.ref.tabulate(quad.pts)                       # (Q,m,2), (Q,m)
    u = np.einsum("qmi,cqid->cqmd", val, quad.Jt) / quad.detJ[:, :, None, None]
    d = div[None, :, :] / quad.detJ[:, :, None]
    s = space.signs
    return u * s[:, None, :, None], d * s[:, None, :]


# ---- element matrices ------------------------------------------------------------------------------

def lb_element_matrices(space, quad):
    _, grad = scalar_basis(space, quad)
    flux = np.einsum("cqij,cqmj->cqmi", quad.ginv, grad)
    return np.einsum("cq,cqai,cqbi->cab", quad.dV * quad.sqrtg, grad, flux)


def lb_extrinsic_element_matrices(space, quad):
    """Extrinsic manifold: cell_map = ambient o chart; 2x3 Jacobian, meas = sqrt(det(J J^T)), pseudo-inverse gradient."""
    _, rg = space.ref.tabulate(quad.pts)
    Jamb = quad.ambient_jac()                                      # (Nc,Q,2,3)
    J = np.einsum("cqij,cqjk->cqik", quad.Jt, Jamb)                # d xyz_k / d xi_i
    G = np.einsum("cqik,cqjk->cqij", J, J)
    meas = np.sqrt(np.linalg.det(G))
    pinv = np.einsum("cqji,cqjk->cqik", J, np.linalg.inv(G))       # J^T (J J^T)^-1 : (3,2)
    grad = np.einsum("cqki,qmi->cqmk", pinv, rg)                   # (Nc,Q,m,3)
    return np.einsum("cq,cqak,cqbk->cab", quad.w[None, :] * meas, grad, grad)


def scalar_mass_element_matrices(test, trial, quad, coef=None):
    """int q p sqrtg [* coef]."""
    vt = test.ref.tabulate(quad.pts)[0]
    vu = trial.ref.tabulate(quad.pts)[0]
    wgt = quad.dV * quad.sqrtg
    if coef is not None:
        wgt = wgt * coef
    return np.einsum("cq,qa,qb->cab", wgt, vt, vu)


def rt_mass_element_matrices(space, quad):
    """int (v . (g u)) / sqrtg."""
    u, _ = rt_basis(space, quad)
    gu = np.einsum("cqij,cqmj->cqmi", quad.g, u)
    return np.einsum("cq,cqai,cqbi->cab", quad.dV / quad.sqrtg, u, gu)


def div_element_matrices(scalar_space, rt_space, quad):
    """int q (div u)  (rows: scalar test, cols: RT trial)."""
    q = scalar_space.ref.tabulate(quad.pts)[0]
    _, d = rt_basis(rt_space, quad)
    return np.einsum("cq,qa,cqb->cab", quad.dV, q, d)


# ---- element vectors / evaluation -----------------------------------------------------------------

def gather(space, x, dirichlet=0.0):
    ids = space.cell_dofs
    xe = np.where(ids >= 0, x[np.maximum(ids, 0)], dirichlet)
    return xe


def eval_scalar(space, quad, x):
    return np.einsum("qm,cm->cq", space.ref.tabulate(quad.pts)[0], gather(space, x))


def eval_scalar_grad(space, quad, x):
    _, grad = scalar_basis(space, quad)
    return np.einsum("cqmi,cm->cqi", grad, gather(space, x))


def eval_rt(space, quad, x):
    u, d = rt_basis(space, quad)
    xe = gather(space, x)
    return np.einsum("cqmi,cm->cqi", u, xe), np.einsum("cqm,cm->cq", d, xe)


def source_element_vectors(space, quad, fq):
    """int f v sqrtg with f given at the quadrature points (Nc,Q)."""
    v = space.ref.tabulate(quad.pts)[0]
    return np.einsum("cq,qa->ca", quad.dV * quad.sqrtg * fq, v)


# ---- assembly --------------------------------------------------------------------------------------

def assemble_matrix(test, trial, Ke):
    """COO -> CSR, duplicates summed, structural zeros kept; rows/cols of Dirichlet DOFs dropped."""
    I = np.repeat(test.cell_dofs[:, :, None], trial.ndofs_loc, axis=2).reshape(-1)
    J = np.repeat(trial.cell_dofs[:, None, :], test.ndofs_loc, axis=1).reshape(-1)
    V = Ke.reshape(-1)
    keep = (I >= 0) & (J >= 0)
    A = sp.coo_matrix((V[keep], (I[keep], J[keep])), shape=(test.num_free, trial.num_free)).tocsr()
    A.sort_indices()
    return A


def assemble_vector(space, fe):
    ids = space.cell_dofs.reshape(-1)
    keep = ids >= 0
    out = np.zeros(space.num_free)
    np.add.at(out, ids[keep], fe.reshape(-1)[keep])
    return out


def csr_pattern(test, trial):
    """Sorted CSR pattern (rowptr, colind) of the union of all cell couplings."""
    A = assemble_matrix(test, trial, np.ones((test.cell_dofs.shape[0], test.ndofs_loc, trial.ndofs_loc)))
    return A.indptr.astype(np.int64), A.indices.astype(np.int64)


# ---- manufactured right-hand side -------------------------------------------------------------------

def surface_laplacian(f, gradf, hessf, mesh, x, cell_to_chart):
    """Closed-form Delta_s f at chart points x (Nc,Q,2) -- src/CellData/SurfaceDiffOps.jl:26-65.

    f: xyz->(..), gradf: xyz->(..,3), hessf: xyz->(..,3,3)."""
    r = mesh.radius
    out = np.empty(x.shape[:2])
    for p in range(6):
        sel = cell_to_chart == p
        xs = x[sel]
        X = geo.csphere_eval(p + 1, r, xs)
        J = geo.csphere_jac(p + 1, r, xs)                  # [i,k]
        H = geo.csphere_hess(p + 1, r, xs)                 # [l,i,k]
        g = geo.csphere_metric(r, xs)
        gi = geo.csphere_inv_metric(r, xs)
        dg = geo.csphere_metric_grad(r, xs)                # [k,i,j]
        dgi = geo.csphere_inv_metric_grad(r, xs)
        meas = np.sqrt(g[..., 0, 0] * g[..., 1, 1] - g[..., 0, 1] ** 2)
        gf = gradf(X)
        hf = hessf(X)
        grad_f = np.einsum("...k,...ik->...i", gf, J)      # chart gradient of f o phi
        # grad(sqrt(det g))_k = 0.5/sqrt(det) * cof(g)_ij dg[k,i,j]
        cof = np.empty_like(g)
        cof[..., 0, 0] = g[..., 1, 1]
        cof[..., 1, 1] = g[..., 0, 0]
        cof[..., 0, 1] = -g[..., 1, 0]
        cof[..., 1, 0] = -g[..., 0, 1]
        grad_meas = (0.5 / meas)[..., None] * np.einsum("...ij,...kij->...k", cof, dg)
        # hessian of f o phi: J hf J^T + gf . H
        gg = np.einsum("...ik,...kl,...jl->...ij", J, hf, J) + np.einsum("...k,...ijk->...ij", gf, H)
        div_gi = np.einsum("...iij->...j", dgi)            # (div ginv)_j = d_i ginv_ij
        t1 = np.einsum("...i,...ij,...j->...", grad_meas, gi, grad_f)
        t2 = meas * np.einsum("...j,...j->...", div_gi, grad_f)
        t3 = meas * np.einsum("...ij,...ij->...", gi, gg)
        out[sel] = (t1 + t2 + t3) / meas
    return out


def surface_gradient(gradf, mesh, x, cell_to_chart):
    """Contravariant components of the surface gradient, ginv . (J grad_X f) -- src/CellData/SurfaceDiffOps.jl:107-115 (_∇s_no_ad)."""
    out = np.empty(x.shape[:2] + (2,))
    for p in range(6):
        sel = cell_to_chart == p
        xs = x[sel]
        X = geo.csphere_eval(p + 1, mesh.radius, xs)
        J = geo.csphere_jac(p + 1, mesh.radius, xs)
        gi = geo.csphere_inv_metric(mesh.radius, xs)
        out[sel] = np.einsum("...ij,...jk,...k->...i", gi, J, gradf(X))
    return out


def surface_divergence(F, gradF, mesh, x, cell_to_chart):
    """divs F = 1/sqrtg div(sqrtg Jdag (F o phi)), Jdag = ginv J -- src/CellData/SurfaceDiffOps.jl:228-259 (_divs_no_ad).

    F: xyz -> (..,3); gradF: xyz -> (..,3,3) with Gridap's convention gradF[l,k] = d F_k / d X_l."""
    r = mesh.radius
    out = np.empty(x.shape[:2])
    for p in range(6):
        sel = cell_to_chart == p
        xs = x[sel]
        X = geo.csphere_eval(p + 1, r, xs)
        J = geo.csphere_jac(p + 1, r, xs)                  # [i,k]
        H = geo.csphere_hess(p + 1, r, xs)                 # [l,i,k] = d_l J[i,k]
        g = geo.csphere_metric(r, xs)
        gi = geo.csphere_inv_metric(r, xs)
        dg = geo.csphere_metric_grad(r, xs)                # [k,i,j]
        dgi = geo.csphere_inv_metric_grad(r, xs)           # [k,i,j]
        meas = np.sqrt(g[..., 0, 0] * g[..., 1, 1] - g[..., 0, 1] ** 2)
        cof = np.empty_like(g)
        cof[..., 0, 0], cof[..., 1, 1] = g[..., 1, 1], g[..., 0, 0]
        cof[..., 0, 1], cof[..., 1, 0] = -g[..., 1, 0], -g[..., 0, 1]
        grad_meas = (0.5 / meas)[..., None] * np.einsum("...ij,...kij->...k", cof, dg)
        Fv, dF = F(X), gradF(X)
        JF = np.einsum("...jk,...k->...j", J, Fv)
        tr1 = np.einsum("...iij,...j->...", dgi, JF)                         # (d_i ginv_ij) (J F)_j
        tr2 = np.einsum("...ij,...ijk,...k->...", gi, H, Fv)                 # ginv_ij (d_i J_jk) F_k
        tr3 = np.einsum("...ij,...jk,...il,...lk->...", gi, J, J, dF)        # ginv_ij J_jk d_i (F_k o phi)
        out[sel] = tr1 + tr2 + tr3 + np.einsum("...i,...ij,...j->...", grad_meas, gi, JF) / meas
    return out
