"""Oracle (test / baseline infrastructure): ctypes wrapper of the C restatement in oracle/c/gg_oracle.c."""
import ctypes as C
import os
import subprocess

import numpy as np

from .refel import LagrangeQ, tensor_quadrature

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "c", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-C", os.path.join(_HERE, "c")], check=True)
        _lib = C.CDLL(_SO)
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def max_threads():
    return int(lib().oracle_max_threads())


def lb_element_matrices(chart_coords, radius, p, degree, nthreads=0):
    cc = np.ascontiguousarray(chart_coords, dtype=np.float64)
    nc = cc.shape[0]
    pts, w = tensor_quadrature(2, degree)
    ref = LagrangeQ(2, p)
    _, g = ref.tabulate(pts)
    g = np.ascontiguousarray(g)
    m = ref.ndofs
    Ke = np.empty((nc, m, m))
    lib().oracle_lb_element_matrices(C.c_int64(nc), _p(cc), C.c_double(radius), C.c_int(len(w)), _p(np.ascontiguousarray(pts)),
                                     _p(np.ascontiguousarray(w)), C.c_int(m), _p(g), _p(Ke), C.c_int(nthreads))
    return Ke


def lb_element_vectors(chart_coords, radius, p, degree, ue, nthreads=0):
    cc = np.ascontiguousarray(chart_coords, dtype=np.float64)
    nc = cc.shape[0]
    pts, w = tensor_quadrature(2, degree)
    ref = LagrangeQ(2, p)
    _, g = ref.tabulate(pts)
    g = np.ascontiguousarray(g)
    m = ref.ndofs
    ue = np.ascontiguousarray(ue, dtype=np.float64)
    fe = np.empty((nc, m))
    lib().oracle_lb_element_vectors(C.c_int64(nc), _p(cc), C.c_double(radius), C.c_int(len(w)), _p(np.ascontiguousarray(pts)),
                                    _p(np.ascontiguousarray(w)), C.c_int(m), _p(g), _p(ue), _p(fe), C.c_int(nthreads))
    return fe


def scatter_csr(cell_dofs, Ke, rowptr, colind, nthreads=1):
    ids = np.ascontiguousarray(cell_dofs, dtype=np.int64)
    rp = np.ascontiguousarray(rowptr, dtype=np.int64)
    ci = np.ascontiguousarray(colind, dtype=np.int64)
    vals = np.zeros(len(ci))
    Ke = np.ascontiguousarray(Ke)
    lib().oracle_scatter_csr(C.c_int64(ids.shape[0]), C.c_int(ids.shape[1]), _p(ids, C.c_int64), _p(Ke), _p(rp, C.c_int64),
                             _p(ci, C.c_int64), _p(vals), C.c_int(nthreads))
    return vals


def scatter_vector(cell_dofs, fe, n, nthreads=1):
    ids = np.ascontiguousarray(cell_dofs, dtype=np.int64)
    out = np.zeros(n)
    fe = np.ascontiguousarray(fe)
    lib().oracle_scatter_vector(C.c_int64(ids.shape[0]), C.c_int(ids.shape[1]), _p(ids, C.c_int64), _p(fe), _p(out), C.c_int(nthreads))
    return out


# ---- explicit-RK half of the path (oracle/c/gg_oracle_rk.c): linear wave and rotating shallow water -------------------------------

def _i8(a):
    return a.ctypes.data_as(C.POINTER(C.c_int8))


class _Tabs:
    """Reference tabulations at the quadrature points of Measure(., degree), in the layout the C functions take."""

    def __init__(self, mesh, p_fe, degree, with_cg=False):
        from .spaces import CGSpace, DGSpace, RTSpace
        self.mesh = mesh
        self.V, self.Q = RTSpace(mesh, p_fe), DGSpace(mesh, p_fe)
        self.pts, self.w = tensor_quadrature(2, degree)
        self.pts, self.w = np.ascontiguousarray(self.pts), np.ascontiguousarray(self.w)
        rv, rd = self.V.ref.tabulate(self.pts)
        self.rt_val, self.rt_div = np.ascontiguousarray(rv), np.ascontiguousarray(rd)
        self.dg_val = np.ascontiguousarray(self.Q.ref.tabulate(self.pts)[0])
        self.signs = np.ascontiguousarray(self.V.signs, dtype=np.int8)
        self.cc = np.ascontiguousarray(mesh.cell_chart_coords, dtype=np.float64)
        self.nc, self.nq = mesh.num_cells, len(self.w)
        self.mu, self.mp = self.V.ref.ndofs, self.Q.ref.ndofs
        if with_cg:
            self.R = CGSpace(mesh, p_fe + 1)
            cv, cg = self.R.ref.tabulate(self.pts)
            self.cg_val, self.cg_grad = np.ascontiguousarray(cv), np.ascontiguousarray(cg)
            self.mr = self.R.ref.ndofs


def rt_mass_matrices(T, nthreads=0):
    Ke = np.empty((T.nc, T.mu, T.mu))
    lib().oracle_rt_mass_matrices(C.c_int64(T.nc), _p(T.cc), C.c_double(T.mesh.radius), C.c_int(T.nq), _p(T.pts), _p(T.w),
                                  C.c_int(T.mu), _p(T.rt_val), _i8(T.signs), _p(Ke), C.c_int(nthreads))
    return Ke


def scalar_mass_matrices(T, val, coef=None, nthreads=0):
    m = val.shape[1]
    Ke = np.empty((T.nc, m, m))
    lib().oracle_scalar_mass_matrices(C.c_int64(T.nc), _p(T.cc), C.c_double(T.mesh.radius), C.c_int(T.nq), _p(T.pts), _p(T.w),
                                      C.c_int(m), _p(val), None if coef is None else _p(np.ascontiguousarray(coef)), _p(Ke),
                                      C.c_int(nthreads))
    return Ke


def csr_pcg(A, b, x0=None, rtol=1e-12, max_iter=2000, nthreads=0):
    rp = np.ascontiguousarray(A.indptr, dtype=np.int64)
    ci = np.ascontiguousarray(A.indices, dtype=np.int64)
    x = np.zeros(len(b)) if x0 is None else np.array(x0, dtype=np.float64)
    lib().oracle_csr_pcg.restype = C.c_int
    it = lib().oracle_csr_pcg(C.c_int64(len(b)), _p(rp, C.c_int64), _p(ci, C.c_int64), _p(np.ascontiguousarray(A.data)),
                              _p(np.ascontiguousarray(b, dtype=np.float64)), _p(x), C.c_double(rtol), C.c_int(max_iter), C.c_int(nthreads))
    if it < 0:
        raise RuntimeError(f"oracle PCG did not converge in {-it} iterations")
    return x, it


def block_solve(Me, be, nthreads=0):
    Me = np.array(Me, dtype=np.float64, order="C")
    be = np.ascontiguousarray(be, dtype=np.float64)
    xe = np.empty_like(be)
    lib().oracle_block_solve(C.c_int64(Me.shape[0]), C.c_int(Me.shape[1]), _p(Me), _p(be), _p(xe), C.c_int(nthreads))
    return xe


class CWaveProblem:
    """Linear wave (TransientWaveEquation.jl:28-77) on the C restatement: residual re-integrated cell by cell at every stage,
    RT mass solve by Jacobi-PCG, DG mass solve block by block."""

    def __init__(self, mesh, p_fe, degree=None, nthreads=0):
        from . import forms
        self.nt = nthreads
        self.T = T = _Tabs(mesh, p_fe, 5 * (p_fe + 1) if degree is None else degree)
        self.nu, self.np_ = T.V.num_free, T.Q.num_free
        rp, ci = forms.csr_pattern(T.V, T.V)
        import scipy.sparse as sp
        self.Mu = sp.csr_matrix((scatter_csr(T.V.cell_dofs, rt_mass_matrices(T, nthreads), rp, ci, nthreads), ci, rp),
                                shape=(self.nu, self.nu))
        self.Mp_e = scalar_mass_matrices(T, T.dg_val, None, nthreads)
        self.iters = 0

    def residual(self, x):
        T = self.T
        ue = np.ascontiguousarray(x[: self.nu][T.V.cell_dofs])
        pe = np.ascontiguousarray(x[self.nu:][T.Q.cell_dofs])
        ru, rp = np.empty_like(ue), np.empty_like(pe)
        lib().oracle_wave_residual_vectors(C.c_int64(T.nc), _p(T.cc), C.c_int(T.nq), _p(T.pts), _p(T.w), C.c_int(T.mu), _p(T.rt_div),
                                           _i8(T.signs), C.c_int(T.mp), _p(T.dg_val), _p(ue), _p(pe), _p(ru), _p(rp), C.c_int(self.nt))
        return scatter_vector(T.V.cell_dofs, ru, self.nu, self.nt), rp

    def slope(self, x):
        r_u, r_pe = self.residual(x)
        ku, it = csr_pcg(self.Mu, -r_u, rtol=1e-12, nthreads=self.nt)
        self.iters += it
        kp = np.zeros(self.np_)
        kp[self.T.Q.cell_dofs] = block_solve(self.Mp_e, -r_pe, self.nt)
        return np.concatenate([ku, kp])

    def step(self, x0, dt, tableau="EXRK_SSP_3_3"):
        from .rk import TABLEAUS
        A, b, c = TABLEAUS[tableau]
        ks = []
        for i in range(len(c)):
            xi = x0.copy()
            for j in range(i):
                xi += dt * A[i, j] * ks[j]
            ks.append(self.slope(xi))
        xF = x0.copy()
        for i in range(len(c)):
            xF += dt * b[i] * ks[i]
        return xF


class CShallowWaterProblem:
    """Rotating shallow water DAE (TransientShallowWater.jl:113-154, RungeKuttaEX.jl:48-134) on the C restatement: as the
    reference, the potential-vorticity block of jac_y is re-assembled at every diagnostic solve; solves by Jacobi-PCG."""

    def __init__(self, mesh, p_fe, coriolis, gravity, topography=None, degree=None, q0_mode="alias", nthreads=0):
        from . import forms
        import scipy.sparse as sp
        self.nt, self.gravity, self.q0_mode = nthreads, gravity, q0_mode
        self.T = T = _Tabs(mesh, p_fe, 4 * (p_fe + 1) if degree is None else degree, with_cg=True)
        self.nu, self.np_, self.nq = T.V.num_free, T.Q.num_free, T.R.num_free
        quad = forms.CellQuadrature(mesh, 4 * (p_fe + 1) if degree is None else degree)
        X = quad.ambient()
        self.fq = np.ascontiguousarray(coriolis(X))
        self.bq = np.ascontiguousarray(topography(X)) if topography is not None else None
        rp, ci = forms.csr_pattern(T.V, T.V)
        self.Mu = sp.csr_matrix((scatter_csr(T.V.cell_dofs, rt_mass_matrices(T, nthreads), rp, ci, nthreads), ci, rp),
                                shape=(self.nu, self.nu))
        self.pat_q = forms.csr_pattern(T.R, T.R)
        self.Mp_e = scalar_mass_matrices(T, T.dg_val, None, nthreads)
        self.iters = self.solves = 0

    def _gather(self, x):
        T = self.T
        return np.ascontiguousarray(x[: self.nu][T.V.cell_dofs]), np.ascontiguousarray(x[self.nu:][T.Q.cell_dofs])

    def diagnostics(self, x):
        import scipy.sparse as sp
        T = self.T
        ue, pe = self._gather(x)
        rq, rF, rP = np.empty((T.nc, T.mr)), np.empty((T.nc, T.mu)), np.empty((T.nc, T.mp))
        pq = np.empty((T.nc, T.nq))
        lib().oracle_swe_diag_vectors(C.c_int64(T.nc), _p(T.cc), C.c_double(T.mesh.radius), C.c_int(T.nq), _p(T.pts), _p(T.w),
                                      C.c_int(T.mu), _p(T.rt_val), _i8(T.signs), C.c_int(T.mp), _p(T.dg_val), C.c_int(T.mr),
                                      _p(T.cg_val), _p(T.cg_grad), _p(self.fq), None if self.bq is None else _p(self.bq),
                                      C.c_double(self.gravity), _p(ue), _p(pe), _p(rq), _p(rF), _p(rP), _p(pq), C.c_int(self.nt))
        rp, ci = self.pat_q
        Mq = sp.csr_matrix((scatter_csr(T.R.cell_dofs, scalar_mass_matrices(T, T.cg_val, pq, self.nt), rp, ci, self.nt), ci, rp),
                           shape=(self.nq, self.nq))
        qv, it1 = csr_pcg(Mq, scatter_vector(T.R.cell_dofs, rq, self.nq, self.nt), rtol=1e-12, nthreads=self.nt)
        Fv, it2 = csr_pcg(self.Mu, scatter_vector(T.V.cell_dofs, rF, self.nu, self.nt), rtol=1e-12, nthreads=self.nt)
        Pv = np.zeros(self.np_)
        Pv[T.Q.cell_dofs] = block_solve(self.Mp_e, rP, self.nt)
        self.iters += it1 + it2
        self.solves += 2
        return np.concatenate([qv, Fv, Pv])

    def residual_x(self, x, y, y0):
        T = self.T
        ue, _ = self._gather(x)
        qv, Fv, Pv = y[: self.nq], y[self.nq: self.nq + self.nu], y[self.nq + self.nu:]
        q0v = y0[: self.nq]
        Fe = np.ascontiguousarray(Fv[T.V.cell_dofs])
        Pe = np.ascontiguousarray(Pv[T.Q.cell_dofs])
        qe, q0e = np.ascontiguousarray(qv[T.R.cell_dofs]), np.ascontiguousarray(q0v[T.R.cell_dofs])
        ru, rp = np.empty_like(ue), np.empty_like(Pe)
        lib().oracle_swe_residual_vectors(C.c_int64(T.nc), _p(T.cc), C.c_double(T.mesh.radius), C.c_int(T.nq), _p(T.pts), _p(T.w),
                                          C.c_int(T.mu), _p(T.rt_val), _p(T.rt_div), _i8(T.signs), C.c_int(T.mp), _p(T.dg_val),
                                          C.c_int(T.mr), _p(T.cg_val), _p(T.cg_grad), C.c_double(self.tau), C.c_double(self.dt),
                                          _p(ue), _p(Fe), _p(Pe), _p(qe), _p(q0e), _p(ru), _p(rp), C.c_int(self.nt))
        return scatter_vector(T.V.cell_dofs, ru, self.nu, self.nt), rp

    def slope(self, x, y, y0):
        r_u, r_pe = self.residual_x(x, y, y0)
        ku, it = csr_pcg(self.Mu, -r_u, rtol=1e-12, nthreads=self.nt)
        self.iters += it
        self.solves += 1
        kp = np.zeros(self.np_)
        kp[self.T.Q.cell_dofs] = block_solve(self.Mp_e, -r_pe, self.nt)
        return np.concatenate([ku, kp])

    def step(self, x0, dt, tableau="EXRK_SSP_3_3"):
        from .rk import TABLEAUS
        A, b, c = TABLEAUS[tableau]
        self.dt, self.tau = dt, dt / 2
        y_start = self.diagnostics(x0)
        ks = []
        for i in range(len(c)):
            xi = x0.copy()
            for j in range(i):
                xi += dt * A[i, j] * ks[j]
            y = self.diagnostics(xi)
            ks.append(self.slope(xi, y, y if self.q0_mode == "alias" else y_start))
        xF = x0.copy()
        for i in range(len(c)):
            xF += dt * b[i] * ks[i]
        return xF, self.diagnostics(xF)
