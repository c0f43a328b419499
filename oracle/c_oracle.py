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
