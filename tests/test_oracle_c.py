"""The C restatement (CPU baseline) agrees with the numpy oracle (CPU only)."""
import numpy as np
import pytest

from oracle import c_oracle, forms
from oracle.mesh import CubedSphereMesh2D
from oracle.spaces import CGSpace


@pytest.mark.parametrize("n,p,degree", [(2, 1, 12), (4, 2, 18), (3, 2, 8)])
def test_c_oracle_matches_numpy_oracle(n, p, degree):
    mesh = CubedSphereMesh2D(n, 1.4)
    V = CGSpace(mesh, p, zeromean=True)
    q = forms.CellQuadrature(mesh, degree)
    Ko = forms.lb_element_matrices(V, q)
    Kc = c_oracle.lb_element_matrices(mesh.cell_chart_coords, 1.4, p, degree)
    assert np.abs(Kc - Ko).max() < 1e-13 * np.abs(Ko).max()
    A = forms.assemble_matrix(V, V, Ko)
    vals = c_oracle.scatter_csr(V.cell_dofs, Kc, A.indptr, A.indices)
    assert np.abs(vals - A.data).max() < 1e-13 * np.abs(A.data).max()
    u = np.random.default_rng(0).standard_normal(V.num_free)
    ue = forms.gather(V, u, 0.25)
    fe = c_oracle.lb_element_vectors(mesh.cell_chart_coords, 1.4, p, degree, ue)
    feo = np.einsum("cab,cb->ca", Ko, ue)
    assert np.abs(fe - feo).max() < 1e-13 * np.abs(feo).max()
    r = c_oracle.scatter_vector(V.cell_dofs, fe, V.num_free)
    assert np.abs(r - forms.assemble_vector(V, feo)).max() < 1e-13 * np.abs(r).max()
    # threaded == serial (static cell ranges, no reduction across threads)
    K1 = c_oracle.lb_element_matrices(mesh.cell_chart_coords, 1.4, p, degree, nthreads=1)
    assert (K1 == Kc).all()
