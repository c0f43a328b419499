"""Vector-valued H1 space with the panel-interface change of basis (src/FESpaces/GradConformingFESpaces.jl) pinned to the
reference's own test of it: test/Projection/L2ProjectionTests.jl (H1 vector projection of (-y, x, 0), slope >= p for p = 1, 2)."""
import numpy as np
import pytest

from oracle import vector_h1 as vh
from oracle.mesh import CubedSphereMesh2D


def uX(X):
    return np.stack([-X[..., 1], X[..., 0], 0 * X[..., 0]], axis=-1)


@pytest.mark.parametrize("p", [1, 2])
def test_l2_projection_converges(p):
    errs, eint = [], []
    lv = (1, 2, 3) if p == 2 else (1, 2, 3, 4)
    for lvl in lv:
        e, ei, V, U = vh.l2_projection(CubedSphereMesh2D(2**lvl), p, uX)
        errs.append(e)
        eint.append(ei)
    h = np.log10([1 / 2.0**l for l in lv])
    assert np.polyfit(h, np.log10(errs), 1)[0] >= p
    assert np.polyfit(h, np.log10(eint), 1)[0] >= p


def test_change_of_basis_blocks():
    mesh = CubedSphereMesh2D(2)
    V = vh.VectorCGSpace(mesh, 2)
    I = np.eye(2)
    changed = np.abs(V.cob - I).reshape(mesh.num_cells, -1).max(axis=1) > 1e-14
    assert changed.any() and not changed.all() or mesh.num_cells == 24          # every cell of a 2x2 panel touches a panel edge
    # interior nodes are never transformed; the blocks are invertible
    assert np.abs(V.cob[:, -1] - I).max() == 0
    assert np.abs(np.linalg.det(V.cob)).min() > 0.1
    # a tangent field interpolated DOF-wise is single valued: all cells around a node agree on the master-frame components
    U = vh.interpolate(V, uX)
    x = V._node_chart_coords()
    from oracle import geometry as geo
    for c in (0, 5, 17, 23):
        p = mesh.cell_to_chart[c]
        J = geo.csphere_jac(p + 1, 1.0, x[c])
        loc = np.einsum("aik,ak->ai", geo.pinvJ(J), uX(geo.csphere_eval(p + 1, 1.0, x[c])))
        mas = np.einsum("aij,aj->ai", np.linalg.inv(V.cob[c]), loc)
        assert np.abs(U[V.cell_dofs[c]].reshape(2, -1).T - mas).max() < 1e-13
