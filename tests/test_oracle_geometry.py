"""Pins the oracle's geometry to the reference's own known-answer tests (CPU only)."""
import numpy as np
import pytest

from oracle import geometry as geo
from oracle import forms
from oracle.mesh import CubedSphereMesh2D

HALF = np.pi / 4
# the four sample points of /root/reference/test/AtlasDiscreteModels/seq/MetricFieldsTests.jl:94-99
PTS = np.array([[0.0, 0.0], [HALF / 2, HALF / 3], [-HALF / 3, HALF / 2], [-0.6 * HALF, -0.4 * HALF]])


@pytest.mark.parametrize("panel", range(1, 7))
def test_metric_is_JJt_and_inverse(panel):
    # MetricFieldsTests.jl:19,34-51,93-110 -- analytic g == J J^T, g g^-1 == I, ATOL 1e-12, r = 1.2
    r = 1.2
    J = geo.csphere_jac(panel, r, PTS)
    g = J @ np.swapaxes(J, -1, -2)
    assert np.abs(g - geo.csphere_metric(r, PTS)).max() < 1e-12
    assert np.abs(geo.csphere_metric(r, PTS) @ geo.csphere_inv_metric(r, PTS) - np.eye(2)).max() < 1e-12
    assert np.abs(np.sqrt(np.linalg.det(g)) - geo.csphere_sqrtg(r, PTS)).max() < 1e-12


def test_metric_panel_independent():
    # MetricFieldsTests.jl:112-122
    r = 1.2
    g1 = geo.csphere_jac(1, r, PTS)
    g1 = g1 @ np.swapaxes(g1, -1, -2)
    for p in range(2, 7):
        J = geo.csphere_jac(p, r, PTS)
        assert np.abs(J @ np.swapaxes(J, -1, -2) - g1).max() < 1e-12


@pytest.mark.parametrize("panel", range(1, 7))
def test_forward_inverse_map(panel):
    # test/Fields/seq/ForwardInverseMapTests.jl:20-40,49-66
    r = 1.7
    corners = np.array([[-HALF, -HALF], [HALF, -HALF], [-HALF, HALF], [HALF, HALF], [0.1, -0.3]])
    X = geo.csphere_eval(panel, r, corners)
    assert np.abs(np.linalg.norm(X, axis=-1) - r).max() < 1e-12   # CubedSphereTests.jl:60-69
    assert np.abs(geo.csphere_inv_eval(panel, X) - corners).max() < 1e-13
    # d(inv)/dxyz . d(fwd)/dalpha = I on the tangent plane
    J = geo.csphere_jac(panel, r, corners)            # [i,k]
    Ji = geo.csphere_inv_jac(panel, X)                # [k,i]
    assert np.abs(J @ Ji - np.eye(2)).max() < 1e-12


@pytest.mark.parametrize("panel", range(1, 7))
def test_jacobian_hessian_finite_differences(panel):
    r, e = 1.3, 1e-5
    for i in range(2):
        d = np.zeros(2)
        d[i] = e
        fd = (geo.csphere_eval(panel, r, PTS + d) - geo.csphere_eval(panel, r, PTS - d)) / (2 * e)
        assert np.abs(fd - geo.csphere_jac(panel, r, PTS)[:, i, :]).max() < 1e-8
        fdJ = (geo.csphere_jac(panel, r, PTS + d) - geo.csphere_jac(panel, r, PTS - d)) / (2 * e)
        assert np.abs(fdJ - geo.csphere_hess(panel, r, PTS)[:, i, :, :]).max() < 1e-8
        fdg = (geo.csphere_metric(r, PTS + d) - geo.csphere_metric(r, PTS - d)) / (2 * e)
        assert np.abs(fdg - geo.csphere_metric_grad(r, PTS)[:, i]).max() < 1e-8
        fdi = (geo.csphere_inv_metric(r, PTS + d) - geo.csphere_inv_metric(r, PTS - d)) / (2 * e)
        assert np.abs(fdi - geo.csphere_inv_metric_grad(r, PTS)[:, i]).max() < 1e-8


def test_shell_metric():
    r, t = 1.0, 0.25
    x = np.array([[0.3, 0.1, -0.2], [0.9, -0.5, 0.6]])
    for p in range(1, 7):
        J = geo.csphere3d_jac(p, r, t, x)
        assert np.abs(J @ np.swapaxes(J, -1, -2) - geo.csphere3d_metric(r, t, x)).max() < 1e-12
    assert np.abs(geo.csphere3d_metric(r, t, x) @ geo.csphere3d_inv_metric(r, t, x) - np.eye(3)).max() < 1e-12


@pytest.mark.parametrize("radius", [1.0, 2.0])
@pytest.mark.parametrize("degree", [2, 4, 6, 8])
def test_surface_area(radius, degree):
    # test/Geometry/SurfaceArea.jl:25-38: |int sqrtg - 4 pi r^2| / (4 pi r^2) < 1e-2 for nref 1..4
    for lvl in range(1, 5):
        m = CubedSphereMesh2D(2**lvl, radius)
        q = forms.CellQuadrature(m, degree)
        area = (q.dV * q.sqrtg).sum()
        assert abs(area - 4 * np.pi * radius**2) / (4 * np.pi * radius**2) < 1e-2


def test_mesh_counts_and_chart_coords():
    # RefinementOrderingTests.jl:42-48,93-100 (6*4^l cells, parent-major) and CellMapTests.jl:16-37
    for lvl in range(0, 4):
        n = 2**lvl
        m = CubedSphereMesh2D(n)
        assert m.num_cells == 6 * 4**lvl
        assert m.num_nodes == 6 * n * n + 2 and m.num_edges == 12 * n * n
        assert (m.cell_to_chart == np.repeat(np.arange(6), n * n)).all()
        h = 2 * HALF / n
        for p in range(6):
            c = m.cell_chart_coords[p * n * n:(p + 1) * n * n]
            i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
            assert np.allclose(c[:, 0, 0], (-HALF + i * h).reshape(-1), atol=1e-14)
            assert np.allclose(c[:, 0, 1], (-HALF + j * h).reshape(-1), atol=1e-14)
            assert np.allclose(c[:, 3] - c[:, 0], h, atol=1e-14)
        assert not m.cell_edge_flip.any()   # DarcyCubedSphereTests.jl:25-32: all edge permutations are identity
        # closed manifold: every edge has exactly two cells, every node is shared by 3 or 4 cells
        assert (np.bincount(m.cell_edges.reshape(-1)) == 2).all()
        cnt = np.bincount(m.cell_nodes.reshape(-1))
        if n > 1:
            assert (cnt == 3).sum() == 8 and (cnt == 4).sum() == m.num_nodes - 8


def test_ambient_continuity_across_panels():
    # nodes shared by cells of different panels map to the same point of the sphere
    m = CubedSphereMesh2D(4, 1.5)
    X = np.full((m.num_nodes, 3), np.nan)
    for c in range(m.num_cells):
        P = geo.csphere_eval(m.cell_to_chart[c] + 1, m.radius, m.cell_chart_coords[c])
        for k in range(4):
            nd = m.cell_nodes[c, k]
            if np.isnan(X[nd, 0]):
                X[nd] = P[k]
            else:
                assert np.abs(X[nd] - P[k]).max() < 1e-13
