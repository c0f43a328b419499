"""Pins the 3-D shell oracle (oracle/shell.py, oracle/geometry.py csphere3d_*) against what the reference's tests assert:
shell volume 4/3 pi ((r+t)^3 - r^3) (test/Distributed/DistributedSurfaceAreaTests.jl:120-134), closed-shell entity counts,
and structural properties of the Laplace-Beltrami matrix (symmetric, constants in the kernel, positive semi-definite)."""
import numpy as np

from oracle import forms, shell


def test_entity_counts_of_the_closed_shell():
    for n, L in ((1, 1), (2, 3), (3, 2)):
        m = shell.CubedSphereShellMesh3D(n, 1.0, 0.19, n_layers=L)
        nv2, ne2, nc2 = 6 * n * n + 2, 12 * n * n, 6 * n * n            # the surface mesh
        assert m.num_cells == nc2 * L and m.num_nodes == nv2 * (L + 1)
        V = shell.HexCGSpace(m, 2)
        assert V.num_edges == ne2 * (L + 1) + nv2 * L and V.num_faces == nc2 * (L + 1) + ne2 * L
        assert V.num_free == m.num_nodes + V.num_edges + V.num_faces + m.num_cells


def test_shell_volume_converges():
    r, t = 2.0, 0.25
    exact = 4.0 / 3.0 * np.pi * ((r + t) ** 3 - r ** 3)
    errs = [abs(shell.shell_volume(shell.CubedSphereShellMesh3D(n, r, t), 4) - exact) / exact for n in (1, 2, 4)]
    assert errs[0] < 1e-2 and errs[2] < errs[1] < errs[0] and errs[2] < 1e-6


def test_laplace_beltrami_matrix_properties():
    m = shell.CubedSphereShellMesh3D(2, 1.0, 0.19)
    V = shell.HexCGSpace(m, 2)
    Ke = shell.lb_element_matrices(V, shell.ShellQuadrature(m, 8))
    assert np.abs(Ke - Ke.transpose(0, 2, 1)).max() < 1e-14 * np.abs(Ke).max()
    A = forms.assemble_matrix(V, V, Ke)
    assert np.abs(A @ np.ones(V.num_free)).max() < 1e-12 * np.abs(A.data).max()
    w = np.linalg.eigvalsh(A.toarray())
    assert w[0] > -1e-10 * w[-1] and w[1] > 1e-6 * w[-1]             # one zero eigenvalue (constants), the rest positive


def test_shell_mass_sums_to_the_volume_and_helmholtz_splits():
    mesh = shell.CubedSphereShellMesh3D(2, 1.3, 0.25, n_layers=2)
    V = shell.HexCGSpace(mesh, 1)
    q = shell.ShellQuadrature(mesh, 8)
    M = shell.mass_element_matrices(V, q)
    vol = 4.0 / 3.0 * np.pi * ((1.3 + 0.25) ** 3 - 1.3**3)
    assert abs(M.sum() - vol) < 1e-6 * vol
    H = shell.helmholtz_element_matrices(V, q)
    assert np.abs(H.sum(axis=(1, 2)) - M.sum(axis=(1, 2))).max() < 1e-12     # the stiffness part annihilates constants


def test_shell_laplace_beltrami_driver_converges():
    """3-D run of test/Laplacian/LaplaceBeltrami.jl (test/Laplacian/mpi/LaplaceBeltramiTests.jl:23-28: two refinement levels,
    p = 2, u = xyz): the shell is a flat 3-D domain in curvilinear coordinates, so Delta u = 0 for u = xyz and the data enter
    through the boundary term alone; slope >= p - 0.1 (src/ConvergenceTools/Tools.jl:4-13)."""
    f = lambda X: X[..., 0] * X[..., 1] * X[..., 2]          # noqa: E731
    lap = lambda X: np.zeros(X.shape[:-1])                   # noqa: E731
    errs = []
    for n in (1, 2):
        mesh = shell.CubedSphereShellMesh3D(n, 1.0, 0.19)
        errs.append(shell.solve_lb(mesh, 2, f, lap, degree=8))
    slope = np.log(errs[0] / errs[1]) / np.log(2.0)
    assert slope >= 2 - 0.1, (errs, slope)
