"""CPU tests of the linear Boussinesq oracle (oracle/boussinesq.py): the hexahedral Raviart-Thomas element and the properties
the semi-discrete system of test/Tutorials/LinearBoussinesq.jl:112-134 must have (the tutorial holds no numbers to pin)."""
import numpy as np
import pytest

from oracle import boussinesq as bq
from oracle import geometry as geo
from oracle.shell import CubedSphereShellMesh3D


@pytest.mark.parametrize("k", [0, 1])
def test_hex_rt_element_is_unisolvent_and_satisfies_the_divergence_theorem(k):
    R = bq.RaviartThomasHex(k)
    assert R.ndofs == 3 * (k + 2) * (k + 1) ** 2
    pts, W = R.moment_points()
    val, _ = R.tabulate(pts)
    assert np.abs(np.einsum("mnc,njc->mj", W, val) - np.eye(R.ndofs)).max() < 1e-12       # dof_m(phi_j) = delta_mj
    # int div(phi_j) = sum of the outward face fluxes = sum of the zeroth face moments
    from oracle.refel import tensor_quadrature
    q, w = tensor_quadrature(3, 2 * k + 2)
    _, div = R.tabulate(q)
    flux = np.zeros(R.ndofs)
    for f in range(6):
        flux[f * R.nface_dofs] = 1.0                                                       # moment against s^0 t^0 of face f
    assert np.abs(w @ div - flux).max() < 1e-12


@pytest.mark.parametrize("k", [0, 1])
def test_normal_flux_is_continuous_across_every_interior_face(k):
    """H(div) conformity on the shell, across panel edges too: the reference normal flux uhat.n of the two cells of a face agrees
    (with opposite outward normals) at matching face points -- the face parametrisations (spanning axes in increasing order) of
    neighbouring hexahedra coincide on this mesh, so sign flips alone make the space conforming."""
    mesh = CubedSphereShellMesh3D(2, 1.0, 0.19, n_layers=2)
    V = bq.HexRTSpace(mesh, k, no_flux=False)
    x = np.random.default_rng(1).standard_normal(V.num_free)
    s, t = np.meshgrid([0.2, 0.7], [0.1, 0.9], indexing="ij")
    s, t = s.reshape(-1), t.reshape(-1)
    seen = {}
    worst = 0.0
    for c in range(mesh.num_cells):
        coef = x[V.cell_dofs[c]] * V.signs[c]
        for f, (ax, val) in enumerate(bq.HEX_FACES):
            spn = [a for a in range(3) if a != ax]
            p = np.zeros((len(s), 3))
            p[:, ax] = val
            p[:, spn[0]], p[:, spn[1]] = s, t
            vals, _ = V.ref.tabulate(p)
            flux = (vals[:, :, ax] @ coef) * (1.0 if val == 1 else -1.0)                    # uhat . outward normal
            fid = V.cell_faces[c, f]
            if fid in seen:
                worst = max(worst, np.abs(flux + seen[fid]).max())
                assert np.abs(flux + seen[fid]).max() < 1e-11 * max(1.0, np.abs(flux).max())
            else:
                seen[fid] = flux
    assert len(seen) == V.num_faces


@pytest.mark.parametrize("k,n,L", [(0, 2, 2), (1, 2, 1), (1, 1, 2)])
def test_energy_balance_and_jacobian(k, n, L):
    mesh = CubedSphereShellMesh3D(n, 1.0, 0.19, n_layers=L)
    P = bq.BoussinesqProblem(mesh, k)
    x = np.random.default_rng(2).standard_normal(P.nu + 2 * P.np_)
    r = P.residual(x)
    u, p, b = P.split(x)
    ru, rp, rb = P.split(r)
    scale = np.abs(u @ ru) + np.abs(p @ rp) + np.abs(b @ rb)
    assert abs(u @ ru + p @ rp / P.c2 + b @ rb / P.N2) < 1e-12 * scale                     # Coriolis skew, couplings cancel
    assert np.abs(P.jacobian() @ x - r).max() < 1e-12 * np.abs(r).max()                    # res is linear
    # no-flux spaces: no DOF on the bottom and top faces
    assert (P.V.cell_dofs[mesh.cell_layer == 0][:, 4 * (k + 1) ** 2: 5 * (k + 1) ** 2] == -1).all()
    assert (P.V.cell_dofs[mesh.cell_layer == L - 1][:, 5 * (k + 1) ** 2: 6 * (k + 1) ** 2] == -1).all()
    # theta = 1/2 conserves the quadratic energy of a skew system exactly; the explicit SSPRK3 march agrees to O(dt^3)
    x0 = P.initial_state()
    dt = 0.01
    x1 = P.theta_step(x0, dt)
    assert abs(P.energy(x1) - P.energy(x0)) < 1e-11 * P.energy(x0)
    x1e = P.step(x0, dt)
    assert np.abs(x1e - x1).max() < 5e-4 * np.abs(x0).max()


def test_rt_interpolant_converges():
    errs = []
    for n in (1, 2, 4):
        mesh = CubedSphereShellMesh3D(n, 1.0, 0.19, n_layers=max(n // 2, 1))
        V = bq.HexRTSpace(mesh, 1, no_flux=False)
        q = bq.ShellRTQuadrature(mesh, 6)
        u, _ = bq.rt_basis(V, q)
        xh = bq.interpolate_rt(V, bq.u0)
        uh = np.einsum("cqmi,cm->cqi", u, xh[V.cell_dofs])
        # exact chart field meas * J^-T v
        ue = np.empty(q.x.shape)
        for p in range(6):
            sel = mesh.cell_to_chart == p
            J = geo.csphere3d_jac(p + 1, 1.0, 0.19, q.x[sel])
            ue[sel] = q.meas[sel][..., None] * np.linalg.solve(np.swapaxes(J, -1, -2), bq.u0(q.ambient()[sel])[..., None])[..., 0]
        e = uh - ue
        errs.append(np.sqrt((np.einsum("cqi,cqij,cqj->cq", e, q.g, e) / q.meas * q.dV).sum()))
    assert errs[1] < errs[0] / 3 and errs[2] < errs[1] / 3
