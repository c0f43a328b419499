"""Worker of tests/test_gpu_dist.py (launched with torch.distributed.run, one rank per GPU): the partitioned RK steppers --
halo exchange and dot products through peer memory inside the solver kernels -- against the single-GPU run of the same
problem on the same mesh."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

OM, G = 1.0, 9.8 * (1 / 7.29e-5) ** 2 / 6.37e6
H0, U0 = 2.94e4 / 9.8 / 6.37e6, 2 * np.pi / (12 * 24 * 3600 * 7.29e-5)


def coriolis(X):
    return 2.0 * OM * X[..., 2] / np.linalg.norm(X, axis=-1)


def depth(X):
    s = X[..., 2] / np.linalg.norm(X, axis=-1)
    return H0 - (OM * U0 + 0.5 * U0 * U0) * s * s / G + 1e-5 * np.cos(3 * X[..., 0]) * X[..., 1]


def velocity(X):
    n = X / np.linalg.norm(X, axis=-1, keepdims=True)
    return U0 * np.stack([-n[..., 1], n[..., 0], 0.3 * n[..., 0] * n[..., 1]], -1)


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def run_checks(gg, ctx, rank, world, n, p):
    """Partitioned (one rank per GPU, or per process on a shared GPU) against single-GPU runs of the same problems; asserts the
    tolerances and returns the measured differences.  Collective: every rank calls it.  Also embedded in `bench.py --gpus N`
    (`dist_parity`), so that the multi-GPU parity is on the driver's record of the scaling run."""
    cpu = dist.get_backend() != "nccl"
    dev = torch.device("cpu") if cpu else torch.device("cuda", ctx.device)
    mesh = gg.CubedSphereMesh(1.0)
    plan = gg.PartitionPlan(n, world, rank, partition=os.environ.get("GG_TEST_PARTITION", "linear"))
    comm = gg.PeerWindow(ctx, rank, world, gg.window_doubles(plan, p))
    comm.connect()
    local_model = gg.partitioned_model(mesh, plan, comm, ctx=ctx)
    full_model = gg.AtlasDiscreteModel(mesh, n=n, ctx=ctx)
    assert local_model.num_cells == plan.num_local_cells

    def spaces(model):
        V = gg.TestFESpace(model, gg.ReferenceFE(gg.raviart_thomas, float, p), conformity="HDiv")
        Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="L2")
        return V, Q

    V, Q = spaces(local_model)
    Vf, Qf = spaces(full_model)
    lv, ov = gg.global_dof_ids(V)
    lq, oq = gg.global_dof_ids(Q)
    lv, lq = lv - 1, lq - 1

    # -- consistent!: ghosts unknown -> owner values
    for S_, l2g, own in ((V, lv, ov), (Q, lq, oq)) if 'cons' not in os.environ.get('GG_DEBUG_SKIP', '') else ():
        g = np.sin(0.1 * np.arange(l2g.max() + 1)) + 2.0
        v = gg.DeviceVector.from_numpy(ctx, np.where(own == rank, g[l2g], -777.0))
        gg.consistent(v, S_)
        assert (v.get() == g[l2g]).all(), "consistent! did not reproduce the owner values"

    # -- interpolation on the sub-model = restriction of the global interpolant; norms add up over the ranks
    u0, h0 = gg.interpolate(velocity, V).get(), gg.interpolate(depth, Q).get()
    u0f, h0f = gg.interpolate(velocity, Vf).get(), gg.interpolate(depth, Qf).get()
    assert rel(u0, u0f[lv]) < 1e-13 and rel(h0, h0f[lq]) < 1e-13
    nsq = torch.tensor([gg.norm(u0, V, 8) ** 2, gg.norm(h0, Q, 8) ** 2], device=dev, dtype=torch.float64)
    dist.all_reduce(nsq)
    assert abs(nsq[0].item() - gg.norm(u0f, Vf, 8) ** 2) < 1e-12 * nsq[0].item()
    assert abs(nsq[1].item() - gg.norm(h0f, Qf, 8) ** 2) < 1e-12 * nsq[1].item()

    x0, x0f = np.concatenate([u0, h0]), np.concatenate([u0f, h0f])
    nsteps = 3
    # -- linear wave
    skip = os.environ.get('GG_DEBUG_SKIP', '')
    dt = gg.dx(full_model) * 0.1
    rk = gg.RungeKutta(gg.CGSolver(rtol=1e-13), None, dt, "EXRK_SSP_3_3")
    st, stf = gg.RKStepper(rk, gg.WaveOperator(local_model, p)), gg.RKStepper(rk, gg.WaveOperator(full_model, p))
    st.set_state(x0); stf.set_state(x0f)
    if 'wave' not in skip:
        st.step(nsteps); stf.step(nsteps)
    x, xf = st.get_state(), stf.get_state()
    e_w = max(rel(x[: st.nu], xf[: stf.nu][lv]), rel(x[st.nu:], xf[stf.nu:][lq]))
    assert e_w < 1e-10, f"wave: partitioned and single-GPU states differ by {e_w}"
    assert comm.error_epoch() == 0
    del st, stf

    # -- rotating shallow water (diagnostic + slope solves, all of them partitioned)
    dt = gg.dx(full_model) * 0.1 / (max(p, 1) * np.sqrt(G * H0))
    rk = gg.RungeKutta(gg.CGSolver(rtol=1e-13), None, dt, "EXRK_SSP_3_3")
    st = gg.RKStepper(rk, gg.ShallowWaterOperator(local_model, p, coriolis, gravity=G))
    stf = gg.RKStepper(rk, gg.ShallowWaterOperator(full_model, p, coriolis, gravity=G))
    st.set_state(x0); stf.set_state(x0f)
    st.step(nsteps); stf.step(nsteps)
    x, xf = st.get_state(), stf.get_state()
    e_s = max(rel(x[: st.nu], xf[: stf.nu][lv]), rel(x[st.nu:], xf[stf.nu:][lq]))
    assert e_s < 1e-9, f"shallow water: partitioned and single-GPU states differ by {e_s}"
    R = gg.TestFESpace(local_model, gg.ReferenceFE(gg.lagrangian, float, p + 1), conformity="H1")
    lr, _ = gg.global_dof_ids(R)
    y, yf = st.get_diagnostics(), stf.get_diagnostics()
    e_q = rel(y[: st.nq], yf[: stf.nq][lr - 1])
    assert e_q < 1e-8, f"shallow water: potential vorticity differs by {e_q}"
    assert comm.error_epoch() == 0
    its, sol = st.stats()
    itf, solf = stf.stats()
    assert sol == solf
    del st, stf

    # -- thermal shallow water: shallow water + buoyancy skeleton terms; the B block of the stage states goes through the halo
    if p >= 1:
        def wb(X):
            return G * (1.0 + 0.05 * H0 ** 2 / depth(X) ** 2) * depth(X)

        dt = gg.dx(full_model) * 0.05 / np.sqrt(G * H0)
        rk = gg.RungeKutta(gg.CGSolver(rtol=1e-13), None, dt, "EXRK_SSP_3_3")
        stt = gg.RKStepper(rk, gg.ThermalShallowWaterOperator(local_model, p, coriolis))
        sttf = gg.RKStepper(rk, gg.ThermalShallowWaterOperator(full_model, p, coriolis))
        B0, B0f = gg.interpolate(wb, Q).get(), gg.interpolate(wb, Qf).get()
        stt.set_state(np.concatenate([x0, B0])); sttf.set_state(np.concatenate([x0f, B0f]))
        stt.step(nsteps); sttf.step(nsteps)
        xt, xtf = stt.get_state(), sttf.get_state()
        nq_ = Q.num_free_dofs
        e_t = max(rel(xt[: stt.nu], xtf[: sttf.nu][lv]), rel(xt[stt.nu: stt.nu + nq_], xtf[sttf.nu: sttf.nu + Qf.num_free_dofs][lq]),
                  rel(xt[stt.nu + nq_:], xtf[sttf.nu + Qf.num_free_dofs:][lq]))
        assert e_t < 1e-9, f"thermal shallow water: partitioned and single-GPU states differ by {e_t}"
        assert comm.error_epoch() == 0
        del stt, sttf

    # -- vector-valued H1 space with the panel-interface change of basis on the partitioned model
    #    (src/Distributed/GradConformingFESpaces.jl:7-35): the blocks of every local cell -- ghost cells included, whose master cell
    #    may lie beyond the ghost layer -- equal those of the serial space; interpolation, element matrices, norms and consistent!
    pv = max(p, 1)
    if pv <= 2:
        W_ = gg.TestFESpace(local_model, gg.ReferenceFE(gg.lagrangian, gg.VectorValue, pv), conformity="H1")
        Wf = gg.TestFESpace(full_model, gg.ReferenceFE(gg.lagrangian, gg.VectorValue, pv), conformity="H1")
        lw, ow = gg.global_dof_ids(W_)
        lw = lw - 1
        cells = plan.local_cells()
        Cl, Cf = gg.change_of_basis(W_), gg.change_of_basis(Wf)
        e_cob = float(np.abs(Cl - Cf[cells]).max())
        assert e_cob < 1e-14, f"vector H1: change-of-basis blocks of the sub-model differ by {e_cob}"
        assert np.abs(Cf - np.eye(2)).max() > 0.1                       # the blocks are not trivially the identity
        w0, w0f = gg.interpolate(velocity, W_).get(), gg.interpolate(velocity, Wf).get()
        assert rel(w0, w0f[lw]) < 1e-13
        dl, df = gg.Measure(gg.Triangulation(local_model), 4 * pv), gg.Measure(gg.Triangulation(full_model), 4 * pv)
        Kl, Kf = gg.element_matrices(gg.VectorMass(dl), W_, W_), gg.element_matrices(gg.VectorMass(df), Wf, Wf)
        assert rel(Kl, Kf[cells]) < 1e-13
        nv = torch.tensor([gg.norm(w0, W_, 4 * pv) ** 2], device=dev, dtype=torch.float64)
        dist.all_reduce(nv)
        assert abs(nv[0].item() - gg.norm(w0f, Wf, 4 * pv) ** 2) < 1e-12 * nv[0].item()
        gw = np.sin(0.2 * np.arange(lw.max() + 1)) - 0.5
        vw = gg.DeviceVector.from_numpy(ctx, np.where(ow == rank, gw[lw], 123.0))
        gg.consistent(vw, W_)
        assert (vw.get() == gw[lw]).all(), "consistent! on the vector H1 space did not reproduce the owner values"
    else:
        e_cob = None

    # -- DG upwind advection: ghost cells of every stage state filled through peer memory
    def wind(X):
        return np.stack([-X[..., 1], X[..., 0], 0.2 * X[..., 0]], -1)

    def blob(X):
        return np.exp(-(X[..., 1] ** 2 + X[..., 2] ** 2))

    dt = gg.dx(full_model) * 0.1
    rk = gg.RungeKutta(None, None, dt, "EXRK_SSP_3_3")
    sa = gg.RKStepper(rk, gg.AdvectionOperator(local_model, p, wind, degree=4 * max(p, 1)))
    saf = gg.RKStepper(rk, gg.AdvectionOperator(full_model, p, wind, degree=4 * max(p, 1)))
    c0, c0f = gg.interpolate(blob, Q).get(), gg.interpolate(blob, Qf).get()
    sa.set_state(c0); saf.set_state(c0f)
    sa.step(4); saf.step(4)
    own = oq == rank
    e_a = rel(sa.get_state(), saf.get_state()[lq])
    assert e_a < 1e-12, f"advection: partitioned and single-GPU states differ by {e_a}"
    assert comm.error_epoch() == 0
    del sa, saf
    # back-to-back exchanges with one rank delayed: the exchange slot is double-buffered by round parity, so a fast rank
    # cannot overwrite values a slow peer has not unpacked yet
    g = np.cos(0.3 * np.arange(lv.max() + 1)) + 3.0
    for rep_ in range(6):
        v = gg.DeviceVector.from_numpy(ctx, np.where(ov == rank, (rep_ + 1) * g[lv], -1.0))
        if rank == rep_ % world:
            import time
            time.sleep(0.02)
        gg.consistent(v, V)
        assert (v.get() == (rep_ + 1) * g[lv]).all(), "consistent! under skew returned stale ghost values"
    assert comm.error_epoch() == 0
    dist.barrier()
    return {"world": world, "n": n, "p": p, "wave": e_w, "swe": e_s, "pv": e_q, "advection": e_a,
            "thermal_swe": e_t if p >= 1 else None, "pcg_iters": int(its), "pcg_iters_single_gpu": int(itf),
            "pcg_iters_equal": bool(its == itf), "solves": int(sol), "consistent_exact": True, "vector_h1_cob": e_cob}


def run_shell_checks(gg, ctx, rank, world, n=4, order=2, degree=8):
    """The distributed 3-D shell (BASELINE configs[4]; test/Laplacian/mpi/LaplaceBeltramiTests.jl:30-34): p8est cell order, equal-count
    chunks of the space-filling curve + ghost layer, hexahedral CG space partitioned through the plan; halo exchange of its DOF
    vectors and the distributed PCG (all-reduces and halo through peer memory) on the assembled operator (grad v . ginv grad u +
    u v) sqrtg, against the single-GPU run."""
    import math
    L = gg._lib
    r, t = 1.0, 0.19
    shell = gg.CubedSphereWithThicknessMesh(r, t)
    plan = gg.PartitionPlan(n, world, rank, radius=r, thickness=t, ordering=2)
    comm = gg.PeerWindow(ctx, rank, world, plan.space(L.SPACE_CG, order)["slot_capacity"])
    comm.connect()
    local = gg.partitioned_model(shell, plan, comm, ctx=ctx)
    full = gg.AtlasDiscreteModel(shell, n=n, ordering=2, ctx=ctx)
    cells = plan.local_cells()
    assert local.num_cells == len(cells) and (local.get_cell_node_ids() == full.get_cell_node_ids()[cells]).all()
    a, b = gg.partition_range(full.num_cells, world, rank)
    assert (plan.first_owned, plan.last_owned) == (a, b) and b - a in (full.num_cells // world, full.num_cells // world + 1)
    V = gg.TestFESpace(local, gg.ReferenceFE(gg.lagrangian, float, order), conformity="H1")
    Vf = gg.TestFESpace(full, gg.ReferenceFE(gg.lagrangian, float, order), conformity="H1")
    l2g, own = gg.global_dof_ids(V)
    l2g = l2g - 1
    assert (V.get_cell_dof_ids() - 1 == l2g.searchsorted(Vf.get_cell_dof_ids()[cells] - 1)).all()      # the global numbering restricted
    g = np.cos(0.05 * np.arange(Vf.num_free_dofs)) + 2.0
    v = gg.DeviceVector.from_numpy(ctx, np.where(own == rank, g[l2g], -5.0))
    gg.consistent(v, V)
    assert (v.get() == g[l2g]).all(), "consistent! on the hexahedral CG space did not reproduce the owner values"

    def system(model, W):
        dO = gg.Measure(gg.Triangulation(model), degree)
        assem = gg.SparseMatrixAssembler(W, W)
        gg.assemble_matrix(gg.LaplaceBeltrami(dO), assem)
        A = gg.assemble_matrix(gg.ScalarMass(dO), assem, add=True)
        xyz = dO.quadrature_points(chart=False)[1]
        f = xyz[..., 0] * xyz[..., 1] + np.sin(3 * xyz[..., 2])
        return A, gg.assemble_vector(gg.Source(dO, f), W)

    A, rhs = system(local, V)
    Af, rhsf = system(full, Vf)
    owned = own == rank
    assert rel(rhs.get()[owned], rhsf.get()[l2g][owned]) < 1e-13          # owned rows are complete without any exchange
    S, Sf = A.to_scipy(), Af.to_scipy()
    xt = np.sin(0.01 * np.arange(Vf.num_free_dofs))
    assert rel((S @ xt[l2g])[owned], (Sf @ xt)[l2g][owned]) < 1e-13
    ns, nsf = gg.CGSolver(rtol=1e-12).numerical_setup(A), gg.CGSolver(rtol=1e-12).numerical_setup(Af)
    x = ns.solve(rhs).get()
    xf = nsf.solve(rhsf).get()
    e_x = rel(x, xf[l2g])
    assert e_x < 1e-9, f"distributed shell solve differs from the single-GPU solve by {e_x}"
    assert comm.error_epoch() == 0
    assert abs(ns.iterations - nsf.iterations) <= 1
    dist.barrier()
    return {"n": n, "order": order, "cells": int(full.num_cells), "dofs": int(Vf.num_free_dofs), "solution": e_x,
            "pcg_iters": int(ns.iterations), "pcg_iters_single_gpu": int(nsf.iterations)}


def main():
    """torchrun worker.  GG_DIST_BACKEND=gloo + GG_DIST_SHARE_GPU=1 are for the world-of-one run of the partitioned code path
    (GG_FORCE_MULTI=1); several ranks must never share a GPU (their kernels wait on one another)."""
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    share = os.environ.get("GG_DIST_SHARE_GPU") == "1"
    assert not share or world == 1, "ranks whose kernels wait on one another must not share a GPU"
    local = 0 if share else local
    torch.cuda.set_device(local)
    if os.environ.get("GG_DIST_BACKEND", "nccl") == "gloo":
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gg = load_package()
    ctx = gg.get_context(local)
    n, p = int(sys.argv[1]) if len(sys.argv) > 1 else 6, int(sys.argv[2]) if len(sys.argv) > 2 else 1
    r = run_checks(gg, ctx, rank, world, n, p)
    if world > 1 and p == 2:
        rs = run_shell_checks(gg, ctx, rank, world)
        if rank == 0:
            print(f"SHELL_OK {rs}")
    if rank == 0:
        print(f"DIST_OK world={world} n={n} p={p} wave_err={r['wave']:.2e} swe_err={r['swe']:.2e} pv_err={r['pv']:.2e} "
              f"adv_err={r['advection']:.2e} pcg_iters={r['pcg_iters']} (single GPU {r['pcg_iters_single_gpu']}) solves={r['solves']}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
