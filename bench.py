#!/usr/bin/env python
"""bench.py -- headline benchmark of the cubed-sphere assembly hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Metric (BASELINE.json): cells/s assembled (residual + Jacobian).  One step = one fused Laplace-Beltrami
Jacobian (CSR values) + residual (DOF vector) assembly over all cells of the rank's mesh slice:
CG-Q2, quadrature degree 18 (10x10 Gauss points per cell) -- the reference's own choice 6*(p+1)
(test/Laplacian/LaplaceBeltrami.jl:29) -- on a 256x256-per-panel cubed sphere (393 216 cells) at N=1.
N>1 is STRONG scaling by default (the named configuration, BASELINE configs[3]: 256x256 per panel partitioned over the N
ranks): every rank assembles its slice of the reference's linear partition (owned cells + vertex-adjacent ghost layer, so
that owned rows are complete without any reverse exchange); there is no data-path collective in assembly.  The weak-scaling
numbers (round(256*sqrt(N))^2 cells per panel) are reported beside it under `weak`; `--scaling weak` makes them the headline.
A small partitioned-vs-single-GPU parity run (tests/dist_worker.py) is embedded in every N>1 run (`dist_parity`).
`--impl reference` times the CPU restatement of the reference algorithm (oracle/c, OpenMP over cells) on the
box's host cores: the real reference (Julia + Gridap) cannot run in this image (no Julia, no network).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BASE, ORDER, DEGREE = 256, 2, 18
SHELL_N_P4EST = 64   # BASELINE configs[4]: 3-D shell, p4est-refined (level 6: 64^3 octants per tree), order 2
SEED = 20261019
# algorithmic FLOPs per cell, Q = 100, m = 9, D = 2
FLOPS_JAC = 100 * (81 * 4 + 9 * 8 + 60)          # 45 600: element matrix
FLOPS_RES = 100 * (4 * 9 * 2 + 8 + 60)           # 14 000: residual
# algorithmic bytes per cell (§8d): ids + chart box + CSR values written once + gather/scatter of the residual
BYTES_JAC = 36 + 64 + 8 * 64
BYTES_RES = 36 + 72 + 32
METRIC = "cells/s assembled (residual+Jacobian)"


def mesh_n(nranks, scaling="strong"):
    return N_BASE if scaling == "strong" else int(round(N_BASE * math.sqrt(nranks)))


def workload_config(n, world, scaling):
    """`config` of the JSON line: the same keys and values in both arms (`--impl b200` and `--impl reference`)."""
    return {"workload": f"laplace_beltrami_cgq2_deg18_n{n}_per_panel", "cells": 6 * n * n, "order": ORDER,
            "quadrature_degree": DEGREE, "partition": f"linear x{world}" if world > 1 else "single", "scaling_mode": scaling,
            "l2": "working set per step (element buffer, maps and CSR values of the rank's slice; ~0.6 GB at N=1) exceeds the 126 MB "
                  "L2 for N <= 4; at N = 8 (75 MB per rank) an L2 flush (write of a 256 MB buffer) runs between timed steps"}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # samples under load = upper half of the observed clocks
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- shallow-water initial states: gridapgeosciences.jl_b200/initial_conditions.py (Williamson 2 as in the reference's
# test/Geophysical/Williamson_functions.jl; Galewsky et al. 2004 restated from the paper -- BASELINE configs[3] names it, the
# reference does not contain it)


def swe_leg(gg, ctx, torch, stream, e0, e1, n=256, p=2, steps=10, dist=None, rank=0, world=1, state="galewsky", partition="linear"):
    """RK steps/s of the shallow-water DAE stepper.  world > 1: the n x n-per-panel model is partitioned over the ranks
    (linear partition + ghost layer of the reference; partition="strips": the opt-in panel strips of gg_plan_create_strips);
    halo exchange and dot products go through peer memory inside the solver kernels, the time is the max over ranks."""
    ctx.set_async(False)
    mesh = gg.CubedSphereMesh(1.0)
    keep = []
    if world > 1:
        plan = gg.PartitionPlan(n, world, rank, partition=partition)
        comm = gg.PeerWindow(ctx, rank, world, gg.window_doubles(plan, p))
        comm.connect()
        model = gg.partitioned_model(mesh, plan, comm, ctx=ctx)
        ncells = plan.num_cells_global
        keep = [plan, comm]
    else:
        model = gg.AtlasDiscreteModel(mesh, n=n, ctx=ctx)
        ncells = model.num_cells
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.raviart_thomas, float, p), conformity="HDiv")
    Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="L2")
    ic = gg.initial_conditions
    vel, dep, H_ref = ((ic.galewsky_velocity, ic.galewsky_depth, 1.0e4 / ic.A_E) if state == "galewsky" else
                       (ic.williamson2_velocity, ic.williamson2_depth, ic.H0_W2))
    x0 = np.concatenate([gg.interpolate(vel, V).get(), gg.interpolate(dep, Q).get()])
    dt = math.sqrt(4 * math.pi / ncells) * 0.1 / (p * math.sqrt(ic.GRAVITY * H_ref))          # TransientShallowWater.jl:147
    op = gg.ShallowWaterOperator(model, p, ic.coriolis, gravity=ic.GRAVITY)
    st = gg.RKStepper(gg.RungeKutta(gg.CGSolver(rtol=1e-12), None, dt, "EXRK_SSP_3_3"), op)
    st.set_state(x0)
    st.step(3)
    it0, so0 = st.stats()
    l0 = ctx.launches
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    st.step(steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        assert keep[1].error_epoch() == 0, "a peer-memory wait timed out"
    it1, so1 = st.stats()
    ph = {name: st.solver_phase_ns(w) for w, name in ((0, "rt_mass"), (1, "pv_mass"))}
    drift = float(np.abs(st.get_state() - x0).max() / np.abs(x0).max())
    out = {"workload": f"rotating shallow water ({'Galewsky jet, perturbed' if state == 'galewsky' else 'Williamson 2'}), {n}x{n} per panel, RT{p} x DG{p} + CG{p + 1} diagnostics, "
                       f"degree {4 * (p + 1)}, SSPRK(3,3), dt = 0.1 dx / (p sqrt(g H0))",
           "n_gpus": world, "cells": ncells, "cells_local_incl_ghosts": model.num_cells,
           "dofs_prognostic_local": st.nu + st.np_, "dofs_diagnostic_local": st.ny,
           "steps_per_s": 1e3 / ms, "ms_per_step": ms, "cell_stage_evaluations_per_s": 6 * ncells / (ms * 1e-3),
           "mass_solves_per_step": (so1 - so0) / steps, "pcg_iters_per_solve": (it1 - it0) / max(so1 - so0, 1),
           "launches_per_step": (ctx.launches - l0) / steps, "state_drift_after_steps": drift,
           "pcg_us_per_iter": {k: {"cells": v[0] / max(v[3], 1) / 1e3, "gather_dot": v[1] / max(v[3], 1) / 1e3,
                                   "update_precond": v[2] / max(v[3], 1) / 1e3} for k, v in ph.items()},
           "solver": "statically condensed element-by-element block-Jacobi PCG, rtol 1e-12, warm-started from the previous "
                     "solution" + ("; halo + all-reduce through CUDA-IPC peer memory inside the kernel" if world > 1 else "")}
    out["roofline"] = swe_roofline(out, model.num_cells, p)
    del st
    return out


def swe_roofline(leg, ncell_local, p):
    """HBM roofline of the dominant kernel of the shallow-water step, the persistent condensed PCG k_pcg_ebe (the largest share of the step in
    profiles/ncu_swe_r02.txt), per iteration of the RT mass solve, from the phase timers of THIS run (%globaltimer inside the
    kernel) and the algorithmic bytes of one iteration: every operand once."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    mb = 4 * (p + 1)                                  # facet DOFs of a cell (RT_p: p + 1 per facet)
    rows = ncell_local * 2 * (p + 1)                  # facet DOFs of the slice (2 facets per cell)
    ncls = min(ncell_local, 65536 if ncell_local >= 65536 else ncell_local)   # geometry classes (256^2: one per (column, row) pair)
    b_cells = ncls * 8.0 * (mb * (mb + 1) // 2) + ncell_local * (mb * 4 + mb * 8 + mb * 8)   # S once per class; ids, x_f in, S x out
    b_gather = rows * (2 * (4 + 8) + 3 * 8)           # two element sources per row (index + value); p, Ap, partial dots
    b_update = rows * (6 * 8 + (p + 1) * 8)           # x, r, z, p read / written; the row's slice of its (p+1)^2 block inverse
    it = leg["pcg_us_per_iter"]["rt_mass"]
    t_us = it["cells"] + it["gather_dot"] + it["update_precond"]
    tot = b_cells + b_gather + b_update
    return {"kernel": "k_pcg_ebe (condensed RT%d mass solve), one CG iteration" % p, "bound": "hbm", "unit": "GB/s",
            "achieved": tot / (t_us * 1e-6) / 1e9 if t_us else None, "peak": hbm,
            "frac": tot / (t_us * 1e-6) / 1e9 / hbm if t_us else None,
            "algorithmic_bytes_per_iteration": tot, "us_per_iteration": t_us,
            "phases": {"cells": {"bytes": b_cells, "us": it["cells"], "gbs": b_cells / (it["cells"] * 1e-6) / 1e9 if it["cells"] else None},
                       "gather_dot": {"bytes": b_gather, "us": it["gather_dot"], "gbs": b_gather / (it["gather_dot"] * 1e-6) / 1e9 if it["gather_dot"] else None},
                       "update_precond": {"bytes": b_update, "us": it["update_precond"], "gbs": b_update / (it["update_precond"] * 1e-6) / 1e9 if it["update_precond"] else None}},
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s", "traffic": None,
            "from_profile": {"file": "profiles/ncu_swe_r02.txt", "k_pcg_ebe<12,12,1>": {"dram_bytes_per_launch": 4.537e9, "ms": 1.285,
                                                                                  "dram_gbs": 3531, "l2_hit_pct": 40.4},
                             "note": "ncu --set full, round 2, N=1, one launch = one whole RT2 mass solve"}}


def adv_leg(gg, ctx, torch, stream, e0, e1, n=128, p=2, steps=200, dist=None, rank=0, world=1):
    """RK steps/s of DG upwind advection (BASELINE configs[2]: 128x128 per panel, DG-Q2, solid-body wind of
    AdvectionUpwinding.jl:100-102, SSPRK(3,3)).  world > 1: partitioned model, the ghost cells of every stage state are
    filled through peer memory (one small cooperative kernel per stage); max over ranks."""
    ctx.set_async(False)
    mesh = gg.CubedSphereMesh(1.0)
    keep = []
    if world > 1:
        plan = gg.PartitionPlan(n, world, rank)
        comm = gg.PeerWindow(ctx, rank, world, gg.window_doubles(plan, p))
        comm.connect()
        model = gg.partitioned_model(mesh, plan, comm, ctx=ctx)
        ncells = plan.num_cells_global
        keep = [plan, comm]
    else:
        model = gg.AtlasDiscreteModel(mesh, n=n, ctx=ctx)
        ncells = model.num_cells
    wind = lambda X: np.stack([-X[..., 1], X[..., 0], 0 * X[..., 0]], -1)   # noqa: E731
    dt = 2 * math.pi / 20000 * 128 / n
    st = gg.RKStepper(gg.RungeKutta(None, None, dt, "EXRK_SSP_3_3"), gg.AdvectionOperator(model, p, wind))
    Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="L2")
    st.set_state(gg.interpolate(lambda X: np.exp(-(X[..., 1] ** 2 + X[..., 2] ** 2)), Q).get())
    st.step(20)
    l0 = ctx.launches
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    st.step(steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        assert keep[1].error_epoch() == 0, "a peer-memory wait timed out"
    out = {"workload": f"DG-Q{p} upwind advection (volume + skeleton terms), {n}x{n} per panel, degree {4 * p}, SSPRK(3,3)",
           "n_gpus": world, "cells": ncells, "cells_local_incl_ghosts": model.num_cells, "dofs_local": st.np_,
           "steps_per_s": 1e3 / ms, "ms_per_step": ms, "cell_stage_evaluations_per_s": 3 * ncells / (ms * 1e-3),
           "launches_per_step": (ctx.launches - l0) / steps}
    del st
    return out


def tswe_leg(gg, ctx, torch, stream, e0, e1, n=256, p=1, steps=10):
    """RK steps/s of the thermal shallow-water stepper (test/Tutorials/ThermalShallowWater.jl: order 1, thermogeostrophic balance,
    Measure(., 4*order); next-row caller of the same kernels)."""
    ctx.set_async(False)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n, ctx=ctx)
    ic = gg.initial_conditions
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.raviart_thomas, float, p), conformity="HDiv")
    Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="L2")
    x0 = np.concatenate([gg.interpolate(ic.williamson2_velocity, V).get(), gg.interpolate(ic.williamson2_depth, Q).get(),
                         gg.interpolate(ic.thermogeostrophic_weighted_buoyancy, Q).get()])
    dt = math.sqrt(4 * math.pi / model.num_cells) * 0.1 / (p * math.sqrt(ic.GRAVITY * ic.H0_W2))
    st = gg.RKStepper(gg.RungeKutta(gg.CGSolver(rtol=1e-12), None, dt, "EXRK_SSP_3_3"),
                      gg.ThermalShallowWaterOperator(model, p, ic.coriolis))
    st.set_state(x0)
    st.step(5)
    it0, so0 = st.stats()
    torch.cuda.synchronize()
    e0.record(stream)
    st.step(steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    it1, so1 = st.stats()
    drift = float(np.abs(st.get_state() - x0).max() / np.abs(x0).max())
    out = {"workload": f"thermal shallow water (thermogeostrophic balance), {n}x{n} per panel, RT{p} x DG{p} x DG{p} + CG{p + 1} "
                       f"diagnostics, degree {4 * p}, SSPRK(3,3)", "cells": model.num_cells, "dofs_prognostic": st.nu + st.np_,
           "steps_per_s": 1e3 / ms, "ms_per_step": ms, "pcg_iters_per_solve": (it1 - it0) / max(so1 - so0, 1),
           "state_drift_after_steps": drift}
    del st
    return out


def shell_leg(gg, ctx, torch, stream, e0, e1, n=48, order=2, degree=18, steps=10, dist=None, rank=0, world=1, ordering=0):
    """cells/s of the Laplace-Beltrami residual + Jacobian assembly on the 3-D shell (hexahedral Q2, 10^3 points per cell in the
    reference; here the tensor Gauss rule is factorised exactly into surface x radial integrals, gg_hex.cuh).  ordering = 2:
    the cell order of the reference's uniformly refined p8est forest (Morton inside each tree); world > 1: every rank takes its
    equal-count chunk of that order -- p4est's partition -- plus the vertex-adjacent ghost layer (BASELINE configs[4])."""
    model = gg.AtlasDiscreteModel(gg.CubedSphereWithThicknessMesh(1.0, 0.19), n=n, ctx=ctx, ordering=ordering)
    ncells_global, owned = model.num_cells, model.num_cells
    if world > 1:
        # every rank assembles its chunk of the (space-filling-curve) cell order (owned + ghost cells; no data-path collective)
        a, b = gg.partition_range(ncells_global, world, rank)
        full = model
        model = full.restrict(np.sort(np.concatenate([np.arange(a, b), full.ghost_cells(a, b)])))
        owned = b - a
        del full
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, order), conformity="H1")
    assem = gg.SparseMatrixAssembler(V, V)
    u = gg.DeviceVector.from_numpy(ctx, np.random.default_rng(SEED).standard_normal(V.num_free_dofs))
    b = gg.DeviceVector(ctx, V.num_free_dofs)
    form = gg.LaplaceBeltrami(dΩ, u=u)
    ctx.set_async(False)
    gg.assemble_matrix_and_vector(form, assem, b)       # builds the gather lists
    ctx.set_async(True)
    for _ in range(3):
        gg.assemble_matrix_and_vector(form, assem, b)
    ctx.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(steps):
        gg.assemble_matrix_and_vector(form, assem, b)
    e1.record(stream)
    ctx.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ctx.profile(True)
    for _ in range(steps):
        gg.assemble_matrix_and_vector(form, assem, b)
    ctx.synchronize()
    prof = {k: ctx.profile_query(k)[1] / steps for k in ("k_hex_tensor_fill", "k_hex_kron", "k_hex_residual", "k_gather_chunks", "k_gather_dofs")}
    ctx.profile(False)
    ctx.set_async(False)
    m = (order + 1) ** 3
    nnz = assem.matrix.nnz
    wr = 8.0 * (m * (m + 1) // 2 * model.num_cells)
    tensor = prof["k_hex_tensor_fill"] > 0
    out = {"workload": f"laplace_beltrami_hexq{order}_deg{degree}_shell_n{n} ({n} layers, r = 1, t = 0.19)"
                       + (", p8est cell order, equal-count SFC chunks + ghost layer" if ordering == 2 else ""), "n_gpus": world,
           "cells": ncells_global, "cells_local_incl_ghosts": model.num_cells, "dofs_local": V.num_free_dofs, "nnz_local": nnz,
           "ms_per_step": ms, "cells_per_s": ncells_global / (ms * 1e-3), "kernels_ms": prof,
           "numeric_phase": "tensor-structured (complete shell): CSR values from assembled surface x radial factors" if tensor else
                            "cell-wise (sub-mesh): packed 27x27 element matrices by Kronecker factors + atomic-free CSR fill",
           "roofline": ({"kernel": "k_hex_tensor_fill", "bound": "hbm", "unit": "GB/s", "achieved": 14.0 * nnz / (prof["k_hex_tensor_fill"] * 1e-3) / 1e9,
                         "algorithmic_bytes": "8 B value written + 6 B of maps read per non-zero"} if tensor else
                        {"kernel": "k_hex_kron<3>", "bound": "hbm", "unit": "GB/s",
                         "achieved": wr / (prof["k_hex_kron"] * 1e-3) / 1e9 if prof["k_hex_kron"] else None,
                         "algorithmic_bytes": "8 B x 378 packed entries per cell written once"})}
    del assem, model
    return out


def cpu_baseline_run(nthreads, target_seconds=12.0):
    """CPU restatement on a bounded sample of the n=256 workload: the first `ns` cells (panel-major order)."""
    from oracle import c_oracle, forms
    from oracle.mesh import CubedSphereMesh2D
    from oracle.spaces import CGSpace
    # a 64x64-per-panel mesh has the same per-cell work (Q2, degree 18); its cells are a bounded sample
    mesh = CubedSphereMesh2D(64, 1.0)
    V = CGSpace(mesh, ORDER)
    rowptr, colind = forms.csr_pattern(V, V)
    u = np.random.default_rng(SEED).standard_normal(V.num_free)
    ue = forms.gather(V, u)
    if nthreads <= 0:
        # all host cores this process may use (the box may export OMP_NUM_THREADS=1)
        nthreads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    threads = nthreads
    # calibrate the sample size on a small chunk
    probe = 1024
    t0 = time.perf_counter()
    c_oracle.lb_element_matrices(mesh.cell_chart_coords[:probe], 1.0, ORDER, DEGREE, nthreads)
    rate = probe / (time.perf_counter() - t0)
    ns = int(min(mesh.num_cells, max(probe, rate * target_seconds * 0.5)))
    reps = max(1, int(round(rate * target_seconds * 0.5 / ns)))
    cc = mesh.cell_chart_coords[:ns]
    t0 = time.perf_counter()
    for _ in range(reps):
        Ke = c_oracle.lb_element_matrices(cc, 1.0, ORDER, DEGREE, nthreads)
        fe = c_oracle.lb_element_vectors(cc, 1.0, ORDER, DEGREE, ue[:ns], nthreads)
        vals = c_oracle.scatter_csr(V.cell_dofs[:ns], Ke, rowptr, colind, nthreads)
        r = c_oracle.scatter_vector(V.cell_dofs[:ns], fe, V.num_free, nthreads)
    dt = time.perf_counter() - t0
    assert np.isfinite(vals).all() and np.isfinite(r).all()
    return ns * reps / dt, threads, f"{reps} pass(es) over {ns} cells of the Q2/degree-18 Laplace-Beltrami residual+Jacobian " \
                                    f"(element matrices, element vectors, CSR/vector scatter), {threads} OpenMP thread(s), {dt:.1f} s"


def cpu_rk_baseline():
    """RK steps/s of the CPU restatement (oracle/c/gg_oracle_rk.c + oracle/c_oracle.py): the reference's structure -- stage
    residual re-integrated cell by cell, the potential-vorticity block of jac_y re-assembled at every diagnostic solve, mass
    solves by Jacobi-PCG at rtol 1e-12 (its tutorial's iterative option; LUSolver() is its default) -- serial and on all host
    cores.  Wave: the full 48x48 configuration.  Shallow water: ONE step on a 16x16-per-panel sample of the order-2
    configuration (same per-cell work), extrapolated to 256x256 by the cell count."""
    from oracle import c_oracle, swe as oswe
    from oracle import rk as ork
    from oracle.mesh import CubedSphereMesh2D
    allc = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    out = {"kind": "port", "solver": "Jacobi-PCG, rtol 1e-12 (test/Tutorials/ShallowWater.jl:175-176)"}
    mesh = CubedSphereMesh2D(48)
    # thread counts are passed explicitly (the box may export OMP_NUM_THREADS=1)
    for name, nt in (("serial", 1), ("all_cores", allc)):
        W = c_oracle.CWaveProblem(mesh, 1, nthreads=nt)
        x = np.concatenate([np.zeros(W.nu), ork.interpolate_scalar(W.T.Q, mesh, lambda X: 1.0 + 0.01 * np.exp(
            -5.0 * ((X[..., 0] - 1.0) ** 2 + X[..., 1] ** 2 + X[..., 2] ** 2)))])
        dt = ork.dx(mesh) * 0.1
        x = W.step(x, dt)
        t0 = time.perf_counter()
        for _ in range(2):
            x = W.step(x, dt)
        out.setdefault("wave", {})[name] = {"steps_per_s": 2 / (time.perf_counter() - t0), "cores": 1 if nt == 1 else allc}
    out["wave"]["sample"] = "2 SSPRK(3,3) steps of the full 48x48-per-panel RT1 x DG1 configuration after one warm-up step"
    ns = 16
    mesh = CubedSphereMesh2D(ns)
    for name, nt in (("serial", 1), ("all_cores", allc)):
        S = c_oracle.CShallowWaterProblem(mesh, 2, oswe.coriolis(), oswe._g, nthreads=nt)
        x0 = np.concatenate([ork.interpolate_rt(S.T.V, mesh, oswe.tangent_component(oswe.velocity())),
                             ork.interpolate_scalar(S.T.Q, mesh, oswe.depth())])
        dt = oswe.reference_dt(mesh, 2)
        t0 = time.perf_counter()
        S.step(x0, dt)
        el = time.perf_counter() - t0
        out.setdefault("shallow_water", {})[name] = {
            "steps_per_s_sample": 1 / el, "steps_per_s_extrapolated_to_n256": 1 / el * (ns / 256.0) ** 2,
            "cell_stage_evaluations_per_s": 6 * mesh.num_cells / el, "cores": 1 if nt == 1 else allc,
            "pcg_iters_per_solve": S.iters / max(S.solves, 1)}
    out["shallow_water"]["sample"] = (f"1 SSPRK(3,3) step (13 mass solves) of the order-2 configuration (RT2 x DG2 + CG3, degree 12, "
                                      f"Williamson 2) on a {ns}x{ns}-per-panel mesh; n=256 figure = sample rate x ({ns}/256)^2")
    return out


def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(args.steps, 1), args.warmup
    per_step = max(2.0, min(20.0, 120.0 / (steps + warm)))
    vals = []
    sample = ""
    for i in range(warm + steps):
        v, threads, sample = cpu_baseline_run(0, per_step)
        if i >= warm:
            vals.append(v)
    value = float(np.mean(vals))
    n = mesh_n(args.gpus, args.scaling)
    ncells = 6 * n * n
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * ncells / value, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n, args.gpus, args.scaling),
        "note": "CPU restatement of the reference algorithm (oracle/c, not Gridap: Julia is absent), all host cores of this "
                "process (OpenMP over cells stands in for the reference's MPI ranks); every step TIMES A BOUNDED SAMPLE of the "
                "workload's cells (see cpu_baseline.sample) and `value` is that sampled rate; ms_per_step extrapolates it to the "
                "full mesh of `config`",
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


class L2Flush:
    """Write of a buffer larger than the 126 MB L2 between timed steps, for rank slices whose working set would otherwise stay
    L2-resident from one step to the next (strong scaling at N >= 4)."""

    def __init__(self, torch, stream, device):
        self.torch, self.stream = torch, stream
        with torch.cuda.stream(stream):
            self.buf = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=device)

    def __call__(self):
        with self.torch.cuda.stream(self.stream):
            self.buf.fill_(1.0)


def timed_steps(torch, stream, step, K, flush=None):
    """ms per step on the library's stream with CUDA events.  Without flush: one event pair around K back-to-back steps.
    With flush: the L2 is flushed before every step and each step has its own event pair (the flush is outside them)."""
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(K):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for a, b in ev:
        flush()
        a.record(stream)
        step()
        b.record(stream)
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / K


def max_over_ranks(torch, dist, world, x):
    if world == 1:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def assembly_leg(gg, ctx, torch, dist, stream, n, rank, world, W, K, profile=True):
    """One rank's slice of the n x n-per-panel Laplace-Beltrami residual + Jacobian assembly (Q2, degree 18)."""
    mesh = gg.CubedSphereMesh(1.0)
    keep = None
    if world > 1:
        # linear partition of the reference + vertex-adjacent ghost layer; local DOF numbering = the global numbering
        # restricted to the local cells (compact: vectors and row pointers have the size of the slice)
        keep = gg.PartitionPlan(n, world, rank)
        model = gg.partitioned_model(mesh, keep, None, ctx=ctx)
        ncells_global, owned = keep.num_cells_global, keep.num_owned_cells
    else:
        model = gg.AtlasDiscreteModel(mesh, n=n, ctx=ctx)
        ncells_global = owned = model.num_cells
    dΩ = gg.Measure(gg.Triangulation(model), DEGREE)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, ORDER), conformity="H1")
    assem = gg.SparseMatrixAssembler(V, V)           # device CSR pattern, built once
    A = assem.matrix
    u_host = torch.from_numpy(np.random.default_rng(SEED + rank).standard_normal(V.num_free_dofs)).pin_memory()
    u = gg.DeviceVector.from_numpy(ctx, u_host.numpy())
    b = gg.DeviceVector(ctx, V.num_free_dofs)
    form = gg.LaplaceBeltrami(dΩ, u=u)

    def step():
        gg.assemble_matrix_and_vector(form, assem, b)

    ws_bytes = 8.0 * (45 * model.num_cells + A.nnz) + 6.0 * 81 * model.num_cells
    flush = L2Flush(torch, stream, ctx.device) if ws_bytes < 2 * 126e6 else None
    ctx.set_async(True)
    for _ in range(W):
        step()
    ctx.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = ctx.launches
    ms = timed_steps(torch, stream, step, K, flush)
    launches = ctx.launches - l0
    ms = max_over_ranks(torch, dist, world, ms)
    out = {"n": n, "cells": ncells_global, "owned": owned, "cells_local_incl_ghosts": model.num_cells, "ms_per_step": ms,
           "value": ncells_global / (ms * 1e-3), "launches": launches, "dofs_local": V.num_free_dofs, "nnz_local": A.nnz,
           "l2_flush_between_steps": flush is not None}
    if profile:
        ctx.profile(True)
        for _ in range(K):
            if flush:
                flush()
            step()
        ctx.synchronize()
        out["prof"] = {name: ctx.profile_query(name) for name in ASSEMBLY_KERNELS}
        ctx.profile(False)
    ctx.set_async(False)
    out["objs"] = (model, V, assem, A, u, u_host, b, form, step, keep, flush)
    return out


# FP64 operations the fused element kernel EXECUTES per cell (Q2, 10x10 points; counted from its SASS: 545 static DFMA, the
# sum-factorised contractions run ~5.1 k DFMA per thread, plus the residual mat-vec, the metric and the rsqrt refinement)
EXECUTED_FLOPS_PER_CELL = 2 * 5100 + 2 * 81 + 12 * 100 + 20 * 60


def assembly_roofline(prof, ncell_local, nnz, ms_per_step, fp64_peak, hbm_peak, have_peaks):
    """Roofline block of the headline step from the per-kernel CUDA-event times of THIS run (taken in a separate profiled
    pass over the same steps).  The fused tile kernel does the FP64 element arithmetic AND writes the CSR values, so it is
    reported against both bounds: the DFMA probe run in this process and the measured copy bandwidth."""
    ms = {k: (v[1] / v[0] if v[0] else 0.0) for k, v in prof.items()}
    fused = ms.get("k_lb_tile", 0.0) > 0.0
    k_name = "k_lb_tile" if fused else "k_lb_fast"
    k_ms = ms[k_name]
    executed = EXECUTED_FLOPS_PER_CELL * ncell_local / (k_ms * 1e-3) / 1e12 if k_ms else None
    naive = (FLOPS_JAC + FLOPS_RES) * ncell_local / (k_ms * 1e-3) / 1e12 if k_ms else None
    work_ms = EXECUTED_FLOPS_PER_CELL * ncell_local / (fp64_peak * 1e12) * 1e3 if fp64_peak else None
    # algorithmic HBM bytes of the step: every CSR value and residual entry written once, DOF ids and state read once
    alg_bytes = 8.0 * nnz + (36 + 72 + 32) * ncell_local
    out = {
        "kernel": ("k_lb_tile<3,10> (fused: Q2 element matrices in registers / shared memory -> CSR values and residual, one CTA per "
                   "16x8-cell tile)" if fused else "k_lb_fast<3,10> (fused Q2 element matrix + residual, one thread per cell)"),
        "bound": "fp64", "unit": "TFLOP/s",
        "achieved": executed, "peak": fp64_peak, "frac": executed / fp64_peak if (executed and fp64_peak) else None,
        "achieved_is": "FP64 operations the kernel executes (sum-factorised: ~13 kFLOP per cell, EXECUTED_FLOPS_PER_CELL in bench.py) "
                       "/ kernel time measured in this run with CUDA events on the library's stream",
        "peak_source": "FP64 DFMA probe run in this process (gg_measure_fp64_peak, also `fp64_tflops_probe`); MEASURED_PEAKS.json has "
                       "no FP64 figure",
        "algorithmic": {"tflops": naive, "frac": naive / fp64_peak if (naive and fp64_peak) else None,
                        "note": "SURVEY.md §8(d)'s per-cell count of the reference algorithm (45 600 + 14 000 FLOP) / the same kernel "
                                "time: exceeds 1 because the kernel does not execute that many operations"},
        "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_per_step if ms_per_step else None,
        "step_level": {"fp64_work_ms_at_probe_peak": work_ms, "frac_of_step": work_ms / ms_per_step if (work_ms and ms_per_step) else None,
                       "kernels_ms": ms},
        "hbm": {"algorithmic_bytes_per_step": alg_bytes, "achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms else None,
                "peak_gbs": hbm_peak, "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak if k_ms else None,
                "algorithmic_bytes": "8 B x nnz (every CSR value written once) + 140 B per cell (DOF ids, state gather, residual)",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if have_peaks else "fallback 6650 GB/s"},
        # dram__bytes_read + write of the dominant kernel per launch: the ncu capture of this exact workload (N=1, full size), else unknown
        "traffic": ((PROFILE_FIGURES["k_lb_tile"]["dram_bytes_per_launch"] if fused else
                     PROFILE_FIGURES["two_kernel_path"]["k_lb_fast"]["dram_bytes_per_launch"]) if ncell_local == 393216 else None),
        "traffic_source": "profiles/ncu_lb_tile_r02.txt (ncu --set full, not measured by this run)",
        "from_profile": PROFILE_FIGURES,
    }
    if not fused:
        g_ms = ms.get("k_gather_chunks", 0.0)
        gather_bytes = 8.0 * (45 * ncell_local + nnz)
        out["csr_fill"] = {"kernel": "k_gather_chunks", "bound": "hbm", "unit": "GB/s", "kernel_ms": g_ms,
                           "achieved": gather_bytes / (g_ms * 1e-3) / 1e9 if g_ms else None, "peak": hbm_peak,
                           "frac": gather_bytes / (g_ms * 1e-3) / 1e9 / hbm_peak if g_ms else None,
                           "algorithmic_bytes": "8 B x (45 packed element entries per cell read + every CSR value written once)"}
    return out


# figures that only a profiler gives (ncu capture of this command, committed under profiles/); NOT measured by bench.py itself
PROFILE_FIGURES = {"file": "profiles/ncu_lb_tile_r02.txt", "workload": "N=1, 393 216 cells, ncu --set full of this command, round 2",
                   "k_lb_tile": {"dram_bytes_per_launch": 266.0e6, "dram_bytes_per_cell": 676.5, "fp64_pipe_active_pct": 56.1,
                                 "note": "65 MB read + 201 MB written; algorithmic: 8 B x nnz = 201 MB of CSR values"},
                   "k_lb_tile_groups": {"dram_bytes_per_launch": 100.4e6, "note": "tile-border rows: 84 MB read, 16 MB written"},
                   "two_kernel_path": {"k_lb_fast": {"dram_bytes_per_launch": 149.0e6, "fp64_pipe_active_pct": 81.3},
                                       "k_gather_chunks": {"dram_bytes_per_launch": 574.6e6}}}


def assembly_limiter(leg, roofline, world):
    """What bounds the strong-scaled assembly step on this rank count, from this run's own numbers."""
    k = roofline["step_level"]["kernels_ms"]
    ksum = sum(v for v in k.values() if v)
    ghost = leg["cells_local_incl_ghosts"] / max(leg["owned"], 1) - 1.0
    gap = leg["ms_per_step"] - ksum
    ctas = leg["cells_local_incl_ghosts"] / 128.0
    if ctas < 3 * 148:
        nk = sum(1 for v in k.values() if v)
        what = (f"less than one wave of CTAs: the slice is {ctas:.0f} 128-cell CTAs for 444 CTA slots (3 per SM), so the element kernel "
                f"runs as ONE wave at the FP64 pipe's rate with nothing to overlap its start-up, tail and the CSR fill behind it; the "
                f"{nk} kernels of a step sum to {ksum * 1e3:.1f} us of a {leg['ms_per_step'] * 1e3:.1f} us step (L2 flushed before every "
                f"step); FP64 work of the slice at the probe's peak: {roofline['step_level']['fp64_work_ms_at_probe_peak'] * 1e3:.1f} us")
    elif gap > 0.25 * leg["ms_per_step"]:
        what = f"launch / tail latency: the kernels of a step sum to {ksum * 1e3:.1f} us of a {leg['ms_per_step'] * 1e3:.1f} us step"
    elif ghost > 0.1:
        what = f"redundant ghost-layer cells: {100 * ghost:.1f} % more cells than owned"
    else:
        what = (f"the same kernels as at N=1 (FP64 pipe + CSR fill) on 1/N of the cells; {ctas / 444:.1f} waves of CTAs (wave quantisation) "
                f"and {100 * ghost:.1f} % ghost cells")
    return {"what": what, "kernel_ms_sum": ksum, "ghost_cell_overhead": ghost, "cells_per_sm": leg["cells_local_incl_ghosts"] / 148.0}


ASSEMBLY_KERNELS = ("k_lb_tile", "k_lb_tile_groups", "k_lb_fast", "k_gather_chunks", "k_gather_dofs")


def bind_host_to_gpu(torch, device):
    """Multi-rank runs: run this rank on the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned allocation (first touch
    puts the pinned pages there), so that the 8 concurrent device -> host copies of the e2e leg do not cross the socket link.
    Returns what was done for the JSON line; a box without the sysfs entries is left alone."""
    try:
        pr = torch.cuda.get_device_properties(device)
        bus = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return {"pci": bus, "numa_node": node, "bound": False}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"pci": bus, "numa_node": node, "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"pci": bus, "numa_node": node, "bound": True, "cpus": len(cpus)}
    except Exception as ex:   # no sysfs, no permission, older torch: not an error
        return {"bound": False, "why": str(ex)[:120]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = the named 256x256-per-panel mesh partitioned over the ranks (default); weak = 256 sqrt(N)")
    ap.add_argument("--no-extras", action="store_true", help="skip the RK-steps/s, e2e and CPU-baseline legs")
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: whatever libraries print on file descriptor 1 (the NCCL version banner
    # of the first communicator, ...) goes to stderr; the line itself is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        run_reference(args, emit)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    numa = bind_host_to_gpu(torch, local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from __graft_entry__ import load_package
    gg = load_package()
    ctx = gg.get_context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    W, K = max(args.warmup, 3), max(args.steps, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    # ---- headline: the rank's slice of the mesh, device-resident timing --------------------------------------------
    n = mesh_n(world, args.scaling)
    sampler = ClockSampler(local) if rank == 0 else None
    leg = assembly_leg(gg, ctx, torch, dist, stream, n, rank, world, W, K)
    model, V, assem, A, u, u_host, b, form, step, _plan, flush = leg["objs"]
    ms_per_step, value, launches, prof = leg["ms_per_step"], leg["value"], leg["launches"], leg["prof"]
    total_owned, ncells_global, ncell_local = leg["cells"], leg["cells"], leg["cells_local_incl_ghosts"]
    # keep the GPU under the same load until the sampler has seen it (the timed region itself is a few ms long)
    ctx.set_async(True)
    t_end = time.perf_counter() + 0.6
    while time.perf_counter() < t_end:
        for _ in range(K):
            step()
        ctx.synchronize()
    ctx.set_async(False)
    clocks = sampler.stop() if sampler else None
    fp64_peak = ctx.measure_fp64_peak()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roofline = assembly_roofline(prof, ncell_local, A.nnz, ms_per_step, fp64_peak, hbm_peak, bool(peaks))

    line = {
        "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(n, world, args.scaling),
        "sizes": {"dofs_local": V.num_free_dofs, "nnz_local": A.nnz, "cells_local_incl_ghosts": ncell_local,
                  "cells_owned_rank0": leg["owned"], "l2_flush_between_steps": leg["l2_flush_between_steps"]},
        "roofline": roofline, "fp64_tflops_probe": fp64_peak, "clocks": clocks, "gpu_launches": launches,
    }
    if world > 1:
        line["limiter"] = assembly_limiter(leg, roofline, world)

    # ---- opt-in variant: one element matrix per geometry class (the 6 panels are congruent), same numbers bit for bit ------
    if not args.no_extras and world == 1:
        try:
            ctx.set_async(True)
            ctx.set_geometry_classes(True)
            assem_c = gg.SparseMatrixAssembler(V, V)
            b_c = gg.DeviceVector(ctx, V.num_free_dofs)
            for _ in range(W):
                gg.assemble_matrix_and_vector(form, assem_c, b_c)
            ctx.synchronize()
            e0.record(stream)
            for _ in range(K):
                gg.assemble_matrix_and_vector(form, assem_c, b_c)
            e1.record(stream)
            torch.cuda.synchronize()
            ms_c = e0.elapsed_time(e1) / K
            ctx.profile(True)
            for _ in range(K):
                gg.assemble_matrix_and_vector(form, assem_c, b_c)
            ctx.synchronize()
            pc = {name: ctx.profile_query(name) for name in ("k_lb_tile", "k_lb_tile_groups", "k_lb_fast", "k_lb_residual_classes", "k_gather_chunks", "k_gather_dofs")}
            ctx.profile(False)
            same = bool((assem_c.matrix.values() == A.values()).all())
            line["geometry_classes"] = {
                "value": total_owned / (ms_c * 1e-3), "unit": "cells/s", "ms_per_step": ms_c, "bitwise_equal_to_headline_path": same,
                "kernels_ms": {k: v[1] / max(v[0], 1) for k, v in pc.items()},
                "note": "opt-in (gg_set_geometry_classes): cells with the same chart box share their intrinsic element matrix, so the "
                        "FP64 kernel runs on one representative per (column, row) class = 1/6 of the cells; every call still "
                        "recomputes the class matrices and refills all CSR values and the residual.  Reported beside `value`, "
                        "which is the per-cell path."}
            del assem_c, b_c
        except Exception as ex:
            line["geometry_classes"] = {"error": str(ex)}
        finally:
            ctx.set_geometry_classes(False)

    # ---- end-to-end through the public API with HOST buffers ------------------------------------------------------
    if not args.no_extras:
        ctx.set_async(False)
        vals_host = torch.empty(A.nnz, dtype=torch.float64).pin_memory()
        b_host = torch.empty(V.num_free_dofs, dtype=torch.float64).pin_memory()

        def e2e_step():
            u.set(u_host.numpy())                    # H2D of the step's input (pinned)
            gg.assemble_matrix_and_vector(form, assem, b)
            A.values(out=vals_host.numpy())          # D2H of the Jacobian values
            b.get(out=b_host.numpy())                # D2H of the residual

        for _ in range(3):
            e2e_step()
        barrier()
        Ke2e = max(3, min(K, 10))
        e0.record(stream)
        for _ in range(Ke2e):
            e2e_step()
        e1.record(stream)
        barrier()
        ms2 = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms2], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms2 = float(t.item())
        line["e2e"] = {"host_numa_binding_rank0": numa,
                       "value": total_owned / (ms2 / Ke2e * 1e-3), "unit": "cells/s", "h2d_bytes_per_step": 8 * V.num_free_dofs,
                       "d2h_bytes_per_step": 8 * (A.nnz + V.num_free_dofs), "ms_per_step": ms2 / Ke2e,
                       "api": "DeviceVector.set(host u) + assemble_matrix_and_vector + CSRMatrix.values(host) + DeviceVector.get(host)",
                       "aggregate_d2h_gbs": world * 8 * (A.nnz + V.num_free_dofs) / (ms2 / Ke2e * 1e-3) / 1e9,
                       "bound": ("PCIe: the 201 MB of CSR values dominate (the kernels are 0.28 ms of the step)" if world == 1 else
                                 "host side of PCIe: the ranks' device -> host copies of their CSR slices run concurrently; the "
                                 "aggregate rate the box delivers (aggregate_d2h_gbs; 92-161 GB/s on the boxes of this pool, against "
                                 "48 GB/s for one GPU) sets the step time, not the kernels")}

        # the same call when the Jacobian stays on the device for a device-side solver (GPULinearSolver in INTEGRATION.md):
        # only the state goes in and the residual comes out
        def e2e_resident_step():
            u.set(u_host.numpy())
            gg.assemble_matrix_and_vector(form, assem, b)
            b.get(out=b_host.numpy())

        for _ in range(3):
            e2e_resident_step()
        barrier()
        e0.record(stream)
        for _ in range(Ke2e):
            e2e_resident_step()
        e1.record(stream)
        barrier()
        ms3 = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms3], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms3 = float(t.item())
        line["e2e"]["matrix_resident"] = {"value": total_owned / (ms3 / Ke2e * 1e-3), "unit": "cells/s", "ms_per_step": ms3 / Ke2e,
                                          "h2d_bytes_per_step": 8 * V.num_free_dofs, "d2h_bytes_per_step": 8 * V.num_free_dofs,
                                          "note": "Jacobian left in HBM for a device-side linear solver; state in, residual out"}

    # ---- RK steps/s (second half of BASELINE.json's metric): linear wave, 48x48 per panel, RT1 x DG1 -------------
    if not args.no_extras and rank == 0:
        try:
            ctx.set_async(False)
            wm = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=48, ctx=ctx)
            dt = gg.dx(wm) * 0.1
            out = {}
            for tab in ("EXRK_SSP_3_3", "EXRK_SSP_2_2"):
                st = gg.WaveStepper(gg.RungeKutta(gg.CGSolver(rtol=1e-12), None, dt, tab), gg.WaveOperator(wm, 1))
                # the reference's initial state: u = 0, h = 1 + 0.01 exp(-5 |x - (1,0,0)|^2)  (TransientWaveEquation.jl:19-25)
                wq = gg.TestFESpace(wm, gg.ReferenceFE(gg.lagrangian, float, 1), conformity="L2")
                x0 = np.zeros(st.nu + st.np_)
                x0[st.nu:] = gg.interpolate(lambda X: 1.0 + 0.01 * np.exp(-5.0 * ((X[..., 0] - 1.0) ** 2 + X[..., 1] ** 2 + X[..., 2] ** 2)), wq).get()
                st.set_state(x0)
                st.step(10)
                e0.record(stream)
                st.step(100)
                e1.record(stream)
                torch.cuda.synchronize()
                it, so = st.stats()
                out[tab] = {"steps_per_s": 100 / (e0.elapsed_time(e1) * 1e-3), "pcg_iters_per_solve": it / max(so, 1)}
            line["rk"] = {"workload": "linear wave, 48x48 per panel, RT1 x DG1, degree 10, dt = 0.1 dx, Gaussian depth bump of the reference test", "dofs": st.nu + st.np_, **out}
            del st, wm
        except Exception as ex:  # the headline line must still be printed
            line["rk"] = {"error": str(ex)}
        # nonlinear rotating shallow water (DAE: 3 diagnostic solves + 3 slope solves per SSPRK3 step), BASELINE configs[3]:
        # 256x256 per panel, order 2 (RT2 x DG2 prognostic, CG3 x RT2 x DG2 diagnostic), degree 12, Williamson-2 state
        if world == 1:
            try:
                line["rk"]["advection"] = adv_leg(gg, ctx, torch, stream, e0, e1)
            except Exception as ex:
                line["rk"]["advection"] = {"error": str(ex)}
        if world == 1:
            try:
                line["rk"]["shallow_water"] = swe_leg(gg, ctx, torch, stream, e0, e1)
            except Exception as ex:
                line["rk"]["shallow_water"] = {"error": str(ex)}
            try:
                line["rk"]["thermal_shallow_water"] = tswe_leg(gg, ctx, torch, stream, e0, e1)
            except Exception as ex:
                line["rk"]["thermal_shallow_water"] = {"error": str(ex)}


    # ---- 3-D shell (BASELINE configs[4]): hexahedral Q2 Laplace-Beltrami residual + Jacobian, 48^3 cells per panel ---------
    if not args.no_extras and rank == 0 and world == 1:
        try:
            line["shell"] = shell_leg(gg, ctx, torch, stream, e0, e1)
        except Exception as ex:
            line["shell"] = {"error": str(ex)}
        try:
            line["shell_p4est"] = shell_leg(gg, ctx, torch, stream, e0, e1, n=SHELL_N_P4EST, steps=5, ordering=2)
        except Exception as ex:
            line["shell_p4est"] = {"error": str(ex)}

    # ---- N > 1: partitioned-vs-single-GPU parity on a small mesh, then the RK legs on the partitioned models ----------------
    if world > 1:
        del leg, model, V, assem, A, u, b, form, step
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import dist_worker
            dp = {}
            for (pn, pp) in ((6, 1), (5, 2)):
                r = dist_worker.run_checks(gg, ctx, rank, world, pn, pp)
                dp[f"n{pn}_p{pp}"] = r
            try:
                shell_par = dist_worker.run_shell_checks(gg, ctx, rank, world)
            except Exception as ex:
                shell_par = {"error": str(ex)[-300:]}
            line["dist_parity"] = {"ok": "error" not in shell_par, "shell_p8est": shell_par, "tolerances": {"wave": 1e-10, "swe": 1e-9, "pv": 1e-8, "thermal_swe": 1e-9, "advection": 1e-12},
                                   "wave": max(v["wave"] for v in dp.values()), "swe": max(v["swe"] for v in dp.values()),
                                   "advection": max(v["advection"] for v in dp.values()),
                                   "pcg_iters_equal": all(v["pcg_iters_equal"] for v in dp.values()),
                                   "consistent_exact": all(v["consistent_exact"] for v in dp.values()), "cases": dp,
                                   "what": "tests/dist_worker.py run_checks: partitioned steppers (halo + all-reduce through peer memory) "
                                           "against the single-GPU run of the same problem; max relative state difference after 3-4 steps"}
        except Exception as ex:
            line["dist_parity"] = {"ok": False, "error": str(ex)[-400:]}
    if not args.no_extras and world > 1:
        nn = mesh_n(world, args.scaling)
        try:
            sw = swe_leg(gg, ctx, torch, stream, e0, e1, n=nn, dist=dist, rank=rank, world=world)
        except Exception as ex:
            sw = {"error": str(ex)}
        line.setdefault("rk", {})["shallow_water"] = sw
        try:
            ss = swe_leg(gg, ctx, torch, stream, e0, e1, n=nn, dist=dist, rank=rank, world=world, partition="strips")
            line["rk"]["shallow_water_panel_strips"] = {k: ss[k] for k in ("ms_per_step", "steps_per_s", "cells_local_incl_ghosts",
                                                                           "pcg_iters_per_solve", "pcg_us_per_iter") if k in ss}
            line["rk"]["shallow_water_panel_strips"]["what"] = ("the same step on the opt-in partition of gg_plan_create_strips (same rows of "
                                                                "cells on all six panels) instead of the reference's linear cut")
        except Exception as ex:
            line["rk"]["shallow_water_panel_strips"] = {"error": str(ex)}
        try:
            ad = adv_leg(gg, ctx, torch, stream, e0, e1, n=128 if args.scaling == "strong" else int(round(128 * math.sqrt(world))),
                         dist=dist, rank=rank, world=world)
        except Exception as ex:
            ad = {"error": str(ex)}
        line["rk"]["advection"] = ad
        try:
            # strong: the p8est-ordered 64^3-per-tree forest cut into `world` equal chunks of the space-filling curve
            if args.scaling == "strong":
                sh = shell_leg(gg, ctx, torch, stream, e0, e1, n=SHELL_N_P4EST, steps=5, dist=dist, rank=rank, world=world, ordering=2)
            else:
                sh = shell_leg(gg, ctx, torch, stream, e0, e1, n=int(round(48 * world ** (1.0 / 3.0))), dist=dist, rank=rank, world=world)
        except Exception as ex:
            sh = {"error": str(ex)}
        line["shell_p4est" if args.scaling == "strong" else "shell"] = sh
        # the other scaling mode beside the headline: same legs, fewer steps
        other = "weak" if args.scaling == "strong" else "strong"
        try:
            no = mesh_n(world, other)
            lw = assembly_leg(gg, ctx, torch, dist, stream, no, rank, world, 3, max(5, K // 2), profile=False)
            lw.pop("objs")
            ow = {"assembly": {k: lw[k] for k in ("n", "cells", "ms_per_step", "value", "cells_local_incl_ghosts", "dofs_local")}}
            del lw
            try:
                ow["shallow_water"] = swe_leg(gg, ctx, torch, stream, e0, e1, n=no, steps=6, dist=dist, rank=rank, world=world)
            except Exception as ex:
                ow["shallow_water"] = {"error": str(ex)}
            try:
                # the same leg on the opt-in panel-strip partition (every rank keeps the six congruent cells of a geometry class)
                ss = swe_leg(gg, ctx, torch, stream, e0, e1, n=no, steps=6, dist=dist, rank=rank, world=world, partition="strips")
                ow["shallow_water_panel_strips"] = {k: ss[k] for k in ("ms_per_step", "steps_per_s", "cells_local_incl_ghosts",
                                                                        "pcg_iters_per_solve", "pcg_us_per_iter") if k in ss}
            except Exception as ex:
                ow["shallow_water_panel_strips"] = {"error": str(ex)}
            line[other] = ow
        except Exception as ex:
            line[other] = {"error": str(ex)}

    # ---- CPU baseline beside it (rank 0, N=1) -------------------------------------------------------------------
    if not args.no_extras and rank == 0 and world == 1:
        v1, c1, s1 = cpu_baseline_run(1, 8.0)
        vN, cN, sN = cpu_baseline_run(0, 12.0)
        line["cpu_baseline"] = {"value": vN, "unit": "cells/s", "cores": cN, "kind": "port", "sample": sN,
                                "serial": {"value": v1, "cores": 1, "sample": s1}}
        try:
            line["cpu_baseline"]["rk"] = cpu_rk_baseline()
        except Exception as ex:
            line["cpu_baseline"]["rk"] = {"error": str(ex)[-300:]}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
