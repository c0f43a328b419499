#!/usr/bin/env python
"""Per-kernel CUDA-event breakdown of one shallow-water SSPRK3 step (development tool; prints JSON).

    python tools/swe_profile.py --n 256 --order 2 --steps 3
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

OMEGA, G, H0, U0 = 1.0, 9.8 * (1 / 7.29e-5) ** 2 / 6.37e6, 2.94e4 / 9.8 / 6.37e6, 2 * np.pi / (12 * 24 * 3600 * 7.29e-5)


def coriolis(X):
    return 2.0 * OMEGA * X[..., 2] / np.linalg.norm(X, axis=-1)


def depth(X):
    s = X[..., 2] / np.linalg.norm(X, axis=-1)
    return H0 - (OMEGA * U0 + 0.5 * U0 * U0) * s * s / G


def velocity(X):
    r = np.linalg.norm(X, axis=-1, keepdims=True)
    n = X / r
    # solid-body rotation about z: u0 cos(phi) e_theta = u0 (-y, x, 0)/r
    return U0 * np.stack([-n[..., 1], n[..., 0], 0 * n[..., 0]], -1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--order", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--rtol", type=float, default=1e-12)
    ap.add_argument("--thermal", action="store_true", help="thermal shallow water (adds the buoyancy field and its skeleton terms)")
    args = ap.parse_args()
    gg = load_package()
    ctx = gg.get_context(0)
    t0 = time.perf_counter()
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=args.n, ctx=ctx)
    p = args.order
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.raviart_thomas, float, p), conformity="HDiv")
    Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="L2")
    x0 = np.concatenate([gg.interpolate(velocity, V).get(), gg.interpolate(depth, Q).get()])
    dt = gg.dx(model) * 0.1 / (max(p, 1) * np.sqrt(G * H0))
    op = gg.ShallowWaterOperator(model, p, coriolis, gravity=G)
    if args.thermal:
        ic = gg.initial_conditions
        x0 = np.concatenate([gg.interpolate(ic.williamson2_velocity, V).get(), gg.interpolate(ic.williamson2_depth, Q).get(),
                             gg.interpolate(ic.thermogeostrophic_weighted_buoyancy, Q).get()])
        dt = gg.dx(model) * 0.1 / (max(p, 1) * np.sqrt(ic.GRAVITY * ic.H0_W2))
        op = gg.ThermalShallowWaterOperator(model, p, ic.coriolis)
    st = gg.RKStepper(gg.RungeKutta(gg.CGSolver(rtol=args.rtol), None, dt, "EXRK_SSP_3_3"), op)
    st.set_state(x0)
    setup_s = time.perf_counter() - t0
    st.step(1)
    t0 = time.perf_counter()
    st.step(args.steps)
    wall = (time.perf_counter() - t0) / args.steps
    it0, so0 = st.stats()
    ctx.profile(True)
    st.step(args.steps)
    names = ["k_swe_diag", "k_swe_stage", "k_scalar_matrix", "k_gather_chunks", "k_gather_dofs", "k_block_inverse", "k_pcg_persistent", "k_pcg_mf", "k_pcg_ebe", "k_ebe_condense", "k_ebe_precond", "k_ebe_scalar_mass", "k_tswe_buoyancy", "k_tswe_stage", "k_consistent", "k_extrapolate"]
    prof = {}
    for nme in names:
        cnt, ms = ctx.profile_query(nme)
        prof[nme] = {"launches_per_step": cnt / args.steps, "ms_per_step": ms / args.steps}
    ctx.profile(False)
    it1, so1 = st.stats()
    x = st.get_state()
    phases = {}
    for which, name in ((0, "rt_mass"), (1, "q_mass")):
        a = st.solver_phase_ns(which)
        if a[3] > 0:
            phases[name] = {"iters": a[3], "us_per_iter": {"cells": a[0] / a[3] / 1e3, "gather_dot": a[1] / a[3] / 1e3, "update_precond": a[2] / a[3] / 1e3}}
    setup = {name: [round(v / 1e3, 1) for v in st.solver_phase_ns(10 + which)] for which, name in ((0, "rt_mass"), (1, "q_mass"))}
    out = {"pcg_setup_us_per_solve(cell pre-pass, rows, first precond, interior recovery)": setup, "n": args.n, "order": p, "cells": model.num_cells, "dofs_prognostic": st.nu + st.np_, "dofs_diagnostic": st.ny,
           "setup_s": setup_s, "ms_per_step": wall * 1e3, "steps_per_s": 1.0 / wall,
           "pcg_iters_per_solve": (it1 - it0) / max(so1 - so0, 1), "solves_per_step": (so1 - so0) / args.steps,
           "kernels": prof, "pcg_phases": phases, "profiled_ms_per_step": sum(v["ms_per_step"] for v in prof.values()),
           "state_drift": float(np.abs(x - x0).max() / np.abs(x0).max())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
