#!/usr/bin/env python
"""Laplace-Beltrami residual + Jacobian assembly on the 3-D shell (BASELINE config 5: hexahedral Q2): cells/s, prints JSON.

    python tools/shell_bench.py --n 48 --order 2 --degree 18 --steps 10
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--degree", type=int, default=18)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    gg = load_package()
    ctx = gg.get_context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=0)
    t0 = time.perf_counter()
    model = gg.AtlasDiscreteModel(gg.CubedSphereWithThicknessMesh(1.0, 0.19), n=args.n, n_layers=args.layers or None, ctx=ctx)
    dΩ = gg.Measure(gg.Triangulation(model), args.degree)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, args.order), conformity="H1")
    t_mesh = time.perf_counter() - t0
    t0 = time.perf_counter()
    assem = gg.SparseMatrixAssembler(V, V)
    t_pattern = time.perf_counter() - t0
    A = assem.matrix
    u = gg.DeviceVector.from_numpy(ctx, np.random.default_rng(1).standard_normal(V.num_free_dofs))
    b = gg.DeviceVector(ctx, V.num_free_dofs)
    form = gg.LaplaceBeltrami(dΩ, u=u)
    gg.assemble_matrix_and_vector(form, assem, b)       # builds the gather lists
    ctx.set_async(True)
    for _ in range(3):
        gg.assemble_matrix_and_vector(form, assem, b)
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        gg.assemble_matrix_and_vector(form, assem, b)
    e1.record(stream)
    ctx.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    ctx.profile(True)
    for _ in range(args.steps):
        gg.assemble_matrix_and_vector(form, assem, b)
    ctx.synchronize()
    prof = {k: ctx.profile_query(k)[1] / args.steps for k in ("k_hex_tensor_fill", "k_hex_kron", "k_hex_residual", "k_lb_fast", "k_scalar_matrix", "k_gather_chunks", "k_gather_dofs")}
    ctx.profile(False)
    nq = (args.degree + 2) // 2
    m = (args.order + 1) ** 3
    out = {"workload": f"laplace_beltrami_hexq{args.order}_deg{args.degree}_shell_n{args.n}", "cells": model.num_cells, "dofs": V.num_free_dofs,
           "nnz": A.nnz, "points_per_cell": nq ** 3, "setup_mesh_space_s": t_mesh, "setup_pattern_s": t_pattern,
           "ms_per_step": ms, "cells_per_s": model.num_cells / (ms * 1e-3), "kernels_ms": prof,
           "bytes_written_min": 8 * (m * (m + 1) // 2 * model.num_cells + A.nnz),
           "algorithmic_flops_per_cell_survey": nq ** 3 * (m * m * 6 + m * 18 + 80)}
    out["algorithmic_tflops"] = out["algorithmic_flops_per_cell_survey"] * out["cells_per_s"] / 1e12
    print(json.dumps(out))


if __name__ == "__main__":
    main()
