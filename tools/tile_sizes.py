#!/usr/bin/env python
"""Tile path against the two-kernel path of the Laplace-Beltrami assembly over mesh sizes (development tool).

    GG_LB_TILES=0|1 python tools/tile_sizes.py 128 181 256
Prints ms per assembly (CUDA events on the library's stream, L2 flushed between steps) for n x n cells per panel.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402


def main():
    import torch
    gg = load_package()
    ctx = gg.get_context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=0)
    out = {}
    for n in [int(a) for a in sys.argv[1:]]:
        model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=n, ctx=ctx)
        dO = gg.Measure(gg.Triangulation(model), bench.DEGREE)
        V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, bench.ORDER), conformity="H1")
        assem = gg.SparseMatrixAssembler(V, V)
        u = gg.DeviceVector.from_numpy(ctx, np.random.default_rng(0).standard_normal(V.num_free_dofs))
        b = gg.DeviceVector(ctx, V.num_free_dofs)
        form = gg.LaplaceBeltrami(dO, u=u)
        step = lambda: gg.assemble_matrix_and_vector(form, assem, b)  # noqa: E731
        flush = bench.L2Flush(torch, stream, ctx.device)
        ctx.set_async(True)
        for _ in range(5):
            step()
        ctx.synchronize()
        torch.cuda.synchronize()
        ms = bench.timed_steps(torch, stream, step, 30, flush)
        ctx.set_async(False)
        out[n] = {"cells": model.num_cells, "tiles": (model.num_cells + 127) // 128, "ms": ms}
    print(json.dumps({"GG_LB_TILES": os.environ.get("GG_LB_TILES"), "sizes": out}))


if __name__ == "__main__":
    main()
