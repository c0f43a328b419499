#!/usr/bin/env python
"""SSPRK steps/s of the linear wave driver at the reference test's size (development tool): python tools/wave_profile.py --n 48"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=48)
    ap.add_argument("--steps", type=int, default=300)
    args = ap.parse_args()
    gg = load_package()
    ctx = gg.get_context(0)
    wm = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=args.n, ctx=ctx)
    st = gg.WaveStepper(gg.RungeKutta(gg.CGSolver(rtol=1e-12), None, gg.dx(wm) * 0.1, "EXRK_SSP_3_3"), gg.WaveOperator(wm, 1))
    wq = gg.TestFESpace(wm, gg.ReferenceFE(gg.lagrangian, float, 1), conformity="L2")
    x0 = np.zeros(st.nu + st.np_)
    x0[st.nu:] = gg.interpolate(lambda X: 1.0 + 0.01 * np.exp(-5.0 * ((X[..., 0] - 1.0) ** 2 + X[..., 1] ** 2 + X[..., 2] ** 2)), wq).get()
    st.set_state(x0)
    st.step(20)
    ctx.synchronize()
    t0 = time.perf_counter()
    st.step(args.steps)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    it, so = st.stats()
    try:
        a, su = st.solver_phase_ns(0), st.solver_phase_ns(10)
        print("  us per iteration (cells, gather+dot, update):", [round(v / max(a[3], 1) / 1e3, 2) for v in a[:3]],
              " us per solve (pre-pass, rows, first precond, recovery):", [round(v / 1e3, 2) for v in su])
    except Exception as ex:  # wave steppers without the condensed solver
        print("  no phase timers:", ex)
    print(f"n={args.n} grid={os.environ.get('GG_EBE_GRID', 'default')} steps/s={args.steps / dt:.1f} us/step={1e6 * dt / args.steps:.1f} iters/solve={it / max(so, 1):.2f}")


if __name__ == "__main__":
    main()
