"""Iterations and time per step of the shallow-water stepper for the initial-guess variants of the mass solves
(GG_EXTRAPOLATE = history levels; 0 = plain warm start).  One process per variant (the setting is read at stepper creation)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import torch
    import bench
    from __graft_entry__ import load_package
    gg = load_package()
    ctx = gg.get_context(0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = bench.swe_leg(gg, ctx, torch, stream, e0, e1, steps=20, state=sys.argv[2])
    print("RESULT " + json.dumps({k: out[k] for k in ("ms_per_step", "pcg_iters_per_solve", "state_drift_after_steps")}))
else:
    for state in ("galewsky",):
        for lv in ("0", "3", "4", "5", "6"):
            env = dict(os.environ, GG_EXTRAPOLATE=lv)
            r = subprocess.run([sys.executable, __file__, "--one", state], env=env, capture_output=True, text=True)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            print(state, "levels", lv, line[0] if line else r.stderr[-400:], flush=True)
