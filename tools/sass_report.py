#!/usr/bin/env python
"""Static evidence for profiles/: ptxas resource table (registers, spills, shared memory) of every kernel from the build logs and
SASS opcode histograms of the hot kernels from the built library (cuobjdump -sass).  No GPU needed.

    python tools/sass_report.py > profiles/sass_ptxas_r02.txt
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gridapgeosciences.jl_b200")
HOT = ["k_lb_tileILi3ELi10ELb0ELb0", "k_lb_tile_groupsILi9", "k_lb_fastILi3ELi10ELb0ELb0", "k_gather_chunks", "k_pcg_ebeILi12ELi12ELb1",
       "k_swe_diag_lanesILi2ELi7", "k_swe_stage_lanesILi2ELb1ELi7", "k_swe_stage_sfILi2ELi1ELb0", "k_swe_diagILi2ELi1", "k_ebe_scalar_massILi3", "k_bq_residualILi36ELi8", "k_hex_tensor_fill"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    except Exception:
        return name


def ptxas_table():
    rows = []
    for log in sorted(glob.glob(os.path.join(PKG, "build", "*.ptxas.log"))):
        txt = open(log).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                             r"ptxas info\s+: Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", txt):
            rows.append((os.path.basename(log).replace(".ptxas.log", ""), m.group(1), int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4)),
                         int(m.group(8) or 0)))
    return rows


def sass_histogram(obj, pattern):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    hists = {}
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1) if pattern in m.group(1) else None
            if cur:
                hists[cur] = collections.Counter()
            continue
        if cur:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                hists[cur][m.group(1).split(".")[0]] += 1
    return hists


def main():
    print("# ptxas resource usage (sm_100a, -O3), kernels with >= 64 registers or any spill; full logs: gridapgeosciences.jl_b200/build/*.ptxas.log")
    print(f"{'unit':<12s} {'regs':>5s} {'stack':>6s} {'spill st':>9s} {'spill ld':>9s} {'smem':>6s}  kernel")
    for unit, name, regs, stack, sst, sld, smem in ptxas_table():
        if regs >= 64 or sst or sld:
            print(f"{unit:<12s} {regs:>5d} {stack:>6d} {sst:>9d} {sld:>9d} {smem:>6d}  {demangle(name)[:150]}")
    print()
    print("# SASS opcode histograms of the hot kernels (static instruction counts; cuobjdump -sass on the objects of libgeosgpu.so)")
    print("# FP64 work is DFMA / DMUL / DADD (tcgen05 has no FP64 path: no UTC*MMA expected); MUFU.RSQ64H is the metric's rsqrt seed")
    objs = sorted(glob.glob(os.path.join(PKG, "build", "*.o")))
    for pat in HOT:
        for obj in objs:
            for fn, h in sass_histogram(obj, pat).items():
                tot = sum(h.values())
                top = ", ".join(f"{k} {v}" for k, v in h.most_common(14))
                print(f"{demangle(fn)[:140]}\n    {tot} instructions: {top}")
                local = h.get("LDL", 0) + h.get("STL", 0)
                print(f"    local-memory instructions (LDL + STL): {local};  tensor-core / TMA opcodes (UTC*MMA, UTMALDG, HMMA): "
                      f"{sum(v for k, v in h.items() if k.startswith(('UTC', 'UTMA', 'HMMA')))}")


if __name__ == "__main__":
    main()
