#!/usr/bin/env python
"""Per-source-line stall samples of one kernel of an .ncu-rep (ncu --set full --import-source on, built with -lineinfo).

    python tools/ncu_lines.py report.ncu-rep kernel_regex [top_n]
Prints, for the source lines with the most warp-stall samples: samples, share, warp instructions executed, dominant stall reasons.
"""
import csv
import io
import subprocess
import sys


def main(rep, kern, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kern, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr, recs = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0] not in ("", "Function Name") and r[0].isdigit():
            d = dict(zip(hdr[4:], r[4:]))
            def I(k):
                try:
                    return int(d.get(k, "0"))
                except ValueError:
                    return 0
            stalls = {k[6:]: I(k) for k in d if k.startswith("stall_") and "Not Issued" not in k}
            recs.append((I("# Samples"), I("Instructions Executed"), fname, int(r[0]), r[1].strip(), stalls,
                         I("L1 Wavefronts Shared Excessive")))
    tot = sum(x[0] for x in recs) or 1
    toti = sum(x[1] for x in recs) or 1
    print(f"# {kern}: {tot} samples, {toti} warp instructions")
    for s, ins, f, ln, src, st, exc in sorted(recs, key=lambda x: -x[0])[:top]:
        why = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
        print(f"{100.0 * s / tot:5.1f}% smp {100.0 * ins / toti:5.1f}% ins  {f}:{ln:<4d} {src[:70]:<70s} | {why}" + (f" | smem excess {exc}" if exc else ""))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
