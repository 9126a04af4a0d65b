#!/usr/bin/env python
"""Top stall sites of one kernel launch of an .ncu-rep (source page: needs -lineinfo at compile time and --import-source on).

    python tools/ncu_source_top.py gpurun_out/prof.ncu-rep --launch-skip 4 [--top 25] [--sass]

Prints, per source line (or per SASS instruction with --sass), the warp-stall samples and the dominant stall reasons."""
import argparse
import csv
import io
import subprocess
import sys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--launch-skip", type=int, default=0)
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--sass", action="store_true")
    a = ap.parse_args()
    cmd = ["ncu", "-i", a.rep, "--page", "source", "--csv", "--launch-skip", str(a.launch_skip), "--launch-count", "1"]
    if not a.sass:
        cmd += ["--print-source", "cuda,sass"] if False else []
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][1][:150])
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    total = 0
    recs = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            n = int(r[idx["# Samples"]])
        except ValueError:
            continue
        total += n
        st = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        recs.append((n, r[idx["Address"]], r[idx["Source"]][:110], st, r[idx["Instructions Executed"]]))
    recs.sort(reverse=True)
    print(f"total samples {total}")
    for n, addr, src, st, ie in recs[: a.top]:
        print(f"{n:6d} {100 * n / max(total, 1):5.1f}%  {addr[-6:]}  {src:110s} inst={ie:>8s}  " + " ".join(f"{c}:{v}" for v, c in st if v))


if __name__ == "__main__":
    main()
