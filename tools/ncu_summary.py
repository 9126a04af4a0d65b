#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report (read here, without a GPU): duration, DRAM bytes, pipe utilisation.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--md profiles/x.md] [--json profiles/x.json]

Launches of the same kernel with the same grid are averaged (the first one of each group is the cold-cache one)."""
import argparse
import csv
import io
import json
import subprocess

METRICS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "sm__cycles_elapsed.max": "cycles",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
}


def to_float(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    if u in ("ms", "msecond"):
        return v * 1e3
    if u in ("ns", "nsecond"):
        return v / 1e3
    if u == "gbyte":
        return v * 1e3
    if u == "kbyte":
        return v / 1e3
    if u == "byte":
        return v / 1e6
    return v


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    recs = []
    for r in rows[2:]:
        rec = {"kernel": r[idx["Kernel Name"]]}
        for m, name in METRICS.items():
            if m in idx and r[idx[m]] not in ("", "n/a"):
                try:
                    rec[name] = to_float(r[idx[m]], units[idx[m]])
                except ValueError:
                    pass
        recs.append(rec)
    return recs


def short(name):
    n = name.replace("void ", "").replace("<unnamed>::", "")
    return n.split("(CUtensorMap")[0].split("(const")[0].strip()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--md", default=None)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    recs = load(a.rep)
    groups = []
    for r in recs:  # consecutive launches of the same kernel with similar duration = one case
        k = short(r["kernel"])
        if groups and groups[-1]["kernel"] == k and abs(groups[-1]["launches"][-1]["time_us"] - r["time_us"]) < 0.25 * r["time_us"]:
            groups[-1]["launches"].append(r)
        else:
            groups.append({"kernel": k, "launches": [r]})
    summary = []
    for g in groups:
        ls = g["launches"]
        s = {"kernel": g["kernel"], "launches": len(ls)}
        for name in METRICS.values():
            vals = [l[name] for l in ls if name in l]
            if vals:
                s[name] = round(sum(vals) / len(vals), 2)
        summary.append(s)
    lines = ["| kernel | n | time us | DRAM rd MB | DRAM wr MB | tensor % | XU % | DRAM % | issue % | regs | grid |", "|---|---|---|---|---|---|---|---|---|---|---|"]
    for s in summary:
        lines.append("| {kernel} | {launches} | {time_us} | {dram_read_MB} | {dram_write_MB} | {tensor_pipe_pct} | {xu_pipe_pct} | {dram_throughput_pct} | {issue} | {registers:.0f} | {grid:.0f} |".format(
            issue=s.get("issue_active_pct", "-"), **{k: s.get(k, "-") for k in ("kernel", "launches", "time_us", "dram_read_MB", "dram_write_MB", "tensor_pipe_pct", "xu_pipe_pct", "dram_throughput_pct")},
            registers=s.get("registers", 0), grid=s.get("grid", 0)))
    print("\n".join(lines))
    if a.md:
        with open(a.md, "a") as f:
            f.write("\n".join(lines) + "\n")
    if a.json:
        json.dump(summary, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
