#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel, --set full --import-source on) into markdown for profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r01_x.md [traffic.json key]
Reads the report with `ncu -i ... --page raw|source --csv`; prints headline metrics, DRAM traffic per launch,
the stall-reason split and the hottest source lines (by stall samples)."""
import csv
import io
import json
import os
import subprocess
import sys


KFILTER = []   # ["--kernel-name", "regex:..."] when the report holds several kernels (B2C_NCU_KERNEL=<regex>)
if os.environ.get("B2C_NCU_KERNEL"):
    KFILTER = ["--kernel-name", "regex:" + os.environ["B2C_NCU_KERNEL"]]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *KFILTER, *args, "--csv"], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw"))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}

    def g(name):
        v = m.get(name, ("", ""))[0].replace(",", "")
        try:
            return float(v)
        except ValueError:
            return None
    name = m.get("Kernel Name", ("?",))[0]
    keys = [("gpu__time_duration.sum", "duration"), ("sm__inst_executed.sum", "warp instructions"),
            ("smsp__inst_executed.avg.per_cycle_active", "IPC per SMSP (active)"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
            ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
            ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
            ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__waves_per_multiprocessor", "waves/SM")]
    lines = [f"# ncu summary: `{name}`", "", f"source report: `{os.path.basename(rep)}` (`ncu --set full --clock-control none --import-source on`, one launch)", "",
             "| metric | value |", "|---|---|"]
    for k, label in keys:
        if k in m:
            lines.append(f"| {label} (`{k}`) | {m[k][0]} {m[k][1]} |")
    rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")

    def tobytes(v, u):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        return v * mult.get(u, 1)
    traffic = None
    if rd is not None and wr is not None:
        traffic = tobytes(rd, m["dram__bytes_read.sum"][1]) + tobytes(wr, m["dram__bytes_write.sum"][1])
        lines.append(f"| DRAM traffic per launch | {traffic / 1e9:.3f} GB |")
    # stall reasons (warp states, per issue-slot sampling)
    stalls = [(h, g(h)) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    stalls = sorted([(h, v) for h, v in stalls if v], key=lambda t: -t[1])[:8]
    if stalls:
        lines += ["", "## warp stall reasons (warps stalled per issue-active cycle)", "", "| reason | ratio |", "|---|---|"]
        for h, v in stalls:
            lines.append(f"| {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} | {v:.2f} |")
    # hottest source lines
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--print-source", "cuda,sass"))))
    cur, h2, agg, tot_i, tot_s = None, None, [], 0, 0
    for r in src:
        if len(r) == 2 and r[0] == "File Path":
            cur = os.path.basename(r[1]); continue
        if len(r) > 5 and r[0] == "Line No":
            h2 = r; continue
        if h2 is None or len(r) != len(h2) or r[2] != "-":
            continue
        ie, sm = int(r[7] or 0), int(r[6] or 0)
        agg.append((sm, ie, cur, r[0], r[1].strip()[:110]))
        tot_i += ie; tot_s += sm
    if agg:
        agg.sort(reverse=True)
        lines += ["", f"## hottest source lines (of {tot_s} stall samples, {tot_i} warp instructions)", "",
                  "| samples % | instr % | line | source |", "|---|---|---|---|"]
        for sm, ie, f, ln, text in agg[:25]:
            lines.append(f"| {100 * sm / max(tot_s, 1):.1f} | {100 * ie / max(tot_i, 1):.1f} | {f}:{ln} | `{text.replace('|', '¦')}` |")
    open(out, "w").write("\n".join(lines) + "\n")
    if len(sys.argv) > 3 and traffic is not None:
        tj = os.path.join(os.path.dirname(out), "traffic.json")
        d = json.load(open(tj)) if os.path.exists(tj) else {}
        d[sys.argv[3]] = traffic
        json.dump(d, open(tj, "w"), indent=1, sort_keys=True)
    print("wrote", out, "traffic", traffic)


if __name__ == "__main__":
    main()
