#!/usr/bin/env python
"""Summarises one `ncu --set full --import-source on` capture (a .ncu-rep read back with `ncu -i`): headline metrics
from the raw page, and the CUDA source lines with the most warp-stall samples from the `cuda,sass` source page.
Usage: summarise_ncu_full.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys
from collections import Counter


def page(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, top=14):
    raw = page(rep, "--page", "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
            "smsp__inst_executed_op_shared_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    print("| metric | value |\n|---|---|")
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                print("| %s | %s %s |" % (w, vals[i], units[i]))
    stalls = Counter()
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                stalls[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] += float(vals[i])
            except ValueError:
                pass
    tot = sum(stalls.values()) or 1.0
    print("\nWarp-stall samples: " + ", ".join("%s %.0f%%" % (k, 100 * v / tot) for k, v in stalls.most_common(7)))
    src = page(rep, "--page", "source", "--print-source", "cuda,sass")
    fname, hdr2, agg, text = None, None, Counter(), {}
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif len(r) > 2 and r[0] == "Line No":
            hdr2 = r
            i_s = hdr2.index("# Samples")
        elif hdr2 and len(r) > i_s and r[0].strip().isdigit():
            try:
                n = int(r[i_s])
            except ValueError:
                continue
            key = (fname, int(r[0]))
            agg[key] += n
            text[key] = r[1].strip()
    total = sum(agg.values()) or 1
    print("\n| share | line | source |\n|---|---|---|")
    for (f, ln), n in agg.most_common(top):
        print("| %.1f%% | `%s:%d` | `%s` |" % (100.0 * n / total, f, ln, text[(f, ln)][:90].replace("|", "\\|")))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14)
