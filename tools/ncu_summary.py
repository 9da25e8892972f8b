#!/usr/bin/env python
"""One-screen summary of an ncu report: python tools/ncu_summary.py prof.ncu-rep [top]  (needs ncu on PATH)"""
import csv
import io
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rep = sys.argv[1]
    top = sys.argv[2] if len(sys.argv) > 2 else "30"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_lsu.sum"]
    for k in keys:
        for i, h in enumerate(hdr):
            if h == k:
                print(f"{k} [{units[i]}]: {[r[i] for r in data]}")
    st = {}
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("not_issued"):
            st[h[33:]] = int(data[0][i])
    tot = sum(st.values()) or 1
    print("stall samples %:", {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]})
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    tmp = "/tmp/_ncu_src.csv"
    open(tmp, "w").write(src)
    subprocess.run([sys.executable, os.path.join(HERE, "ncu_lines.py"), tmp, top])


if __name__ == "__main__":
    main()
