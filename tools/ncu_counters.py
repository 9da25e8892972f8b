#!/usr/bin/env python
"""Extract the counters bench.py quotes beside its live timing from an `ncu --set full` report and merge them into
profiles/kernel_counters.json under a workload key.

    python tools/ncu_counters.py gpurun_out/r2c_prof.ncu-rep AntUMaze-v0:65536 ["note"]

Per launch: warp instructions, fp32 operations (FADD + FMUL + 2 FFMA, thread level), DRAM bytes, shared-memory
wavefronts, and the utilisations ncu derived from them (issue slots, FMA pipe, LSU data pipe) with the stall split."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, key = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rows = list(csv.reader(raw.splitlines()))
    names, units, vals = rows[0], rows[1], rows[2]
    m = {n: (v, u) for n, u, v in zip(names, units, vals)}

    def f(name, scale=None):
        v, u = m[name]
        v = float(v.replace(",", ""))
        mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}.get(u, 1.0)
        return v * mult

    cycles = f("sm__cycles_elapsed.max")
    flop_per_cycle = (f("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed")
                      + f("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed")
                      + 2 * f("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed"))
    stalls = {n.split("issue_stalled_")[1].split("_per_issue")[0]: float(v) for n, (v, u) in m.items()
              if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("per_issue_active.ratio")}
    tot = sum(stalls.values()) or 1.0
    out = {
        "source": os.path.basename(rep), "kernel": m["Kernel Name"][0], "note": note,
        "captured_ms": f("gpu__time_duration.sum") * 1e3, "sm_cycles": cycles,
        "warp_inst": f("smsp__inst_executed.sum"),
        "fp32_flop": flop_per_cycle * f("sm__cycles_elapsed.avg"),
        "dram_bytes": f("dram__bytes_read.sum") + f("dram__bytes_write.sum"),
        "smem_wavefronts": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "issue_active_pct": f("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        "fma_pipe_pct": f("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "lsu_data_pipe_pct": f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "warps_per_sm": f("sm__warps_active.avg.per_cycle_active"),
        "registers": int(f("launch__registers_per_thread")),
        "stall_share_pct": {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v / tot > 0.005},
        "cache_control": "none (--cache-control none is not set: ncu flushes caches before each replay pass, so dram_bytes is a COLD-cache figure like the bench's flushed L2)",
    }
    path = os.path.join(ROOT, "profiles", "kernel_counters.json")
    db = json.load(open(path)) if os.path.exists(path) else {}
    db[key] = out
    json.dump(db, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
