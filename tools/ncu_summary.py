#!/usr/bin/env python
"""Text summary of one `ncu --set full` capture of step_kernel (the metrics profiles/README.md quotes) and, with
--constants ROBOT ENVS, the per-env-step figures bench.py's roofline reads from profiles/ncu_constants.json.

usage: python tools/ncu_summary.py report.ncu-rep "header comment" [--constants UnitreeA1 4096] > profiles/rNN_..._ncu_summary.txt
"""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "gcc__average_cache_request_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__cycles_elapsed.max", "sm__icc_request_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed")


def main():
    rep, comment = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units, vals = rows[0], rows[1], rows[-1]
    d = {n: (u, v) for n, u, v in zip(names, units, vals)}
    print("# " + comment)
    for n in sorted(d):
        if n in KEEP or (n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio")):
            print("%-90s %-16s %s" % (n, d[n][0], d[n][1]))
    if "--constants" in sys.argv:
        i = sys.argv.index("--constants")
        robot, envs = sys.argv[i + 1], int(sys.argv[i + 2])
        f = lambda k: float(d[k][1].replace(",", ""))
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
        dram = sum(f(k) * scale[d[k][0]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        cyc = f("sm__cycles_elapsed.max")
        flops = cyc * (2 * f("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed") +
                       f("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed") +
                       f("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed"))
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_constants.json")
        c = json.load(open(path)) if os.path.exists(path) else {}
        c[robot] = {"dram_bytes_per_env_step": dram / envs, "fp32_flops_per_env_step": flops / envs,
                    "capture": os.path.basename(rep), "kernel_ms": f("gpu__time_duration.sum") * (1e-6 if d["gpu__time_duration.sum"][0] == "ns" else (1e-3 if d["gpu__time_duration.sum"][0] == "us" else 1.0))}
        json.dump(c, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
