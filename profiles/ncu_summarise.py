#!/usr/bin/env python
"""Turn `ncu -i X.ncu-rep --page raw --csv` into one line per profiled launch.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
    python profiles/ncu_summarise.py raw.csv > profiles/rNN_xxx_summary.txt
"""
import csv
import re
import sys


def main(path):
    r = csv.reader(open(path))
    h = next(r)
    next(r)
    ix = {k: i for i, k in enumerate(h)}

    def f(row, k, d=0.0):
        try:
            return float(row[ix[k]].replace(",", ""))
        except Exception:
            return d

    stalls = [k for k in h if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
    for row in r:
        name = re.sub(r"void pcd::pcd_kernel<pcd::K|, pcd::\w+>\(.*", "", row[ix["Kernel Name"]])[:22]
        st = sorted(((f(row, k), k.split("stalled_")[1].split("_per_")[0]) for k in stalls), reverse=True)[:4]
        ffma = f(row, "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed")
        fadd = f(row, "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed")
        fmul = f(row, "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed")
        print(f"{name:22s} t={f(row, 'gpu__time_duration.sum'):8.1f}us grid={row[ix['Grid Size']]:>14s} "
              f"regs={f(row, 'launch__registers_per_thread'):4.0f} smem={f(row, 'launch__shared_mem_per_block'):6.1f}K "
              f"occ%={f(row, 'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} "
              f"issue%={f(row, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):5.1f} "
              f"inst={f(row, 'smsp__inst_executed.sum') / 1e6:7.2f}M "
              f"fp32/clk/SM={(ffma + fadd + fmul) / 148:6.1f} (ffma {ffma / 148:5.1f} of 128) "
              f"fma_pipe%={f(row, 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed'):5.1f} "
              f"dram={f(row, 'dram__bytes_read.sum'):7.1f}+{f(row, 'dram__bytes_write.sum'):7.1f}{u_dram(h, path)} "
              f"bankconf={f(row, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum') / 1e6:6.2f}M | "
              + ", ".join(f"{n}={v:.1f}" for v, n in st))


_U = {}


def u_dram(h, path):
    if path not in _U:
        r = csv.reader(open(path))
        hh = next(r)
        uu = next(r)
        _U[path] = uu[hh.index("dram__bytes_read.sum")]
    return _U[path]


if __name__ == "__main__":
    main(sys.argv[1])
