#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the handful of numbers the roofline argument needs.

    python profiles/summarize_ncu.py gpurun_out/x.ncu-rep [algorithmic_bytes] [algorithmic_flops] > profiles/x.json

Runs `ncu -i <rep> --page raw --csv` (no GPU needed) and keeps: duration, SM clock, DRAM bytes read / written (the
`traffic` of bench.py's roofline object), DRAM and L2 throughput %, tensor-pipe active %, XU (MUFU) %, issue-slot
utilisation, shared-memory wavefront share of the tensor core, registers, and -- when algorithmic bytes / flops are
given -- achieved GB/s / TFLOP/s under the profiler (cold cache, one launch: compare shares, not absolutes)."""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "sm__cycles_elapsed.max": "sm_cycles",
    "gpc__cycles_elapsed.avg.per_second": "sm_clock",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed": "xu_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_elapsed": "fma_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_elapsed": "alu_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_slots_pct",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_tensor_wavefronts_pct",
    "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed": "smem_bank_reads_pct",
    "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed": "smem_bank_writes_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "ms": 1e-3, "us": 1e-6, "s": 1.0,
         "ns": 1e-9, "Ghz": 1e9, "Mhz": 1e6}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None}
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP and v != "":
                x = float(v.replace(",", ""))
                d[KEEP[h]] = x * SCALE[u] if u in SCALE else x
        if "dram_read" in d and "dram_write" in d:
            d["dram_bytes_per_launch"] = d["dram_read"] + d["dram_write"]
        if len(sys.argv) > 2 and float(sys.argv[2]) > 0:
            d["algorithmic_bytes"] = float(sys.argv[2])
            d["achieved_gbs_under_ncu"] = d["algorithmic_bytes"] / d["duration"] / 1e9
            d["traffic_over_algorithmic"] = d["dram_bytes_per_launch"] / d["algorithmic_bytes"]
        if len(sys.argv) > 3 and float(sys.argv[3]) > 0:
            d["algorithmic_flops"] = float(sys.argv[3])
            d["achieved_tflops_under_ncu"] = d["algorithmic_flops"] / d["duration"] / 1e12
        res.append(d)
    json.dump({"report": rep, "note": "one launch under ncu --set full --clock-control none: cold cache, serialised",
               "launches": res}, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
