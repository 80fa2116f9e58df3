#!/usr/bin/env python
"""Digest of `ncu --set full` reports: the handful of metrics DESIGN.md quotes, per captured launch.
    python tools/ncu_digest.py title1=rep1.ncu-rep title2=rep2.ncu-rep ... > profiles/<round>_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("sm__cycles_elapsed.avg.per_second", "SM clock"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active % of elapsed"),
        ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active % of elapsed"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active (realtime, triage) %"),
        ("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "bf16 tensor ops % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers/thread"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("smsp__inst_executed.sum", "warp instructions")]
print("ncu --set full --clock-control none, one attack step of the bench workload (tools/profile_step.py); per launch.")
for arg in sys.argv[1:]:
    title, rep = arg.split("=", 1)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"\n== {title}")
    for r in rows[2:]:
        print(f"  {r[hdr.index('Kernel Name')][:70]}  grid {r[hdr.index('Grid Size')] if 'Grid Size' in hdr else ''}")
        for key, label in WANT:
            cols = [i for i, h in enumerate(hdr) if h == key]
            if cols:
                print(f"      {label:36s} {r[cols[0]]} {units[cols[0]]}")
