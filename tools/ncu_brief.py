#!/usr/bin/env python
"""Condensed view of an .ncu-rep: python tools/ncu_brief.py gpurun_out/x.ncu-rep"""
import csv
import subprocess
import sys

raw = subprocess.check_output(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], text=True)
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[0], rows[2:]
W = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "us"), ("launch__registers_per_thread", "regs"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
     ("smsp__inst_executed.sum", "winst"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
     ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
     ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf"),
     ("lts__t_bytes.sum", "l2bytes"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr")]
ST = ["long_scoreboard", "short_scoreboard", "barrier", "wait", "not_selected", "math_pipe_throttle", "mio_throttle", "lg_throttle", "selected",
      "no_instructions", "dispatch_stall", "membar"]
for r in data:
    out = []
    for k, n in W:
        if k in hdr:
            v = r[hdr.index(k)]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                v = v.replace("void ", "").split("(")[0]
            out.append(f"{n}={v}")
    st = []
    for s in ST:
        k = "smsp__pcsamp_warps_issue_stalled_" + s
        if k in hdr:
            st.append(f"{s}:{r[hdr.index(k)]}")
    print(" ".join(out))
    print("    stalls " + " ".join(st))
