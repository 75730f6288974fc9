#!/usr/bin/env python
"""zkstark_b200/ncu_counts.json from an `ncu --set full` capture of the Merkle leaf pass (tree_pass_kernel<values> over
2^23 leaves, the first tree_pass_kernel launch of tools/profile_prove.py 20 1):
    python tools/ncu_counts.py gpurun_out/r2_leafpass.ncu-rep
Dynamic thread-instructions per leaf, the ALU-pipe share of them, DRAM bytes per launch -- and the sha256 of the sources
the kernel is built from, so that bench.py can tell when the counts no longer describe the code it is timing."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["zkstark_b200/csrc/sha256.cuh", "zkstark_b200/csrc/tree.cu", "zkstark_b200/csrc/field.cuh", "zkstark_b200/csrc/common.cuh"]


def source_hash():
    h = hashlib.sha256()
    for s in SOURCES:
        h.update(open(os.path.join(ROOT, s), "rb").read())
    return h.hexdigest()


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, body = rows[0], rows[2:]
    col = lambda r, k: float(r[hdr.index(k)].replace(",", ""))
    import re
    r = next(r for r in body if re.search(r"tree_pass_kernel<(\(int\))?0,", r[hdr.index("Kernel Name")]))   # SRC_VALUES = 0
    grid = int(r[hdr.index("Grid Size")].strip("()").split(",")[0])
    leaves = 1 << 23
    inst = col(r, "smsp__inst_executed.sum") * 32 / leaves
    alu_pct = col(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")
    issue_pct = col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    out = {
        "kernel": r[hdr.index("Kernel Name")].split("(")[0].strip() + f", grid {grid} x 256 threads, 2^23 leaves, production launch shape (CTA epilogue included)",
        "inst_dyn_per_leaf": inst,
        "alu_dyn_per_leaf": inst * (alu_pct / 100.0 * 0.5) / (issue_pct / 100.0),
        "alu_pipe_pct": alu_pct, "issue_active_pct": issue_pct,
        "fma_pipe_pct": col(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "us_under_ncu": col(r, "gpu__time_duration.sum") if "gpu__time_duration.sum" in hdr else None,
        "dram_bytes_per_launch": (col(r, "dram__bytes_read.sum") + col(r, "dram__bytes_write.sum")) * 1e6
        if rows[1][hdr.index("dram__bytes_read.sum")].lower().startswith("mbyte") else col(r, "dram__bytes_read.sum") + col(r, "dram__bytes_write.sum"),
        "source_sha256": source_hash(),
        "source": f"{os.path.basename(sys.argv[1])}: smsp__inst_executed.sum*32/leaves thread-instructions per leaf; ALU-pipe instructions = "
                  "that x (sm__inst_executed_pipe_alu % x 0.5 inst/cycle/SMSP) / (issue active %); dram = dram__bytes_read.sum + "
                  "dram__bytes_write.sum of the launch.  Regenerate with tools/ncu_counts.py when the sources change.",
    }
    json.dump(out, open(os.path.join(ROOT, "zkstark_b200", "ncu_counts.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
