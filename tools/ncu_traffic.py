#!/usr/bin/env python
"""DRAM and L2 traffic per launch of the stream / transform kernels bench.py reports in `roofline_hbm_kernels`.

  # on the GPU box, one op per ncu run (the op runs twice: a warm-up and the measured pass, an L2 flush in front of each)
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none \\
      --csv --log-file gpurun_out/traffic_<which>_<log>.csv python tools/ncu_traffic.py run <which> <log>
  # here
  python tools/ncu_traffic.py collect gpurun_out/traffic_*.csv      -> zkstark_b200/ncu_traffic.json

`which` is the zk_bench_kernel id (2 LDE, 3 interpolation, 4 FRI fold, 5 composition polynomial, 1 whole Merkle tree).  The
json carries the sha256 of the kernel sources, so bench.py can tell when the numbers no longer describe the code it times."""
import csv
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["zkstark_b200/csrc/ntt.cu", "zkstark_b200/csrc/tree.cu", "zkstark_b200/csrc/field.cuh", "zkstark_b200/csrc/common.cuh",
           "zkstark_b200/csrc/sha256.cuh"]
NAMES = {2: "lde (all passes)", 3: "interpolate (all passes)", 4: "fri_fold layer0", 5: "cp_eval", 1: "merkle tree (all passes)"}


def source_hash():
    h = hashlib.sha256()
    for s in SOURCES:
        h.update(open(os.path.join(ROOT, s), "rb").read())
    return h.hexdigest()


def run(which, log):
    sys.path.insert(0, ROOT)
    import zkstark_b200 as zk
    ctx = zk.Context(0)
    ms, _, _ = ctx.bench_kernel(which, log, iters=1, flush_l2=True)
    print(f"which {which} 2^{log}: {ms * 1e3:.1f} us")


def collect(paths):
    out = {"source_sha256": source_hash(), "kernels": {},
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum "
                     "--clock-control none python tools/ncu_traffic.py run <which> <log>: the launches behind the last L2 flush"}
    for p in sorted(paths):
        m = re.search(r"traffic_(\d+)_(\d+)\.csv$", p)
        which, log = int(m.group(1)), int(m.group(2))
        rows = [r for r in csv.reader(l for l in open(p) if l.startswith('"'))]
        hdr = rows[0]
        ki, mi, ui, vi, idi = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
        launches = {}
        for r in rows[1:]:
            d = launches.setdefault(int(r[idi]), {"name": r[ki]})
            v = float(r[vi].replace(",", ""))
            u = r[ui].lower()
            scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3,
                     "msecond": 1e3}.get(u, 1)
            d[r[mi]] = v * scale
        ids = sorted(launches)
        last_flush = max(i for i in ids if "flush_kernel" in launches[i]["name"])
        op = [launches[i] for i in ids if i > last_flush and "fill_kernel" not in launches[i]["name"]]
        name = NAMES[which] + ("" if log in (20, 23) else f" @2^{log}")
        out["kernels"][f"{which}:{log}"] = {
            "name": name, "launches": [l["name"].split("(")[0].strip() for l in op],
            "dram_bytes": sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in op),
            "l2_bytes": sum(l["lts__t_bytes.sum"] for l in op),
            "us_under_ncu": sum(l["gpu__time_duration.sum"] for l in op)}
    json.dump(out, open(os.path.join(ROOT, "zkstark_b200", "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(int(sys.argv[2]), int(sys.argv[3]))
    else:
        collect(sys.argv[2:])
