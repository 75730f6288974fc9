"""One-line digest of a bench.py JSON line read from stdin (experiments)."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
b, s = d.get("stage_ms_batched_step", {}), d.get("stage_ms", {})
print(sys.argv[1] if len(sys.argv) > 1 else "", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
      "| batched interp/lde ms", round(b.get("interpolate", 0), 3), round(b.get("lde", 0), 3),
      "| single interp/lde/total ms", round(s.get("interpolate", 0), 4), round(s.get("lde", 0), 4), round(s.get("total", 0), 3),
      "| lde G/s single/batched", round(d.get("lde_elems_per_s", 0) / 1e9, 1), round(d.get("lde_elems_per_s_batched", 0) / 1e9, 1),
      {k: round(v["ms"] * 1e3, 1) for k, v in d.get("roofline_hbm_kernels", {}).items()})
