"""Static SASS instruction mix of the hot kernels in libzkstark_b200.so (cuobjdump -sass).

Writes zkstark_b200/sass_counts.json.  bench.py uses `alu_per_leaf` -- the number of ALU-pipe
instructions (SHF/LOP3/IADD3/VIADD/PRMT/ISETP/SEL/...) one thread of tree_pass_kernel<values> issues
per leaf, i.e. one leaf hash + one node hash (the loop body; a 2^k-leaf thread runs 2^k leaf hashes and
2^k - 1 node hashes) -- to turn a measured launch time into achieved integer-pipe throughput.
IMAD* goes to the FMA pipe and is counted separately.
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "zkstark_b200", "libzkstark_b200.so")
OUT = os.path.join(ROOT, "zkstark_b200", "sass_counts.json")

ALU = {"SHF", "LOP3", "IADD3", "VIADD", "PRMT", "ISETP", "SEL", "LEA", "IADD", "LOP", "SGXT", "BMSK", "VIMNMX", "IABS", "FLO", "POPC", "BREV", "MOV", "PLOP3", "CS2R"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2"}


def kernel_mix(sass, mangled_substr):
    mix = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and mangled_substr in cur:
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                mix[m.group(1).split(".")[0]] += 1
    return mix


def main():
    sass = subprocess.check_output(["cuobjdump", "-sass", SO], text=True)
    out = {}
    for name, key in [("tree_pass_values", "tree_pass_kernelILi0E"), ("tree_pass_digests", "tree_pass_kernelILi1E"),
                      ("tree_pass_fold", "tree_pass_kernelILi2E"), ("tree_pass_cp", "tree_pass_kernelILi3E"),
                      ("ntt_strided", "ntt_strided_kernel"), ("ntt_final", "ntt_final_kernel"),
                      ("fold", "fold_kernel"), ("cp", "9cp_kernel")]:
        mix = kernel_mix(sass, key)
        out[name] = {"total": sum(mix.values()), "alu": sum(v for k, v in mix.items() if k in ALU),
                     "fma": sum(v for k, v in mix.items() if k in FMA), "mix": dict(mix.most_common(12))}
    out["alu_per_leaf"] = out["tree_pass_values"]["alu"]
    out["fma_per_leaf"] = out["tree_pass_values"]["fma"]
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps({k: (v if not isinstance(v, dict) else {kk: v[kk] for kk in ("total", "alu", "fma")}) for k, v in out.items()}))


if __name__ == "__main__":
    main()
