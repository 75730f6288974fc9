"""GPU experiment: integer throughput of the ALU pipe, the FMA pipe (IMAD) and a 1:1 mix (zk_bench_kernel(6, mode))."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zkstark_b200 as zk
ctx = zk.Context(0)
for mode, name in ((0, "ALU  SHF/LOP3/IADD3"), (1, "FMA  IMAD"), (2, "1:1 ALU/IMAD mix")):
    _, _, peak = ctx.bench_kernel(6, mode, iters=5, flush_l2=False)
    print(f"{name:22s} {peak:7.2f} Tiop/s  = {peak*1e12/148/1.965e9:6.1f} thread-instr/clk/SM")
