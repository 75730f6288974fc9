#!/usr/bin/env python
"""bench.py -- headline benchmark of the zkstark_b200 hot path (contract: see DESIGN.md "Measurement").

A "step" is `--batch` (default 32) independent prove() calls (interpolate -> LDE -> Merkle -> CP -> FRI ->
decommit; prover.rs:9-293) per rank, each of one synthetic FibonacciSq trace of BASELINE.json configs[2]:
trace 2^20, blow-up 8 (8,388,608-point LDE), secret 3141592 + global proof index.  The proofs of a step run
as one batched launch sequence (the latency-bound upper Merkle levels and Fiat-Shamir steps of one proof
overlap the others); single-proof latency is reported beside it as prove_ms_single.  Independent proofs
shard across ranks with no data-path collective (weak scaling): value = proofs/s summed over all ranks.

  python bench.py                          # 1 GPU
  torchrun --nproc-per-node N bench.py --gpus N
  python bench.py --impl reference         # the CPU arm: the oracle's NTT route on all host threads
                                           # (the reference's own Rust code cannot be built here and is O(n^3))
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "proofs_per_s"
UNIT = "proofs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=20, help="log2 of the trace domain (BASELINE configs[2] = 20)")
    ap.add_argument("--batch", type=int, default=32, help="independent proofs per rank per step (proofs of one step run as one batched launch sequence)")
    ap.add_argument("--no-extras", action="store_true", help="skip roofline / cpu_baseline / stage breakdown")
    ap.add_argument("--workload", default="batch", choices=["batch", "sharded"],
                    help="batch: independent proofs per rank (headline, weak scaling); sharded: ONE proof of 2^log-n "
                         "split over all ranks (BASELINE configs[4], strong scaling; use --log-n 24)")
    return ap.parse_args()


def workload_name(log_n, batch):
    w = f"FibonacciSq synthetic trace 2^{log_n}, blowup 8 (LDE 2^{log_n + 3}), secret 3141592+rank"
    if batch > 1:
        w += f", {batch} independent proofs per step per GPU"
    return w


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.idx), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.p:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(power),
                       reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def run_reference(args):
    """CPU arm.  The reference crate is Rust (no cargo/rustc in the image) and its algorithms are O(n^3) with
    sizes hard-coded to n = 1024, so it cannot run this workload at all; what is timed is the oracle's NTT
    route (oracle/fast.c + merkle.c: same transcript bytes), OpenMP over all host threads, SHA-NI if present."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    threads = oracle.default_threads()
    k = args.log_n
    for _ in range(min(args.warmup, 1)):
        oracle.prove_fibsq(k, 3141592, oracle.ROUTE_FAST, threads)
    t0 = time.perf_counter()
    lde_s = 0.0
    for i in range(args.steps):
        _, _, info = oracle.prove_fibsq(k, 3141592 + i, oracle.ROUTE_FAST, threads)
        lde_s += info.t_interp + info.t_lde
    dt = time.perf_counter() - t0
    v = args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 (F_p, p=3*2^30+1) + SHA-256", "data": "synthetic",
        "config": {"workload": workload_name(k, 1), "log_n": k, "timing": "host wall clock, whole prove() incl. trace generation"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} full prove() calls at 2^{k} (oracle NTT route, OpenMP x{threads}, "
                                   f"sha_ni={bool(oracle.lib().orc_sha256_uses_shani())})"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lde_elems_per_s": args.steps * (8 << k) / lde_s if lde_s > 0 else None,
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def hashlib_sha(hex_list):
    import hashlib
    return hashlib.sha256("".join(hex_list).encode()).hexdigest()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import zkstark_b200 as zk

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    k, nb = args.log_n, args.batch
    n = 1 << k
    # all library work and all torch work of this process go to ONE side stream, so torch.cuda.Event sees the kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = zk.Context(local, stream=stream.cuda_stream)
    tl = zk.transcript_len(k)
    if args.workload == "sharded":
        run_sharded(args, ctx, zk, torch, dist, world, rank, local, stream, barrier)
        return

    # synthetic inputs: secret_j = 3141592 + rank*batch + j.  Trace generation is sequential host
    # work (prover.rs:32-39) and is an INPUT to the hot path: timed separately, reported as trace_gen_ms.
    t0 = time.perf_counter()
    traces_np = np.stack([zk.trace_fibsq(k, 3141592 + rank * nb + j) for j in range(nb)])
    trace_gen_ms = 1e3 * (time.perf_counter() - t0) / nb
    traces_host = torch.from_numpy(traces_np.view(np.int32)).pin_memory()                 # rows of n-1 (host API layout)
    traces_dev = torch.zeros((nb, n), dtype=torch.int32, device="cuda")
    traces_dev[:, : n - 1] = torch.from_numpy(traces_np.view(np.int32)).cuda()
    tr_out = torch.empty((nb, tl), dtype=torch.uint8).pin_memory()
    st_out = torch.empty((nb, 32), dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def step_resident():
        ctx.prove_fibsq_resident_ptr(k, nb, traces_dev.data_ptr(), tr_out.data_ptr(), st_out.data_ptr())

    def step_e2e():
        ctx.prove_fibsq_batch_ptr(k, nb, traces_host.data_ptr(), tr_out.data_ptr(), st_out.data_ptr())

    def timed(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        w0 = time.perf_counter()
        for a, b in evs:
            flush.fill_(1)            # L2 flush between timed iterations (not inside the event pair)
            a.record(stream)
            fn()
            b.record(stream)
        barrier()
        wall = time.perf_counter() - w0
        ms = [a.elapsed_time(b) for a, b in evs]
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()), ms, wall

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    tot_ms, ms_list, wall = timed(step_resident, args.steps)
    # parity of the N>1 path: every rank's transcripts, hashed, gathered in global order
    from zkstark_b200 import dist as zd
    digests = zd.gather_transcript_digests([tr_out[j].numpy().tobytes() for j in range(nb)], device="cuda")
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else {}
    first_hash = __import__("hashlib").sha256(tr_out[0].numpy().tobytes()).hexdigest()

    for _ in range(2):
        step_e2e()
    e2e_ms, _, _ = timed(step_e2e, args.steps)

    value = world * nb * args.steps / (tot_ms * 1e-3)
    e2e_value = world * nb * args.steps / (e2e_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": tot_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 (F_p, p=3*2^30+1) + SHA-256", "data": "synthetic",
        "config": {"workload": workload_name(k, nb), "log_n": k, "proofs_per_step_per_gpu": nb,
                   "l2": "512 MiB-free: 256 MiB buffer rewritten between timed iterations (L2 flush); per-proof working set 1.75 GB >> L2",
                   "timing": "CUDA events per step on the ctx stream (torch current stream), sum over steps, max over ranks"},
        "prove_ms": tot_ms / args.steps / nb,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(4 * (n - 1) * nb),
                "d2h_bytes_per_step": int((((tl + 63) // 64) * 64 + 256) * nb), "ms_per_step": e2e_ms / args.steps,
                "api": "zk_prove_fibsq_batch: pinned host trace in, transcript + state out (H2D/D2H inside)"},
        "trace_gen_ms": trace_gen_ms,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "transcript_sha256": first_hash,
        "transcripts_sha256_all_ranks": hashlib_sha(digests),
        "wall_s_timed_region": wall,
    }

    if rank == 0 and not args.no_extras:
        extras(args, ctx, zk, line, clocks, world, step_resident)
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_sharded(args, ctx, zk, torch, dist, world, rank, local, stream, barrier):
    """BASELINE configs[4]: ONE 2^log_n proof over all ranks (zkstark_b200/sharded.py).  Strong scaling."""
    import hashlib
    from zkstark_b200.sharded import LocalComm, ShardedProver, TorchComm
    k = args.log_n
    comm = TorchComm() if world > 1 else LocalComm(1)
    prover = ShardedProver(ctx, k, comm)
    import numpy as np
    trace = torch.from_numpy(zk.trace_fibsq(k, 3141592).view(np.int32)).pin_memory()   # host input, pinned (async H2D)
    for _ in range(max(args.warmup, 1)):
        prover.prove(trace)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    barrier()
    for a, b in evs:
        a.record(stream)
        tr, st = prover.prove(trace)
        b.record(stream)
    barrier()
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else {}
    tot = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms = float(tot.item()) / args.steps
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": 1e3 / ms, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32 (F_p, p=3*2^30+1) + SHA-256", "data": "synthetic",
            "config": {"workload": f"ONE FibonacciSq proof, trace 2^{k}, blowup 8 (LDE 2^{k + 3}), sharded over {world} GPU(s): "
                                   "coset-split LDE + NCCL all-to-all, blocked-cyclic ownership", "log_n": k,
                       "timing": "CUDA events per proof on the ctx stream, max over ranks; working set >> L2"},
            "prove_ms": ms, "lde_elems_total": 8 << k, "nvlink_bytes_per_proof": prover.comm_bytes,
            "e2e": {"value": 1e3 / ms, "unit": UNIT, "h2d_bytes_per_step": int(4 * ((1 << k) - 1)),
                    "d2h_bytes_per_step": int(zk.transcript_len(k) + 256),
                    "note": "the sharded driver takes the host trace and returns the host transcript: H2D/D2H are inside"},
            "gpu_launches": int(launches), "clocks": clocks, "transcript_sha256": hashlib.sha256(tr).hexdigest()}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def extras(args, ctx, zk, line, clocks, world, step_resident):
    """Stage breakdown, per-kernel rooflines (live CUDA-event timings) and the CPU baseline."""
    import numpy as np
    k = args.log_n
    n, N = 1 << k, 8 << k
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"

    # stage breakdown and latency of ONE prove() (device time, CUDA events inside the library)
    ctx.set_stage_timing(True)
    tr = zk.trace_fibsq(k, 3141592)
    best, call_ms = None, None
    for _ in range(5):
        t0 = time.perf_counter()
        ctx.prove_fibsq_from_trace(k, tr)
        dt = 1e3 * (time.perf_counter() - t0)
        call_ms = dt if call_ms is None else min(call_ms, dt)
        st = ctx.last_stage_ms()
        if best is None or st["total"] < best["total"]:
            best = st
    stages = best
    # wall clock of the whole host call for one proof: pageable trace upload, kernels, transcript download, host channel tail
    line["prove_call_ms_single"] = call_ms
    # the same breakdown for one batched step (all proofs of the step advance through a stage together)
    bst = None
    for _ in range(3):
        step_resident()
        st = ctx.last_stage_ms()
        if bst is None or st["total"] < bst["total"]:
            bst = st
    ctx.set_stage_timing(False)
    line["stage_ms_batched_step"] = bst
    if bst["interpolate"] + bst["lde"] > 0:
        line["lde_elems_per_s_batched"] = args.batch * N / ((bst["interpolate"] + bst["lde"]) * 1e-3)
    line["prove_ms_single"] = stages["total"]
    line["stage_ms"] = stages
    if stages["interpolate"] + stages["lde"] > 0:
        line["lde_elems_per_s"] = N / ((stages["interpolate"] + stages["lde"]) * 1e-3)

    # dominant kernel: the Merkle leaf pass (SHA-256, integer-pipe bound).  achieved = ALU-pipe instructions
    # actually issued (static SASS count per leaf x leaves) / live launch time; peak = the same pipe measured
    # live with a SHF/LOP3/IADD3 microbenchmark.  The HBM view of the same launch is given beside it.
    counts = json.load(open(os.path.join(ROOT, "zkstark_b200", "sass_counts.json")))
    try:
        counts.update(json.load(open(os.path.join(ROOT, "zkstark_b200", "ncu_counts.json"))))
    except Exception:
        pass
    log_leaves = k + 3
    ms_leaf, ms_leaf_min, _ = ctx.bench_kernel(0, log_leaves, iters=10, flush_l2=True)
    _, _, int_peak = ctx.bench_kernel(6, 0, iters=5, flush_l2=False)
    leaves = 1 << log_leaves
    # dynamic ALU-pipe instructions per leaf (one leaf hash + one node hash; loops are rolled, so this is
    # the ncu count smsp__inst_executed_pipe_alu / leaves of the committed profile, not the static SASS size)
    alu_ops = counts.get("alu_dyn_per_leaf", counts["alu_per_leaf"]) * leaves
    ach = alu_ops / (ms_leaf * 1e-3) / 1e12
    alg_bytes = 4 * leaves + 32 * (2 * leaves - (leaves >> 5))
    line["roofline"] = {
        "kernel": "tree_pass_kernel<values> (Merkle leaf pass, 2^%d leaves, 32 leaves + 31 nodes per thread)" % log_leaves,
        "bound": "int", "achieved": ach, "peak": int_peak, "unit": "Tiop/s", "frac": ach / int_peak if int_peak else None,
        "peak_source": "measured live: zk_bench_kernel(6) SHF/LOP3/IADD3 microbenchmark on this GPU",
        "ms_per_launch": ms_leaf, "compressions_per_s": 3 * leaves / (ms_leaf * 1e-3),
        "hbm": {"achieved": alg_bytes / (ms_leaf * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes / (ms_leaf * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                "bound": "hbm", "traffic": counts.get("dram_bytes_per_launch") if log_leaves == 23 else None,
                "note": "the contract's HBM view of the same launch: 6 % of the copy bandwidth, i.e. NOT the bound of this kernel"},
        # dram__bytes_read.sum + dram__bytes_write.sum of this launch in the committed `ncu --set full` capture
        "traffic": counts.get("dram_bytes_per_launch") if log_leaves == 23 else None,
        "share_of_step": None,
        "note": "SHA-256 Merkle hashing is integer-pipe bound (north_star: INT roofline for hashing, HBM roofline for NTT/fold)",
    }
    tree_ms = stages["f_tree"] + stages["cp_tree"] + stages["fri"]
    line["roofline"]["share_of_step"] = tree_ms / stages["total"] if stages["total"] else None
    # HBM-bound kernels (north_star target: NTT / fold >= 60 % of the HBM roofline at 2^20)
    rk = {}
    for name, which, ls, nbytes in [
        ("lde (all passes)", 2, k, 4 * n + 4 * N),
        ("interpolate (all passes)", 3, k, 8 * n),
        ("fri_fold layer0", 4, k + 3, 4 * N + 2 * N),
        ("cp_eval", 5, k, 8 * N),
        ("merkle tree (all passes)", 1, k + 3, 4 * N + 32 * (2 * N - 1)),
        # the same stream kernels at the size of BASELINE configs[4] (2^24 trace): the 2^20 working set is L2-sized and
        # a 16 us launch is mostly ramp-up, these show what the kernels sustain out of HBM
        ("fri_fold layer0 @2^27 points (2^24-trace size)", 4, 27, 6 << 27),
        ("lde (all passes) @2^24 trace", 2, 24, 36 << 24),
    ]:
        ms, msmin, _ = ctx.bench_kernel(which, ls, iters=10 if ls <= k + 3 else 4, flush_l2=True)
        gbs = nbytes / (ms * 1e-3) / 1e9
        rk[name] = {"bound": "hbm", "ms": ms, "algorithmic_bytes": nbytes, "achieved": gbs, "peak": hbm_peak,
                    "unit": "GB/s", "frac": gbs / hbm_peak}
    line["roofline_hbm_kernels"] = rk

    if world == 1:
        import oracle
        oracle.build()
        threads = oracle.default_threads()
        reps = 2
        oracle.prove_fibsq(max(k - 4, 3), 1, oracle.ROUTE_FAST, threads)  # thread-pool warm-up
        t0 = time.perf_counter()
        for i in range(reps):
            otr, ost, info = oracle.prove_fibsq(k, 3141592 + i, oracle.ROUTE_FAST, threads)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": reps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{reps} full prove() calls at 2^{k} with the oracle's NTT route (OpenMP x{threads}, "
                      f"sha_ni={bool(oracle.lib().orc_sha256_uses_shani())}); the reference's own O(n^3) algorithm "
                      "only exists for n = 1024",
            "ms_per_proof": 1e3 * dt / reps,
        }
        # parity spot check in the same run: the GPU transcript of secret 3141592 equals the oracle's
        tr_g, st_g = ctx.prove_fibsq_from_trace(k, tr)
        otr, ost, _ = oracle.prove_fibsq(k, 3141592, oracle.ROUTE_FAST, threads)
        line["parity_checked"] = bool(tr_g == otr and st_g == ost)


if __name__ == "__main__":
    main()
