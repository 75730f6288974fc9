#!/usr/bin/env python
"""bench.py -- headline benchmark of the zkstark_b200 hot path (contract: see DESIGN.md "Measurement").

A "step" is `--batch` (default 32) independent prove() calls (interpolate -> LDE -> Merkle -> CP -> FRI ->
decommit; prover.rs:9-293) per rank, each of one synthetic FibonacciSq trace of BASELINE.json configs[2]:
trace 2^20, blow-up 8 (8,388,608-point LDE), secret 3141592 + global proof index.  The proofs of a step run
as one batched launch sequence (the latency-bound upper Merkle levels and Fiat-Shamir steps of one proof
overlap the others); single-proof latency is reported beside it as prove_ms_single.  Independent proofs
shard across ranks with no data-path collective (weak scaling): value = proofs/s summed over all ranks.

After the headline the same line carries the other BASELINE configs, each with parity against a committed golden:
  sharded_2p24      configs[4]: ONE 2^24 proof over all N ranks (zk_prove_fibsq_sharded, NCCL inside the library), beside
                    the single-GPU time measured on rank 0 in the same run;
  batch_2p16x1024   configs[3]: 1024 independent 2^16 proofs split over the N ranks (strong scaling);
  configs           single-GPU latency of one proof at 2^10 (the reference's own configuration), 2^16 and 2^24 with the
                    CPU baselines of the small ones (2^10: the LITERAL single-thread restatement of the reference).

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
                    help="batch: independent proofs per rank (headline, weak scaling) followed by the other configs; sharded: "
                         "only ONE proof of 2^log-n split over all ranks through the Python driver (zkstark_b200/sharded.py)")
    ap.add_argument("--no-configs", action="store_true", help="skip sharded_2p24 / batch_2p16x1024 / configs")
    ap.add_argument("--sharded-log-n", type=int, default=24)
    ap.add_argument("--cpu-2p24", action="store_true", help="also time the oracle at 2^24 (35 s per proof)")
    return ap.parse_args()


def workload_name(log_n, batch):
    w = f"FibonacciSq synthetic trace 2^{log_n}, blowup 8 (LDE 2^{log_n + 3}), secret 3141592+index"
    if batch > 1:
        w += f", {batch} independent proofs per step per GPU"
    return w


def make_config(log_n, batch):
    """The `config` object of BOTH arms (the driver compares them)."""
    return {"workload": workload_name(log_n, batch), "log_n": log_n, "proofs_per_step_per_gpu": batch,
            "scope": "secret in -> proof out: trace generation (prover.rs:32-39) + prove() (prover.rs:60-292)"}


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
    warm = max(args.warmup, 3)
    for i in range(warm):
        oracle.prove_fibsq(k, 3141592 + i, oracle.ROUTE_FAST, threads)
    t0 = time.perf_counter()
    lde_s = 0.0
    for i in range(args.steps):
        _, _, info = oracle.prove_fibsq(k, 3141592 + i, oracle.ROUTE_FAST, threads)
        lde_s += info.t_interp + info.t_lde
    dt = time.perf_counter() - t0
    v = args.steps / dt
    cfg = make_config(k, args.batch)
    cfg["timing"] = "host wall clock around every step"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 (F_p, p=3*2^30+1) + SHA-256", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"each step = ONE proof of the workload's {args.batch} (secret in -> proof out, trace generation "
                                   f"included): {args.steps} prove() calls at 2^{k}, oracle NTT route, OpenMP x{threads}, "
                                   f"sha_ni={bool(oracle.lib().orc_sha256_uses_shani())}"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lde_elems_per_s": args.steps * (8 << k) / lde_s if lde_s > 0 else None,
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def leaf_hash_logic_ops():
    """SHF + LOP3 operations one Merkle leaf costs at the very least (merkle.rs:30-46: leaf = SHA256(BE32(v)), node =
    SHA256(L || R) = data block + constant padding block), with 3-input logic and funnel shifts:
      round: Sigma0, Sigma1 (3 SHF + 1 LOP3 each), Ch, Maj (1 LOP3 each) = 10; round 0 of a block that starts from the IV
             has constant inputs (0);
      schedule word: sigma0, sigma1 (3 SHF + 1 LOP3 each) = 8, nothing for a sigma of a compile-time constant: the leaf
             block is one value + 15 constant words, the padding block is all constants."""
    rounds = 10 * 64 * 3 - 10 * 2                       # leaf block and node data block start from the IV
    node_sched = 8 * 48
    const = [False] + [True] * 15                       # W0 = v, W1..W15 constants (0x80000000, 0.., 32)
    leaf_sched = 0
    for t in range(16, 64):
        leaf_sched += (0 if const[t - 2] else 4) + (0 if const[t - 15] else 4)
        const.append(const[t - 2] and const[t - 7] and const[t - 15] and const[t - 16])
    return rounds + node_sched + leaf_sched             # 1900 + 384 + 292 = 2576


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
    # the trace workers of zk_prove_stream_submit: the ranks of a node share its cores (one process per GPU), and the
    # thread that issues the launches needs one of them
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    host_threads = max(1, cores // max(local_world, 1) - (1 if local_world > 1 else 0))
    os.environ.setdefault("ZK_HOST_THREADS", str(host_threads))
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

    def step_from_trace():
        ctx.prove_fibsq_batch_ptr(k, nb, traces_host.data_ptr(), tr_out.data_ptr(), st_out.data_ptr())

    secrets_np = np.array([(3141592 + rank * nb + j) & 0xFFFFFFFF for j in range(nb)], dtype=np.uint32)

    def timed(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        w0 = time.perf_counter()
        for i, (a, b) in enumerate(evs):
            flush.fill_(1)            # L2 flush between timed iterations (not inside the event pair)
            a.record(stream)
            fn(i, steps)
            b.record(stream)
        barrier()
        wall = time.perf_counter() - w0
        ms = [a.elapsed_time(b) for a, b in evs]
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()), ms, wall

    def step_e2e(i, steps):
        # secrets in -> transcripts out through the pipelined API: the traces of step i + 1 are generated (host threads) and
        # uploaded (side stream) while step i is proved; the first step's trace generation is exposed and inside its events
        if i == 0:
            ctx.prove_stream_submit(k, secrets_np)
        if i + 1 < steps:
            ctx.prove_stream_submit(k, secrets_np)
        ctx.prove_stream_collect_ptr(tr_out.data_ptr(), st_out.data_ptr())

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    tot_ms, ms_list, wall = timed(lambda i, steps: step_resident(), args.steps)
    # parity of the N>1 path: every rank's transcripts, hashed, gathered in global order
    from zkstark_b200 import dist as zd
    digests = zd.gather_transcript_digests([tr_out[j].numpy().tobytes() for j in range(nb)], device="cuda")
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else {}
    first_hash = __import__("hashlib").sha256(tr_out[0].numpy().tobytes()).hexdigest()

    timed(step_e2e, warm)     # warm-up of the pipelined path (allocates its two trace slots)
    tr_out.zero_()
    e2e_ms, e2e_list, _ = timed(step_e2e, args.steps)
    e2e_digests = zd.gather_transcript_digests([tr_out[j].numpy().tobytes() for j in range(nb)], device="cuda")
    for _ in range(2):
        step_from_trace()
    ft_ms, _, _ = timed(lambda i, steps: step_from_trace(), args.steps)
    # one un-pipelined call, secrets in -> transcripts out (trace generation exposed)
    t_single = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        ctx.prove_fibsq_batch_secrets_ptr(k, nb, secrets_np.ctypes.data, tr_out.data_ptr(), st_out.data_ptr())
        t_single.append(1e3 * (time.perf_counter() - t0))

    value = world * nb * args.steps / (tot_ms * 1e-3)
    e2e_value = world * nb * args.steps / (e2e_ms * 1e-3)
    cfg = make_config(k, nb)
    cfg.update({"l2": "256 MiB buffer rewritten between timed iterations (L2 flush); per-proof working set 1.75 GB >> L2",
                "timing": "CUDA events per step on the ctx stream (torch current stream), sum over steps, max over ranks"})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": tot_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 (F_p, p=3*2^30+1) + SHA-256", "data": "synthetic",
        "config": cfg,
        "value_scope": "traces resident in HBM -> transcripts on the host (zk_prove_fibsq_resident)",
        "prove_ms": tot_ms / args.steps / nb,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(4 * (n - 1) * nb),
                "d2h_bytes_per_step": int((((tl + 63) // 64) * 64 + 256) * nb), "ms_per_step": e2e_ms / args.steps,
                "first_step_ms": e2e_list[0],
                "api": "zk_prove_stream_submit / zk_prove_stream_collect: host secrets in, traces generated by host threads one "
                       "step ahead and uploaded from pinned memory on a side stream (H2D), transcripts + states out (D2H); the "
                       "first step's trace generation is not overlapped and is inside the timed region",
                "h2d_note": "the library uploads the traces it generates: 4 (n - 1) bytes per proof, counted here",
                "host_threads_per_rank": int(os.environ.get("ZK_HOST_THREADS", "0")),
                "parity": bool(e2e_digests == digests)},
        "e2e_from_trace": {"value": world * nb * args.steps / (ft_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(4 * (n - 1) * nb),
                           "api": "zk_prove_fibsq_batch: pinned host traces in (H2D inside), transcripts out"},
        "e2e_single_call_ms": min(t_single),
        "e2e_single_call_note": "zk_prove_fibsq_batch_secrets, one call, host wall clock: the trace generation of the (only) "
                                "sub-batch is exposed in front of the proving",
        "trace_gen_ms": trace_gen_ms,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "transcript_sha256": first_hash,
        "transcripts_sha256_all_ranks": hashlib_sha(digests),
        "wall_s_timed_region": wall,
    }

    if rank == 0 and not args.no_extras:
        extras(args, ctx, zk, line, clocks, world, step_resident)
    if not args.no_configs:
        del traces_dev, flush
        torch.cuda.empty_cache()
        other_configs(args, ctx, zk, torch, dist, world, rank, local, stream, barrier, line)
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def other_configs(args, ctx, zk, torch, dist, world, rank, local, stream, barrier, line):
    """BASELINE configs[3], configs[4] and the single-GPU latencies, each checked against a committed golden."""
    import hashlib

    import numpy as np
    from zkstark_b200 import dist as zd

    def dev_max(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    gold_large = json.load(open(os.path.join(ROOT, "tests", "golden", "large_digests.json")))
    ctx.trim()

    # ---- configs[3]: 1024 independent 2^16 proofs split over the ranks (strong scaling, no collective)
    k16, total = 16, 1024
    lo, hi = zd.shard_range(total, world, rank)
    secrets = np.array([zd.proof_secret(j) for j in range(lo, hi)], dtype=np.uint32)
    tl16 = zk.transcript_len(k16)
    tr16 = torch.empty((hi - lo, tl16), dtype=torch.uint8).pin_memory()
    st16 = torch.empty((hi - lo, 32), dtype=torch.uint8).pin_memory()

    def batch16():
        ctx.prove_fibsq_batch_secrets_ptr(k16, hi - lo, secrets.ctypes.data, tr16.data_ptr(), st16.data_ptr())

    for _ in range(2):
        batch16()
    reps = 5
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    barrier()
    for a, b in evs:
        a.record(stream)
        batch16()
        b.record(stream)
    barrier()
    ms16 = dev_max(sum(a.elapsed_time(b) for a, b in evs)) / reps
    d16 = zd.gather_transcript_digests([tr16[j].numpy().tobytes() for j in range(hi - lo)], device="cuda")
    try:
        gold16 = json.load(open(os.path.join(ROOT, "tests", "golden", "batch_2p16_digest.json")))
        parity16 = bool(hashlib.sha256("".join(d16).encode()).hexdigest() == gold16["checksum_of_checksums"][str(total)])
    except Exception:
        parity16 = None
    line["batch_2p16x1024"] = {
        "workload": "BASELINE configs[3]: 1024 independent 2^16-trace proofs (secret 3141592 + j) split over the ranks, "
                    "secrets in -> transcripts out (zk_prove_fibsq_batch_secrets)",
        "proofs_per_s": total / (ms16 * 1e-3), "ms": ms16, "n_gpus": world, "scaling": "strong",
        "parity_vs_golden": parity16, "golden": "tests/golden/batch_2p16_digest.json (oracle, checksum of checksums)"}
    del tr16, st16
    ctx.trim()

    # ---- configs[4]: ONE 2^24 proof over all ranks (native driver: NCCL inside the library)
    ks = args.sharded_log_n
    trace_np = zk.trace_fibsq(ks, 3141592)
    trace = torch.from_numpy(trace_np.view(np.int32)).pin_memory()
    tls = zk.transcript_len(ks)
    trs = torch.empty(tls, dtype=torch.uint8).pin_memory()
    sts = torch.empty(32, dtype=torch.uint8).pin_memory()
    ms_1gpu = ms_1gpu_min = None
    if rank == 0:
        # the single-GPU time of the same proof, same run, same GPU model: the monolithic pipeline on rank 0 alone
        for _ in range(2):
            ctx.prove_fibsq_from_trace_ptr(ks, trace.data_ptr(), trs.data_ptr(), tls, sts.data_ptr())
        e = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
        for a, b in e:
            a.record(stream)
            ctx.prove_fibsq_from_trace_ptr(ks, trace.data_ptr(), trs.data_ptr(), tls, sts.data_ptr())
            b.record(stream)
        torch.cuda.synchronize()
        ms_1gpu_all = [a.elapsed_time(b) for a, b in e]
        ms_1gpu, ms_1gpu_min = sum(ms_1gpu_all) / len(ms_1gpu_all), min(ms_1gpu_all)
        mono_hash = hashlib.sha256(trs.numpy().tobytes()).hexdigest()
        ctx.trim()
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(zk.shard_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        comm = zk.ShardComm(ctx, world, rank, bytes(uid.cpu().numpy()))
    else:
        comm = zk.ShardComm.local(ctx, 1)
    for _ in range(2):
        comm.prove_ptr(ks, trace.data_ptr(), trs.data_ptr(), tls, sts.data_ptr())
    reps = 10
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    barrier()
    phases = []
    for a, b in evs:
        a.record(stream)
        comm.prove_ptr(ks, trace.data_ptr(), trs.data_ptr(), tls, sts.data_ptr())
        b.record(stream)
        phases.append(comm.last_phase_ms())
    barrier()
    # every repetition: the slowest rank's time (element-wise MAX over ranks); then mean / median / min over repetitions
    per_rep = torch.tensor([a.elapsed_time(b) for a, b in evs], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(per_rep, op=dist.ReduceOp.MAX)
    per_rep = sorted(float(x) for x in per_rep.cpu())
    ms_sh, ms_sh_min, ms_sh_med = sum(per_rep) / reps, per_rep[0], per_rep[reps // 2]
    sh_hash = hashlib.sha256(trs.numpy().tobytes()).hexdigest()
    all_hash = zd.gather_transcript_digests([trs.numpy().tobytes()], device="cuda")
    gold = gold_large.get(str(ks), {})
    best = min(phases, key=lambda p: p["total"])
    line["sharded_2p%d" % ks] = {
        "workload": f"BASELINE configs[4]: ONE FibonacciSq proof, trace 2^{ks} (LDE 2^{ks + 3}), sharded over {world} GPU(s): "
                    "coset-split LDE, one exchange through NVLink peer memory fused into the unpack kernel (NCCL all-to-all "
                    "as fallback), blocked-cyclic ownership (zk_prove_fibsq_sharded); pinned host trace in, transcript out",
        "ms": ms_sh, "ms_median": ms_sh_med, "ms_min": ms_sh_min, "reps": reps,
        "ms_note": "mean over the repetitions of the slowest rank's device time per repetition (host part of the call included)",
        "ms_1gpu": ms_1gpu, "ms_1gpu_min": ms_1gpu_min,
        "efficiency": (ms_1gpu / (world * ms_sh)) if ms_1gpu else None, "n_gpus": world,
        "scaling": "strong", "nvlink_bytes_per_rank": comm.last_comm_bytes(), "first_tail_layer": comm.last_first_tail(),
        "phase_ms": {a: round(b, 4) for a, b in best.items()},
        "parity": bool(gold and sh_hash == gold.get("sha256") and sts.numpy().tobytes().hex() == gold.get("state")
                       and len(set(all_hash)) == 1),
        "parity_1gpu": bool(gold and ms_1gpu is not None and mono_hash == gold.get("sha256")) if rank == 0 else None,
        "golden": "tests/golden/large_digests.json (oracle)"}
    comm.close()
    ctx.trim()

    # ---- single-GPU latency of ONE proof per config (rank 0), beside its CPU baseline
    if rank == 0:
        cfgs = {}
        ctx.set_stage_timing(True)
        tdata = json.load(open(os.path.join(ROOT, "tests", "golden", "transcripts.json")))
        for kk in (10, 16):
            tr = zk.trace_fibsq(kk, 3141592)
            best_dev, best_call, out = None, None, None
            for _ in range(8):
                t0 = time.perf_counter()
                out = ctx.prove_fibsq_from_trace(kk, tr)
                dt = 1e3 * (time.perf_counter() - t0)
                best_call = dt if best_call is None else min(best_call, dt)
                sm = ctx.last_stage_ms()["total"]
                best_dev = sm if best_dev is None else min(best_dev, sm)
            ent = {"prove_ms_device": best_dev, "prove_call_ms_host_wall": best_call}
            g = tdata.get(str(kk)) if isinstance(tdata, dict) else None
            if g and "sha256" in g:
                ent["parity"] = bool(hashlib.sha256(out[0]).hexdigest() == g["sha256"])
            cfgs["2^%d" % kk] = ent
        ctx.set_stage_timing(False)
        cfgs["2^%d" % ks] = {"prove_ms_device": ms_1gpu, "parity": line["sharded_2p%d" % ks]["parity_1gpu"]}
        try:
            import oracle
            oracle.build()
            th = oracle.default_threads()
            t0 = time.perf_counter()
            otr, _, _ = oracle.prove_fibsq(10, 3141592, oracle.ROUTE_LITERAL, 1)
            cfgs["2^10"]["cpu_ms"] = 1e3 * (time.perf_counter() - t0)
            cfgs["2^10"]["cpu_kind"] = ("port: LITERAL restatement of the reference's own O(n^3) algorithm (oracle/prover.c), 1 thread, "
                                        "1 sample -- what `cargo run --release` does on a single core")
            oracle.prove_fibsq(16, 1, oracle.ROUTE_FAST, th)
            t0 = time.perf_counter()
            for i in range(5):
                oracle.prove_fibsq(16, 3141592 + i, oracle.ROUTE_FAST, th)
            cfgs["2^16"]["cpu_ms"] = 1e3 * (time.perf_counter() - t0) / 5
            cfgs["2^16"]["cpu_kind"] = f"port: oracle NTT route, OpenMP x{th}, 5 samples"
            if args.cpu_2p24:
                t0 = time.perf_counter()
                oracle.prove_fibsq(ks, 3141592, oracle.ROUTE_FAST, th)
                cfgs["2^%d" % ks]["cpu_ms"] = 1e3 * (time.perf_counter() - t0)
        except Exception as ex:      # the CPU legs are a reported baseline, never a reason to lose the line
            cfgs["cpu_error"] = repr(ex)
        line["configs"] = cfgs


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
    fresh = False
    try:
        nc = json.load(open(os.path.join(ROOT, "zkstark_b200", "ncu_counts.json")))
        import hashlib
        h = hashlib.sha256()
        for src in ("sha256.cuh", "tree.cu", "field.cuh", "common.cuh"):
            h.update(open(os.path.join(ROOT, "zkstark_b200", "csrc", src), "rb").read())
        fresh = nc.get("source_sha256") == h.hexdigest()      # the capture describes the code that is being timed
        counts.update(nc)
    except Exception:
        pass
    log_leaves = k + 3
    ms_leaf, ms_leaf_min, _ = ctx.bench_kernel(0, log_leaves, iters=10, flush_l2=True)
    _, _, int_peak = ctx.bench_kernel(6, 0, iters=5, flush_l2=False)
    leaves = 1 << log_leaves
    # ALGORITHMIC integer work per leaf (one leaf hash + one node hash = 3 compressions): the funnel shifts and 3-input
    # logic operations SHA-256 cannot do without (leaf_hash_logic_ops below).  They can only run on the ALU pipe; the
    # additions can (and in this kernel do) go to the FMA pipe.
    alg_ops_per_leaf = leaf_hash_logic_ops()
    ach_alg = alg_ops_per_leaf * leaves / (ms_leaf * 1e-3) / 1e12
    # EXECUTED ALU-pipe instructions per leaf (ncu capture of this very code, else the algorithmic count)
    alu_dyn = counts.get("alu_dyn_per_leaf") if fresh else None
    ach = (alu_dyn if alu_dyn else alg_ops_per_leaf) * leaves / (ms_leaf * 1e-3) / 1e12
    alg_bytes = 4 * leaves + 32 * (2 * leaves - (leaves >> 5))
    traffic = counts.get("dram_bytes_per_launch") if (log_leaves == 23 and fresh) else None
    line["roofline"] = {
        "kernel": "tree_pass_kernel<values> (Merkle leaf pass, 2^%d leaves, 32 leaves + 31 nodes per thread + one CTA epilogue level: "
                  "the production launch shape)" % log_leaves,
        "bound": "int", "achieved": ach, "peak": int_peak, "unit": "Tiop/s", "frac": ach / int_peak if int_peak else None,
        "achieved_is": ("executed ALU-pipe instructions (ncu capture of this source, zkstark_b200/ncu_counts.json) / live launch time"
                        if alu_dyn else "algorithmic SHF/LOP3 operations / live launch time (no ncu capture of this exact source)"),
        "achieved_algorithmic": ach_alg, "frac_algorithmic": ach_alg / int_peak if int_peak else None,
        "algorithmic_ops_per_leaf": alg_ops_per_leaf, "executed_alu_per_leaf": alu_dyn, "ncu_counts_match_source": fresh,
        "int_peak_tiops": int_peak,
        "peak_source": "measured live in this run: zk_bench_kernel(6) SHF/LOP3/IADD3 microbenchmark on this GPU (MEASURED_PEAKS.json has no integer figure)",
        "ms_per_launch": ms_leaf, "compressions_per_s": 3 * leaves / (ms_leaf * 1e-3),
        "hbm": {"achieved": alg_bytes / (ms_leaf * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes / (ms_leaf * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                "bound": "hbm", "traffic": traffic,
                "note": "the contract's HBM view of the same launch: ~6 % of the copy bandwidth, i.e. NOT the bound of this kernel"},
        # dram__bytes_read.sum + dram__bytes_write.sum of this launch in the committed `ncu --set full` capture
        "traffic": traffic,
        "share_of_step": None,
        "note": "SHA-256 Merkle hashing is integer-pipe bound (north_star: INT roofline for hashing, HBM roofline for NTT/fold)",
    }
    tree_ms = stages["f_tree"] + stages["cp_tree"] + stages["fri"]
    line["roofline"]["share_of_step"] = tree_ms / stages["total"] if stages["total"] else None
    # HBM-bound kernels (north_star target: NTT / fold >= 60 % of the HBM roofline at 2^20)
    rk = {}
    traffic_rec, traffic_ok = {}, False
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import ncu_traffic
        traffic_rec = json.load(open(os.path.join(ROOT, "zkstark_b200", "ncu_traffic.json")))
        traffic_ok = traffic_rec.get("source_sha256") == ncu_traffic.source_hash()
    except Exception:
        pass
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
        # dram__bytes_read + write and lts__t_bytes of the same launches in the committed ncu capture (tools/ncu_traffic.py),
        # used only while the capture describes the sources being timed
        t = traffic_rec.get("kernels", {}).get(f"{which}:{ls}") if traffic_ok else None
        rk[name]["traffic"] = t["dram_bytes"] if t else None
        rk[name]["l2_bytes"] = t["l2_bytes"] if t else None
        rk[name]["traffic_source"] = ("zkstark_b200/ncu_traffic.json (ncu, per launch sequence behind an L2 flush; matches the sources)"
                                      if t else "no capture of the current sources (tools/ncu_traffic.sh)")
    # the transforms are integer-ISSUE bound, not HBM bound (p > 2^31: a butterfly is 10 instructions, 3 of them 64-bit-result
    # multiplications): their roofline is the butterfly rate of the same arithmetic without any memory traffic, measured live
    try:
        _, _, tbf = ctx.bench_kernel(8, 0, iters=5, flush_l2=False)       # tera-butterflies / s
        _, _, tmm = ctx.bench_kernel(8, 1, iters=5, flush_l2=False)       # tera-Montgomery-products / s
        for name, kk, arrays, extra_mul in (("lde (all passes)", k, 8, 3), ("lde (all passes) @2^24 trace", 24, 8, 5), ("interpolate (all passes)", k, 1, 2)):
            pts = arrays << kk
            bf = pts * kk / 2                                             # butterflies of the transform(s)
            passes = (kk + 9) // 10
            mm = pts * (2 * (passes - 1) + (1 if arrays == 8 else 0))     # inter-pass twiddle (apply + chain) and coset products
            floor_ms = 1e3 * (bf / (tbf * 1e12) + mm / (tmm * 1e12))
            rk[name]["instr_roofline"] = {"bound": "int-issue", "butterflies": bf, "extra_products": mm, "floor_ms": floor_ms,
                                          "frac": floor_ms / rk[name]["ms"],
                                          "peak": tbf, "peak_unit": "T butterflies/s (zk_bench_kernel(8), live)"}
    except Exception as ex:
        rk["instr_roofline_error"] = repr(ex)
    line["roofline_hbm_kernels"] = rk

    if world == 1:
        import oracle
        oracle.build()
        threads = oracle.default_threads()
        reps = 5
        oracle.prove_fibsq(k, 1, oracle.ROUTE_FAST, threads)  # warm-up (thread pool, page faults)
        t0 = time.perf_counter()
        for i in range(reps):
            otr, ost, info = oracle.prove_fibsq(k, 3141592 + i, oracle.ROUTE_FAST, threads)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": reps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{reps} proofs of the workload (secret in -> proof out, trace generation included) at 2^{k} with the oracle's NTT route (OpenMP x{threads}, "
                      f"sha_ni={bool(oracle.lib().orc_sha256_uses_shani())}); the reference's own O(n^3) algorithm "
                      "only exists for n = 1024",
            "ms_per_proof": 1e3 * dt / reps,
        }
        # parity spot check in the same run: the GPU transcript of secret 3141592 equals the oracle's
        tr_g, st_g = ctx.prove_fibsq_from_trace(k, tr)
        otr, ost, _ = oracle.prove_fibsq(k, 3141592, oracle.ROUTE_FAST, threads)
        line["parity_checked"] = bool(tr_g == otr and st_g == ost)


def _only_the_json_line_on_stdout():
    """The contract is ONE JSON line on stdout.  Libraries underneath print there too (NCCL writes its version banner
    to stdout from C when NCCL_DEBUG is set in the environment), so file descriptor 1 is pointed at stderr for the
    whole run and Python's sys.stdout keeps the real one: only print() of this script reaches it."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


if __name__ == "__main__":
    _only_the_json_line_on_stdout()
    main()
