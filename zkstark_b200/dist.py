"""Multi-GPU plumbing (one process per GPU, torch.distributed): host-side logic only.

Independent proofs are the unit of sharding (no data-path collective); these helpers are what
bench.py and the batch drivers use to split work and to combine timings/results.  Backend-agnostic: NCCL on the
GPU box, gloo in the CPU tests (tests/test_dist_cpu.py)."""
import hashlib
import os


def world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_items, world_size, rank):
    """Contiguous, balanced split of n_items units: ranks < n_items % world_size get one extra."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size/rank")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def proof_secret(global_index, base_secret=3141592):
    """Synthetic inputs of BASELINE configs[3]: secret_j = 3141592 + j."""
    return (base_secret + global_index) & 0xFFFFFFFF


def max_over_ranks(value, device=None):
    """Timing rule: every multi-GPU number is the max over ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_transcript_digests(transcripts, device=None):
    """All ranks learn sha256 of every rank's transcripts, in global proof order (a checksum of checksums):
    returns a list of hex digests, rank-major.  `transcripts` is a list of bytes owned by this rank."""
    import torch
    import torch.distributed as dist
    mine = b"".join(hashlib.sha256(t).digest() for t in transcripts)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [mine[i:i + 32].hex() for i in range(0, len(mine), 32)]
    ws = dist.get_world_size()
    counts = torch.tensor([len(transcripts)], dtype=torch.int64, device=device or "cpu")
    all_counts = [torch.zeros_like(counts) for _ in range(ws)]
    dist.all_gather(all_counts, counts)
    mx = max(int(c.item()) for c in all_counts)
    buf = torch.zeros(mx * 32, dtype=torch.uint8, device=device or "cpu")
    if mine:
        buf[: len(mine)] = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(buf.device)
    bufs = [torch.zeros_like(buf) for _ in range(ws)]
    dist.all_gather(bufs, buf)
    out = []
    for c, b in zip(all_counts, bufs):
        raw = bytes(b[: int(c.item()) * 32].cpu().numpy())
        out += [raw[i:i + 32].hex() for i in range(0, len(raw), 32)]
    return out
