"""One FibonacciSq proof sharded over several GPUs (BASELINE configs[4]).

SPMD, one process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch) for the exchanges, the C-ABI
`zk_shard_*_dev` kernels for all field / hash work.  The transcript is byte-identical to the single-GPU one.

Layout.  Ranks first split the LDE by COSET (f_eval[8j+r] = NTT_n(c_k (5 h^r)^k)[j]: eight independent
size-n transforms, no communication inside a transform); ONE all-to-all then re-owns the evaluations
blocked-cyclically: global index i -> rank (i >> log_block) % G.  With that ownership
  * every FRI fold pairs i with i + len/2 on the same rank (no exchange per layer),
  * the bottom log_block levels of every Merkle tree are rank-local; only the block roots travel
    (all-gather of len/2^log_block digests), the small upper tree and the Fiat-Shamir step run replicated
    on every rank's device, so all ranks draw identical challenges without a host round trip,
  * the composition polynomial needs a 16-value halo per block (all-gather of block heads).
When a layer no longer has two blocks per rank it is all-gathered and the tail of FRI runs replicated.

The same code runs G "virtual ranks" inside one process on one GPU (LocalComm) -- that is how the single-GPU
test tier covers the multi-rank logic; TorchComm is the real multi-process path.
"""
import hashlib
import struct

import numpy as np

P = 3221225473
POST_NONE, POST_F, POST_CP, POST_FRI = 0, 1, 2, 3


class LocalComm:
    """G virtual ranks in one process: collectives are tensor shuffles on one device."""

    def __init__(self, world_size):
        self.world_size = world_size
        self.my_ranks = list(range(world_size))

    def all_to_all(self, send):  # send[r]: tensor whose dim 0 is split evenly by destination
        import torch
        G = self.world_size
        chunks = [s.chunk(G, dim=0) for s in send]
        return [torch.cat([chunks[src][dst] for src in range(G)], dim=0) for dst in range(G)]

    def all_gather(self, xs):  # -> per local rank: tensor [G, ...]
        import torch
        full = torch.stack(list(xs), dim=0)
        return [full for _ in xs]

    def all_reduce_max(self, xs):
        import torch
        m = torch.stack(list(xs), dim=0).amax(dim=0)
        return [m for _ in xs]


class TorchComm:
    """One real rank per process (torch.distributed already initialised, NCCL on GPUs)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.my_ranks = [dist.get_rank(group)]

    def all_to_all(self, send):
        import torch
        out = torch.empty_like(send[0])
        self.dist.all_to_all_single(out, send[0].contiguous(), group=self.group)
        return [out]

    def all_gather(self, xs):
        import torch
        x = xs[0].contiguous()
        out = torch.empty((self.world_size,) + tuple(x.shape), dtype=x.dtype, device=x.device)
        self.dist.all_gather_into_tensor(out.view(-1), x.view(-1), group=self.group)
        return [out]

    def all_reduce_max(self, xs):
        self.dist.all_reduce(xs[0], op=self.dist.ReduceOp.MAX, group=self.group)
        return [xs[0]]


def pack_cosets(E, G, log_block):
    """E[c][j] = f_eval[8 j + (r + c G)] (the C = 8/G coset transforms of rank r, j < n) -> all-to-all send buffer
    [d][c][bl][tt] with j = (bl G + d) jb + tt: destination rank d owns the blocks bl G + d."""
    C_, n = E.shape
    jb = 1 << (log_block - 3)                  # j-values per block (a block holds all 8 cosets of jb points)
    nbl = n // jb // G                         # blocks per rank
    return E.view(C_, nbl, G, jb).permute(2, 0, 1, 3).contiguous()


def unpack_blocks(recv, G, log_block):
    """all-to-all result [g][c][bl][tt] (from source rank g) -> rank-local evaluations in blocked-cyclic order:
    local[(bl jb + tt) 8 + (c G + g)]."""
    jb = 1 << (log_block - 3)
    C_ = 8 // G
    nbl = recv.numel() // (G * C_ * jb)
    return recv.view(G, C_, nbl, jb).permute(2, 3, 1, 0).contiguous().view(-1)


def local_to_global(loc, G, rank, log_block):
    """index map of the blocked-cyclic ownership (csrc/common.cuh shard_global)"""
    return ((((loc >> log_block) * G) + rank) << log_block) | (loc & ((1 << log_block) - 1))


class _Tree:
    """One committed layer on one rank: local values, local heap (bottom levels), replicated upper heap."""
    def __init__(self, vals, heap, upper, length, distributed):
        self.vals, self.heap, self.upper, self.length, self.distributed = vals, heap, upper, length, distributed


class ShardedProver:
    def __init__(self, ctx, log_n, comm, log_block=10, device=None, tail_len=1 << 18):
        import torch
        self.torch = torch
        self.ctx, self.k, self.comm, self.lb = ctx, log_n, comm, log_block
        self.G = comm.world_size
        self.log_N = log_n + 3
        self.N = 1 << self.log_N
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        # torch ops (copies, packs, collectives) and the ctx kernels are interleaved through raw pointers: they must be
        # ordered by ONE stream.  Create the context on torch's current stream (zk.Context(dev, stream=s.cuda_stream)).
        if ctx.stream() != torch.cuda.current_stream(self.dev).cuda_stream:
            raise ValueError("ShardedProver: the zk context must enqueue on torch's current stream "
                             "(zk.Context(device, stream=torch.cuda.current_stream().cuda_stream))")
        if self.G not in (1, 2, 4, 8):
            raise ValueError("world size must be 1, 2, 4 or 8 (the LDE has 8 cosets)")
        if log_block < 5 or (self.N // self.G) < (1 << (log_block + 1)):
            raise ValueError("proof too small to shard: need at least two blocks per rank")
        # layers of at most tail_len points are gathered and finished replicated on every rank: below ~2^18 points a
        # sharded layer costs more in launches and collectives (~110 us) than hashing it whole on one GPU
        self.tail_len = tail_len
        self.comm_bytes = 0
        self.profile = None      # set to {} to collect per-phase wall times (synchronising: for diagnosis only)
        self._t_last = None

    def _mark(self, name):
        if self.profile is None:
            return
        import time
        self.torch.cuda.synchronize()
        now = time.perf_counter()
        if self._t_last is not None:
            self.profile[name] = self.profile.get(name, 0.0) + (now - self._t_last)
        self._t_last = now

    # ------------------------------------------------------------------ helpers
    def _i32(self, *shape):
        return self.torch.empty(shape, dtype=self.torch.int32, device=self.dev)

    def _u8(self, *shape):
        return self.torch.empty(shape, dtype=self.torch.uint8, device=self.dev)

    def _commit_tree(self, vals_by_rank, length, post, fri_layer, states, transcripts, fold_from=None, cp_from=None):
        """Distributed Merkle tree of a layer of global length `length`; returns per-rank _Tree.
        fold_from = (cur_by_rank, layer_k): the layer does not exist yet -- the leaf pass of the block subtrees
        folds layer_k into vals_by_rank first (one fused pass).  cp_from = (f_by_rank, halo_by_rank): likewise for the
        composition polynomial."""
        t, lb, G = self.torch, self.lb, self.G
        local_len = length // G
        nbl = local_len >> lb
        nbt = length >> lb
        heaps, roots = [], []
        for idx, r in enumerate(self.comm.my_ranks):
            heap = self._u8(32 * (2 * local_len - 1))
            if cp_from is not None:
                f_loc, halos = cp_from
                self.ctx.shard_cp_blocks(f_loc[idx].data_ptr(), halos[idx].data_ptr(), local_len, self.k, G, r, lb,
                                         states[idx].data_ptr(), vals_by_rank[idx].data_ptr(), heap.data_ptr())
            elif fold_from is None:
                self.ctx.shard_merkle_blocks(vals_by_rank[idx].data_ptr(), local_len, lb, heap.data_ptr())
            else:
                cur, layer_k = fold_from
                self.ctx.shard_fold_blocks(cur[idx].data_ptr(), 2 * local_len, layer_k, self.log_N, G, r, lb,
                                           states[idx].data_ptr(), vals_by_rank[idx].data_ptr(), heap.data_ptr())
            heaps.append(heap)
            roots.append(heap[32 * (nbl - 1): 32 * (2 * nbl - 1)].view(nbl, 32))
        gathered = self.comm.all_gather(roots)           # [G, nbl, 32] on every rank
        self.comm_bytes += 32 * nbt
        trees = []
        log_nbt = nbt.bit_length() - 1
        for idx, r in enumerate(self.comm.my_ranks):
            upper = self._u8(32 * (2 * nbt - 1))
            upper[32 * (nbt - 1):] = gathered[idx].permute(1, 0, 2).reshape(-1)   # global block b = bl*G + g
            self.ctx.shard_upper_tree(upper.data_ptr(), log_nbt, post, fri_layer, self.k,
                                      states[idx].data_ptr(), transcripts[idx].data_ptr())
            trees.append(_Tree(vals_by_rank[idx], heaps[idx], upper, length, True))
        return trees

    # ------------------------------------------------------------------ the proof
    def prove(self, trace):
        """trace: numpy uint32 array of 2^k - 1 values (identical on every rank).  Returns (transcript, state)."""
        t, ctx, k, lb, G = self.torch, self.ctx, self.k, self.lb, self.G
        n, N, log_N = 1 << k, self.N, self.log_N
        ranks = self.comm.my_ranks
        nr = len(ranks)
        tl = ctx_transcript_len(k)
        if isinstance(trace, t.Tensor):                 # host tensor of n-1 int32 words, ideally pinned (async H2D)
            host = trace.view(t.int32)
        else:
            host = t.from_numpy(np.ascontiguousarray(trace, dtype=np.uint32).view(np.int32))
        assert host.numel() == n - 1
        a_first, a_last = int(host[0].item()) & 0xFFFFFFFF, int(host[n - 2].item()) & 0xFFFFFFFF
        self.comm_bytes = 0
        self._t_last = None
        self._mark("start")

        states = [t.zeros(64, dtype=t.int32, device=self.dev) for _ in ranks]
        transcripts = [t.zeros(tl + 64, dtype=t.uint8, device=self.dev) for _ in ranks]
        for i in range(nr):
            ctx.shard_state_init(states[i].data_ptr(), a_first, a_last)

        # Part 1a: interpolation, replicated (1/8 of the transform work).  Each real rank uploads only its
        # 1/G slice of the trace over PCIe; the slices are all-gathered over NVLink.
        if nr == 1 and G > 1:
            chunk = n // G
            lo = ranks[0] * chunk
            hi = min(lo + chunk, n - 1)
            part = self._i32(chunk)
            if hi - lo < chunk:
                part[hi - lo:].zero_()
            part[: hi - lo].copy_(host[lo:hi], non_blocking=True)
            y = self.comm.all_gather([part])[0].view(-1)
            self.comm_bytes += 4 * n * (G - 1) // G
        else:
            y = self._i32(n)
            y[n - 1:].zero_()
            y[: n - 1].copy_(host, non_blocking=True)
        coeffs = self._i32(n)
        ctx.interpolate_subgroup_dev(y.data_ptr(), n - 1, k, coeffs.data_ptr())
        self._mark("h2d+interpolate")
        # Part 1b: LDE, cosets split over ranks, then the all-to-all that re-owns the evaluations by blocks
        C_ = 8 // G
        jb = 1 << (lb - 3)                       # j-values per block (a block holds all 8 cosets of jb points)
        nbl = (N // G) >> lb                     # blocks per rank
        work = self._i32(n)
        send = []
        for r in ranks:
            E = self._i32(C_, n)
            for c in range(C_):
                ctx.shard_lde_coset(coeffs.data_ptr(), k, r + c * G, work.data_ptr(), E[c].data_ptr())
            send.append(pack_cosets(E, G, lb))
        self._mark("lde_cosets+pack")
        recv = self.comm.all_to_all(send)
        self._mark("all_to_all")
        self.comm_bytes += 4 * N * (G - 1) // G
        f_local = []
        for idx in range(nr):
            f_local.append(unpack_blocks(recv[idx], G, lb))
        del send, recv, work
        self._mark("unpack")
        f_trees = self._commit_tree(f_local, N, POST_F, 0, states, transcripts)
        self._mark("f_tree")

        # Part 2: composition polynomial (needs f at i, i+8, i+16: 16-value halo per block)
        heads = [f.view(nbl, 1 << lb)[:, :16].contiguous() for f in f_local]
        heads_all = self.comm.all_gather(heads)  # [G, nbl, 16]
        self.comm_bytes += 4 * 16 * (N >> lb)
        halos = []
        for idx, r in enumerate(ranks):
            if r + 1 < G:
                halos.append(heads_all[idx][r + 1].contiguous())
            else:
                halos.append(t.roll(heads_all[idx][0], shifts=-1, dims=0).contiguous())   # block b+1 wraps to rank 0, next bl
        cp_local = [self._i32(N // G) for _ in ranks]
        self._mark("halo")
        layers = [self._commit_tree(cp_local, N, POST_CP, 0, states, transcripts, cp_from=(f_local, halos))]
        self._mark("cp_tree")

        # Part 3: FRI.  Folds stay rank-local while a layer has >= 2 blocks per rank.
        cur, length, i = cp_local, N, 0
        while (length // G) >= (2 << lb) and length > self.tail_len:
            nxt = [self._i32(length // G // 2) for _ in ranks]
            layers.append(self._commit_tree(nxt, length // 2, POST_FRI, i + 1, states, transcripts, fold_from=(cur, i)))
            i += 1
            length //= 2
            cur = nxt
        self._mark("fri_sharded_layers")
        # replicated tail: gather layer i (one block per rank) and finish on every rank
        first_tail = i
        full = self.comm.all_gather([c.view(-1, 1 << lb) for c in cur])     # [G, nbl=1, blk]
        self.comm_bytes += 4 * length
        tails = []
        for idx in range(nr):
            arena_vals = self._i32(2 * length)
            arena_heaps = self._u8(64 * 2 * length)
            arena_vals[:length] = full[idx].permute(1, 0, 2).reshape(-1)
            ctx.shard_fri_tail(k, first_tail, arena_vals.data_ptr(), arena_heaps.data_ptr(),
                               states[idx].data_ptr(), transcripts[idx].data_ptr())
            tails.append((arena_vals, arena_heaps))
        ctx.synchronize()
        self._mark("fri_tail")

        # Part 4: decommit (prover.rs:263-289).  Every rank gathers the values and sibling digests it owns straight
        # into its (otherwise identical) transcript buffer; a byte-wise MAX all-reduce completes it everywhere.
        for idx, r in enumerate(ranks):
            ptrs = []
            for tree in [f_trees[idx]] + [layers[li][idx] for li in range(first_tail + 1)]:
                ptrs += [tree.vals.data_ptr(), tree.heap.data_ptr(), tree.upper.data_ptr()]
            ctx.shard_decommit(k, G, r, lb, first_tail, ptrs, tails[idx][0].data_ptr(), tails[idx][1].data_ptr(),
                               states[idx].data_ptr(), transcripts[idx].data_ptr())
        full = self.comm.all_reduce_max(transcripts)
        self.comm_bytes += tl
        results = []
        flag_off = (tl + 3) & ~3
        for idx in range(nr):
            host_tr = full[idx].cpu().numpy()
            if int(host_tr[flag_off: flag_off + 4].view(np.uint32)[0]) != 0:     # set by zk_shard_decommit_dev
                raise RuntimeError("last FRI layer is not constant (prover.rs:238)")
            tr = host_tr[:tl].tobytes()
            st_host = states[idx].cpu().numpy().view(np.uint32)
            # channel.rs:19-26 over the decommitment records (host; k + 4 sequential hashes)
            state = b"".join(struct.pack(">I", int(w)) for w in st_host[47:55])
            o2 = 84 + 36 * k
            for rr in range(4 + k):
                ln_rec = 12 + 32 * log_N if rr < 4 else 24 + 64 * (log_N - (rr - 4))
                state = hashlib.sha256(state + tr[o2: o2 + ln_rec]).digest()
                o2 += ln_rec
            assert o2 == tl, (o2, tl)
            results.append((tr, state))
        self._mark("decommit+host_tail")
        self.first_tail_layer = first_tail
        return results[0] if nr == 1 else results


def ctx_transcript_len(k):
    from ._binding import transcript_len
    return transcript_len(k)
