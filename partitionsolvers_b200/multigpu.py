"""Multi-GPU drivers: one process per GPU, torch.distributed for the plumbing.

Two ways the path shards:

* :func:`solve_replicas` -- independent instances (randomization-test replicas, T-sweep points).  Rank r
  takes a contiguous block of the instances, runs ONE batched engine call, and the per-instance results
  are gathered on rank 0.  No collective touches the data path.

* :class:`ShardedSolver` -- one very large instance.  Rows of a DP level are independent given the
  previous level, so 128-row blocks are dealt block-cyclically over the ranks (equal rows AND equal
  cells of the triangle per rank); after every level the ranks all-gather their value chunks
  (n*w bytes in total: 8 MB at n=1e6) and scatter them into the next level's records.  Split indices stay
  with their owner; the backtrace is T point look-ups, each broadcast from the owning rank.

The engine is injected so the host logic can be exercised with the gloo backend on CPU
(tests/test_dist_gloo.py); :class:`CudaShardEngine` is the real one (C ABI, ps_shard_*).
"""
import ctypes as C

import numpy as np

from . import capi

BR = 128  # rows per block == tile rows of the level kernel


# ------------------------------------------------------------------------------------ replicas
def replica_slice(R, rank, world):
    """Contiguous block of instances owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(R, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def solve_replicas(solve_batch, a, b, group=None):
    """a, b: [R, n] host arrays, identical on every rank.  `solve_batch(a_blk, b_blk)` -> dict of arrays with a
    leading instance dimension.  Returns the concatenated dict on rank 0 (None elsewhere)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = replica_slice(a.shape[0], rank, world)
    mine = solve_batch(a[lo:hi], b[lo:hi]) if hi > lo else {}
    if world == 1:
        return mine
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0, group=group)
    if rank != 0:
        return None
    keys = [k for k in gathered[0].keys()] if gathered[0] else []
    return {k: np.concatenate([g[k] for g in gathered if g]) for k in keys}


# ------------------------------------------------------------------------------------ row sharding
class CudaShardEngine:
    """ps_shard_* over one capi.Context; level vectors live in torch CUDA tensors."""

    def __init__(self, ctx, device):
        import torch
        self.ctx, self.torch, self.device = ctx, torch, device
        self.lib = ctx.lib
        self._h = C.c_void_p()
        self._keep = None

    def sort(self, a, b, dtype):
        """Device radix sort by priority; returns (perm_host, a_sorted_dev, b_sorted_dev)."""
        torch = self.torch
        tdt = torch.float64 if dtype == np.float64 else torch.float32
        a_d = torch.as_tensor(np.ascontiguousarray(a, dtype), device=self.device)
        b_d = torch.as_tensor(np.ascontiguousarray(b, dtype), device=self.device)
        n = a_d.shape[0]
        perm_d = torch.empty(n, dtype=torch.int32, device=self.device)
        self.ctx.wait_torch(self.device)  # the engine's stream is not ordered with torch's: a_d, b_d must have arrived
        # T=1: sort + level 1 only; we only want the sort index, so level 1 takes the closed prefix form instead of
        # walking every row (O(n^2) additions in RUNNING mode: 0.2 s at n = 1e6)
        self.ctx.solve_device(a_d.data_ptr(), b_d.data_ptr(), n, 1, 2, True, dtype, d_perm_out=perm_d.data_ptr(),
                              sum_mode=capi.PS_SUM_PREFIX)
        idx = perm_d.long()
        return perm_d.cpu().numpy(), a_d[idx].contiguous().to(tdt), b_d[idx].contiguous().to(tdt)

    def create(self, a_s, b_s, n, T, dist_id, risk_part, dtype, rank, world, sum_mode=capi.PS_SUM_RUNNING):
        p = capi.PsProblem(n=n, T=T, dist=dist_id, risk_part=int(risk_part), sum_mode=sum_mode, sort_mode=0,
                           perm=None, sweep=0, batch=1, on_device=1, no_fast_math=0, chunk_len=0, screen=0)
        fn = self.lib.ps_shard_create_f64 if dtype == np.float64 else self.lib.ps_shard_create_f32
        self.ctx.wait_torch(self.device)  # a_s, b_s were gathered on torch's stream
        self.ctx._check(fn(self.ctx._h, C.byref(p), C.c_void_p(a_s.data_ptr()), C.c_void_p(b_s.data_ptr()), rank,
                           world, C.byref(self._h)))
        self._keep = (a_s, b_s)
        self.chunk_rows = self.lib.ps_shard_chunk_rows(self._h)
        self.tdt = self.torch.float64 if dtype == np.float64 else self.torch.float32
        return self

    @property
    def screened(self):
        """True when this shard's levels run the screened kernel (csrc/screen.cuh)."""
        return bool(self.lib.ps_shard_screened(self._h))

    def new_chunk(self):
        return self.torch.zeros(self.chunk_rows, dtype=self.tdt, device=self.device)

    def new_gathered(self, world):
        return self.torch.zeros(world * self.chunk_rows, dtype=self.tdt, device=self.device)

    def level1(self, chunk):
        self.ctx.wait_torch(self.device)  # the chunk's zero fill runs on torch's stream
        self.ctx._check(self.lib.ps_shard_level1(self._h, C.c_void_p(chunk.data_ptr())))

    def set_prev(self, gathered):
        self.ctx.wait_torch(self.device)  # the all-gather that filled `gathered` runs on torch's stream
        self.ctx._check(self.lib.ps_shard_set_prev(self._h, C.c_void_p(gathered.data_ptr())))

    def level(self, j, chunk):
        self.ctx._check(self.lib.ps_shard_level(self._h, j, C.c_void_p(chunk.data_ptr())))

    def next_start(self, j, i):
        out = C.c_int(0)
        self.ctx._check(self.lib.ps_shard_next_start(self._h, j, i, C.byref(out)))
        return out.value

    def range_score(self, a_seg, b_seg, dist_id, risk_part, dtype):
        return self.ctx.range_score(a_seg, b_seg, dist_id, risk_part, dtype=dtype)

    def subset_scores(self, bounds, dtype):
        """Scores of the subsets [bounds[t], bounds[t+1]) in one launch (one thread per subset)."""
        bounds = np.ascontiguousarray(bounds, np.int32)
        out = np.empty(len(bounds) - 1, dtype)
        self.ctx._check(self.lib.ps_shard_subset_scores(self._h, len(out), bounds.ctypes.data_as(C.c_void_p),
                                                        out.ctypes.data_as(C.c_void_p)))
        return out

    def close(self):
        if self._h:
            self.lib.ps_shard_destroy(self._h)
            self._h = C.c_void_p()


def owner_of_row(i, world):
    return (i // BR) % world


def natural_from_gathered(gathered, n, world, chunk_rows):
    """[world][chunk] block-cyclic layout -> natural row order (host-side mirror of shard_unpermute_kernel)."""
    g = np.asarray(gathered).reshape(world, chunk_rows)
    out = np.empty(n, g.dtype)
    nblk = (n + BR - 1) // BR
    for rb in range(nblk):
        r, m = rb % world, rb // world
        lo, hi = rb * BR, min(n, rb * BR + BR)
        out[lo:hi] = g[r, m * BR:m * BR + (hi - lo)]
    return out


class ShardedSolver:
    """Row-sharded solve of ONE instance across the ranks of `group`.  Every rank passes the same a, b."""

    def __init__(self, engine, a, b, T, dist_id, risk_part, dtype=np.float64, group=None, keep_levels=False):
        import torch
        import torch.distributed as dist
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        import time
        n = len(a)
        Tc = min(T, n)
        self.n, self.T = n, Tc
        t_start = time.perf_counter()
        perm, a_s, b_s = engine.sort(a, b, dtype)
        self.perm = perm
        eng = engine.create(a_s, b_s, n, Tc, dist_id, risk_part, dtype, self.rank, self.world)
        chunk = eng.new_chunk()
        gathered = eng.new_gathered(self.world)
        self.levels = []

        def exchange():
            if self.world > 1:
                dist.all_gather_into_tensor(gathered, chunk, group=group)
            else:
                gathered.copy_(chunk)
            if gathered.is_cuda:
                # the engine runs on its own stream: finish the collective before the next engine call reads it
                torch.cuda.current_stream(gathered.device).synchronize()
            if keep_levels:
                self.levels.append(natural_from_gathered(gathered.cpu().numpy(), n, self.world, eng.chunk_rows))

        t_setup = time.perf_counter()
        eng.level1(chunk)
        exchange()
        t_l1 = time.perf_counter()
        for j in range(2, Tc + 1):
            eng.set_prev(gathered)
            eng.level(j, chunk)
            exchange()
        t_levels = time.perf_counter()
        # backtrace: T point look-ups, each answered by the row's owner (DP_impl.hpp:131-139)
        bounds = [0]
        cur = 0
        for t in range(Tc, 0, -1):
            if t == 1:
                nxt = n  # nextStart_[1][i] = n for every i (DP_impl.hpp:93)
            else:
                own = owner_of_row(cur, self.world)
                val = torch.zeros(1, dtype=torch.int64)
                if own == self.rank:
                    val[0] = eng.next_start(t, cur)
                if self.world > 1:
                    buf = val.to(gathered.device) if gathered.is_cuda else val
                    dist.broadcast(buf, src=own, group=group)
                    val = buf.cpu()
                nxt = int(val[0])
            bounds.append(nxt)
            cur = nxt
        self.bounds = np.array(bounds, np.int32)
        # objective value = maxScore_[T][0], held by rank 0's chunk slot 0 after the last exchange
        self.max_score_T0 = float(gathered[0].item())
        if self.rank != 0:
            self.subset_scores = None
        elif hasattr(eng, "subset_scores"):
            self.subset_scores = eng.subset_scores(self.bounds, dtype)
        else:
            a_host = np.asarray(a, dtype)[perm]
            b_host = np.asarray(b, dtype)[perm]
            self.subset_scores = np.array(
                [engine.range_score(a_host[lo:hi], b_host[lo:hi], dist_id, risk_part, dtype)
                 for lo, hi in zip(bounds[:-1], bounds[1:])], dtype)
        self.subsets = [perm[lo:hi] for lo, hi in zip(bounds[:-1], bounds[1:])]
        eng.close()
        t_end = time.perf_counter()
        # host wall-clock per phase on this rank (every engine call blocks on its stream)
        self.seconds = {"sort_and_setup": t_setup - t_start, "level1": t_l1 - t_setup,
                        "levels_with_allgather": t_levels - t_l1, "backtrace": t_end - t_levels}
