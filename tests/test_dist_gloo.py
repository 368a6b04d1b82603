"""CPU tests of the N>1 host logic with the gloo backend (world_size 2): the block-cyclic row-sharded driver
(partitionsolvers_b200/multigpu.ShardedSolver) and the replica sharding, with a numpy/oracle stand-in for the
CUDA engine so the exchange pattern, chunk layout, owner look-ups and backtrace can run without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from partitionsolvers_b200 import multigpu, synth

BR = multigpu.BR


class FakeShardEngine:
    """Same interface as multigpu.CudaShardEngine; arithmetic = the CPU oracle (tests may use it)."""

    def __init__(self):
        from oracle import oracle_py as O
        self.O = O

    def sort(self, a, b, dtype):
        a = np.asarray(a, dtype)
        b = np.asarray(b, dtype)
        perm = np.argsort((a / b).astype(dtype), kind="stable").astype(np.int32)
        return perm, torch.from_numpy(a[perm].copy()), torch.from_numpy(b[perm].copy())

    def create(self, a_s, b_s, n, T, dist_id, risk_part, dtype, rank, world, sum_mode=0):
        self.a, self.b = a_s.numpy(), b_s.numpy()
        self.n, self.T, self.dist, self.rp, self.dtype = n, T, dist_id, risk_part, dtype
        self.rank, self.world = rank, world
        nblk = (n + BR - 1) // BR
        self.chunk_rows = ((nblk + world - 1) // world) * BR
        self.owned = [i for rb in range(rank, nblk, world) for i in range(rb * BR, min(n, rb * BR + BR))]
        self.slot = {i: (i // BR // world) * BR + i % BR for i in self.owned}
        self.ns = {}
        self.prev = None
        return self

    def new_chunk(self):
        return torch.zeros(self.chunk_rows, dtype=torch.float64 if self.dtype == np.float64 else torch.float32)

    def new_gathered(self, world):
        return torch.zeros(world * self.chunk_rows, dtype=torch.float64 if self.dtype == np.float64 else torch.float32)

    def level1(self, chunk):
        c = chunk.numpy()
        for i in self.owned:
            c[self.slot[i]] = self.O.range_score(self.a, self.b, self.dist, self.rp, i, self.n, dtype=self.dtype)

    def set_prev(self, gathered):
        self.prev = multigpu.natural_from_gathered(gathered.numpy(), self.n, self.world, self.chunk_rows)

    def level(self, j, chunk):
        kmax = self.n - j + 1
        nrows = 1 if j == self.T else kmax
        rows = np.array([i for i in self.owned if i < nrows], np.int32)
        c = chunk.numpy()
        c[:] = 0
        if len(rows):
            best, arg = self.O.dp_rows(self.a, self.b, self.prev, kmax, self.dist, self.rp, rows, dtype=self.dtype)
            for i, v, k in zip(rows, best, arg):
                c[self.slot[int(i)]] = v
                self.ns[(j, int(i))] = int(k)

    def next_start(self, j, i):
        return self.ns[(j, i)]

    def range_score(self, a_seg, b_seg, dist_id, risk_part, dtype):
        return self.O.range_score(a_seg, b_seg, dist_id, risk_part, 0, len(a_seg), dtype=dtype)

    def close(self):
        pass


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, T, dist_id, rp, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = synth.gaussian_mixture(n, 5, seed=77, dtype=np.float64)
        if dist_id == 1:
            a = np.abs(a) + 0.5
        s = multigpu.ShardedSolver(FakeShardEngine(), a, b, T, dist_id, rp, dtype=np.float64, keep_levels=True)
        # replicas: 5 instances over 2 ranks
        ar, br = synth.poisson_null(300, seed=5, dtype=np.float64, replicas=5)

        def solve_batch(a_blk, b_blk):
            from oracle import oracle_py as O
            res = [O.solve("port", a_blk[r], b_blk[r], 4, 1, True, dtype=np.float64, levels=False) for r in
                   range(a_blk.shape[0])]
            return {"score": np.array([float(x["score"]) for x in res])}
        rep = multigpu.solve_replicas(solve_batch, ar, br)
        if rank == 0:
            q.put((s.bounds, s.perm, s.max_score_T0, s.subset_scores, np.stack(s.levels), rep["score"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dist_id,rp", [(2, False), (1, True)])
def test_row_sharded_driver_and_replicas_world2(oracle, dist_id, rp):
    n, T, world = 700, 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, T, dist_id, rp, q)) for r in range(world)]
    for p in procs:
        p.start()
    bounds, perm, top, sscores, levels, rep_scores = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, b = synth.gaussian_mixture(n, 5, seed=77, dtype=np.float64)
    if dist_id == 1:
        a = np.abs(a) + 0.5
    ref = oracle.solve("port", a, b, T, dist_id, rp, dtype=np.float64, perm_in=perm)
    # levels 1..T in natural order equal the unsharded oracle rows (written cells)
    for j in range(1, T + 1):
        nrows = n if j == 1 else (1 if j == T else n - j + 1)
        assert np.array_equal(levels[j - 1][:nrows], ref["max_score"][j][:nrows])
    # boundaries == walking the oracle's nextStart_
    cur, want = 0, [0]
    for t in range(T, 0, -1):
        cur = int(ref["next_start"][t][cur])
        want.append(cur)
    assert bounds.tolist() == want
    assert top == float(ref["max_score"][T][0])
    assert abs(float(np.sum(sscores)) - float(ref["optimal_score_member"])) <= 1e-9 * abs(float(ref["optimal_score_member"]))
    # replicas: gathered on rank 0 in instance order
    ar, br = synth.poisson_null(300, seed=5, dtype=np.float64, replicas=5)
    want_rep = [float(oracle.solve("port", ar[r], br[r], 4, 1, True, dtype=np.float64, levels=False)["score"])
                for r in range(5)]
    assert rep_scores.tolist() == want_rep


def test_replica_slices_cover_everything():
    for R in (1, 7, 4096):
        for world in (1, 2, 3, 8):
            spans = [multigpu.replica_slice(R, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == R
            assert all(x[1] == y[0] for x, y in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_block_cyclic_layout_roundtrip():
    n, world = 1000, 3
    nblk = (n + BR - 1) // BR
    chunk = ((nblk + world - 1) // world) * BR
    g = np.full((world, chunk), -1.0)
    for i in range(n):
        rb = i // BR
        g[rb % world, (rb // world) * BR + i % BR] = i
    nat = multigpu.natural_from_gathered(g.reshape(-1), n, world, chunk)
    assert np.array_equal(nat, np.arange(n, dtype=float))
    assert [multigpu.owner_of_row(i, world) for i in (0, 127, 128, 256, 384)] == [0, 0, 1, 2, 0]
