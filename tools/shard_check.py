"""Row-sharded solve of ONE instance through the C ABI (ps_solve_sharded_*), one process per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/shard_check.py [--size 1000000] [--parts 32] [--dtype f64] [--check] [--reps 3]

The unique id of the communicator travels over a gloo process group (any transport would do: the library only needs the 128
bytes); NCCL itself is used by the library alone.  --check compares every rank's result with the single-GPU solve."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partitionsolvers_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=1000000)
    ap.add_argument("--parts", dest="T", type=int, default=32)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--data", default="counts", choices=["counts", "real"])
    ap.add_argument("--screen", type=int, default=0)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    ctx = capi.Context(local)
    ids = [ctx.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(ids, src=0)
    comm = ctx.comm_create(ids[0], rank, world)
    dt = np.float64 if args.dtype == "f64" else np.float32
    if args.data == "counts":
        a, b = synth.poisson_planted(args.n, args.T, seed=synth.SEED + 5, dtype=dt)
        dist_id = 1
    else:  # real-valued mixture, Rational score: sums round, the levels run on the reference's running sums
        a, b = synth.gaussian_mixture(args.n, args.T, seed=synth.SEED + 5, dtype=dt)
        dist_id = 2
    res, secs = None, []
    for _ in range(args.reps):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        res = ctx.solve_sharded(comm, a, b, args.T, dist_id, False, dtype=dt, screen=args.screen)
        secs.append(time.perf_counter() - t0)
    tm = ctx.timings()
    ok = True
    if args.check:
        one = ctx.solve_raw(a, b, args.T, dist_id, False, dtype=dt, screen=args.screen)
        ok = all(np.array_equal(np.asarray(res[k]).ravel(), np.asarray(one[k]).ravel()) for k in ("perm", "bounds", "subset_scores"))
    print(f"rank {rank}/{world}: n={args.n} T={args.T} {args.dtype} seconds {['%.4f' % s for s in secs]} device_ms {tm['total_ms']:.2f} "
          f"levels_ms {tm['levels_ms']:.2f} sort_ms {tm['sort_scan_ms']:.2f} screen {tm['screen_used']} comm_bytes {tm['comm_bytes']} "
          f"same_as_one_gpu {ok if args.check else 'n/a'}", flush=True)
    oks = [ok]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, ok)
        oks = gathered
    ctx.comm_destroy(comm)
    if rank == 0 and all(oks):
        print("SHARD_CHECK_OK", flush=True)
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if all(oks) else 1)


if __name__ == "__main__":
    main()
