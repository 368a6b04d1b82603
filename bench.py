#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on its headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype f64|f32]

A "step" is ONE complete consecutive-partitions solve (priority sort, records, level 1, levels 2..T,
backtrace) of the headline workload: synthetic n=100000, T=16 Poisson multiple clustering
(BASELINE.json configs[2], synth.CONFIGS["C3"]).  metric = DP cell-updates/s = n^2*T / time (nominal
count, BASELINE.json); `actual_cells_per_s` is the exact (i,k) count / time.

  value   inputs resident in HBM, timed on the device (CUDA events on the library's stream)
  e2e     same solve through the host-buffer C-ABI call: pinned host a,b in, sort index + boundaries +
          subset scores out, copies inside the timed region (wall clock around the blocking call)
  N > 1   one process per GPU (torchrun); every rank solves its own instance of the workload
          (independent instances shard with no collective: SURVEY.md 8e) -> weak scaling,
          value = N * n^2*T / max-over-ranks time.  The replica workload (configs[3]) is reported in
          "mc_replicas" (4096 instances split over the ranks).

Extras on the same JSON line (skipped by --no-extras): other_level_kernel (the exhaustive kernel on the headline workload),
variants (other precision, other sum mode), c3_all_scores (every score x objective x precision at n=100000, T=16),
mc_replicas + cpu_replicas_baseline (configs[3] next to the reference's ThreadPool variant), ols_sweep, null_table_paper_scale,
small_configs (configs[0], [1]), sharded (N > 1: configs[4], rows sharded with an NCCL all-gather per level), cpu_baseline,
cpu_baseline_all_cores.  stdout carries exactly one line.

--impl reference times the reference's own CPU solver (oracle/_ref = the unmodified reference headers;
the port only if that library is absent) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from partitionsolvers_b200 import synth  # noqa: E402

METRIC = "dp_cell_updates_per_s"
UNIT = "cells/s"
ALG_FLOPS_PER_CELL = {  # SURVEY.md 8d, (dist, risk_part) -> algorithmic flops per actual (i,k) cell
    (1, False): 10, (1, True): 7, (0, False): 10, (0, True): 7, (2, False): 6, (2, True): 6}
ALG_HBM_BYTES_PER_ROW = lambda w: 4 * w + 4  # noqa: E731  read prev,x,y; write value + int split (per level row)


def dist_env():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """nvidia-smi style clock / throttle-reason sampling DURING the timed region (pynvml)."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else \
                    self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.ok:
            self.t.start()

    def stop(self):
        self._stop.set()
        if self.ok and self.t.is_alive():
            self.t.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU baselines
def cpu_reference_sample(dtype, target_s, quiet=True, instances=1):
    """Time the reference's own solver on a bounded sample of the headline workload: same score,
    objective, T and input distribution, smaller n (its three n x n tables do not fit at n=100k)."""
    from oracle import oracle_py as O
    cfg = synth.CONFIGS["C3"]
    kind = "reference" if O.have("ref") else "port"
    if kind == "port":
        O.build(("port",))
    which = "ref" if kind == "reference" else "port"
    n_cal = 3000
    a, b = cfg["gen"](dtype, n_cal)
    t_cal, _ = O.time_solve(which, a, b, cfg["T"], cfg["dist"], cfg["risk_part"], dtype=dtype)
    w = np.dtype(dtype).itemsize
    try:
        import psutil
        ram = psutil.virtual_memory().available
    except Exception:
        ram = 32 << 30
    # the reference allocates three (n+1)^2 tables (score_impl.hpp:238-239): memory, not a fixed cap, bounds the sample
    n_mem = int(np.sqrt(0.6 * ram / (3 * w) / max(1, instances))) if kind == "reference" else 10 ** 9
    n_s = int(n_cal * np.sqrt(max(target_s, t_cal) / max(t_cal, 1e-4)))
    n_s = max(n_cal, min(n_s, n_mem, 100000))
    n_s = (n_s // 500) * 500
    return which, kind, n_s


def run_cpu_solver(which, dtype, n_s):
    from oracle import oracle_py as O
    cfg = synth.CONFIGS["C3"]
    a, b = cfg["gen"](dtype, n_s)
    t, score = O.time_solve(which, a, b, cfg["T"], cfg["dist"], cfg["risk_part"], dtype=dtype)
    return t, score


def hostile_inputs(kind, n, dt):
    """Inputs on which exact branch-and-bound cannot prune (tests/test_gpu_ties.py uses the same generators)."""
    rs = np.random.RandomState(77)
    if kind == "constant_ratio":  # a = 3 b: every priority ties and every cell of a row ties in exact arithmetic
        b = rs.randint(1, 200, n).astype(np.float64)
        return (3.0 * b).astype(dt), b.astype(dt)
    if kind == "one_two":  # gtest_all.cpp:924-958: a in {1, 2}, b = 1
        return rs.randint(1, 3, n).astype(dt), np.ones(n, dt)
    if kind == "uniform_counts":  # no planted structure
        return rs.randint(1, 1000, n).astype(dt), rs.randint(1, 1000, n).astype(dt)
    raise ValueError(kind)


_FULL_REF_SNIPPET = r'''
import json, sys, time
import numpy as np
sys.path.insert(0, %(root)r)
from oracle import oracle_py as O
from partitionsolvers_b200 import synth
cfg = synth.CONFIGS["C3"]
a, b = cfg["gen"](np.float32)
t, score = O.time_solve("ref", a, b, cfg["T"], cfg["dist"], cfg["risk_part"], dtype=np.float32)
print(json.dumps({"seconds": t, "score": score}))
'''


def cpu_reference_full_size(limit_s=330.0):
    """SAME CONFIGURATION leg: the unmodified reference DPSolver<float> at n = 100000, T = 16 (its three (n+1)^2 float
    tables need 120 GB; the f64 tables, 240 GB, do not fit this pool's 196 GB boxes).  Runs in a child process with a hard
    time limit; skipped with a stated reason where memory or the predicted time does not allow it."""
    from oracle import oracle_py as O
    cfg = synth.CONFIGS["C3"]
    n = cfg["n"]
    need = 3.0 * (n + 1) ** 2 * 4
    if os.environ.get("PS_BENCH_SKIP_FULL_REF"):
        return {"ran": False, "why": "PS_BENCH_SKIP_FULL_REF set"}
    if not O.have("ref"):
        return {"ran": False, "why": "oracle/_ref/libps_ref.so absent (the reference was not compiled into this tree)"}
    try:
        import psutil
        avail = float(psutil.virtual_memory().available)
    except Exception:
        avail = 0.0
    if avail < 1.12 * need:
        return {"ran": False, "why": f"needs {need / 1e9:.0f} GB of tables, {avail / 1e9:.0f} GB of host memory available"}
    n_cal = 6000
    a, b = cfg["gen"](np.float32, n_cal)
    t_cal, _ = O.time_solve("ref", a, b, cfg["T"], cfg["dist"], cfg["risk_part"], dtype=np.float32)
    predicted = t_cal * (n / n_cal) ** 2
    if predicted > 0.9 * limit_s:
        return {"ran": False, "why": f"predicted {predicted:.0f} s from n={n_cal} ({t_cal:.2f} s) exceeds the {limit_s:.0f} s limit"}
    try:
        out = subprocess.run([sys.executable, "-c", _FULL_REF_SNIPPET % {"root": ROOT}], capture_output=True, text=True,
                             timeout=limit_s)
        rec = json.loads(out.stdout.strip().splitlines()[-1])
    except subprocess.TimeoutExpired:
        return {"ran": False, "why": f"did not finish within {limit_s:.0f} s (predicted {predicted:.0f} s)"}
    except Exception as e:
        return {"ran": False, "why": f"child failed: {e}"}
    return {"ran": True, "seconds": rec["seconds"], "score": rec["score"], "predicted_s": predicted,
            "value": synth.nominal_cells(n, cfg["T"]) / rec["seconds"], "n": n, "T": cfg["T"], "dtype": "f32",
            "table_bytes": need}


def cpu_threads_used(kind):
    # reference: <= 9 fill threads (score_impl.hpp:10-12,241,258-267) + single-threaded DP loop
    hw = os.cpu_count() or 1
    return min(9, hw) if kind == "reference" else hw


def reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    dtype = np.float64 if args.dtype == "f64" else np.float32
    cfg = synth.CONFIGS["C3"]
    budget = 150.0
    per_step = max(1.0, min(8.0, budget / max(1, args.steps + args.warmup)))
    # --gpus N: the GPU arm solves N instances at once (one per GPU, weak scaling), so this arm solves N instances at once
    # on the box's host cores (N processes, each the reference's own <= 9 threads) -- the same job on the same box
    inst = max(1, args.gpus)
    which, kind, n_s = cpu_reference_sample(dtype, per_step, instances=inst)

    pool = None
    if inst > 1:
        import multiprocessing as mp
        pool = mp.get_context("spawn").Pool(inst)  # fresh interpreters: the reference's library is loaded per worker

    def one_step():
        if pool is None:
            return run_cpu_solver(which, dtype, n_s)[0]
        # the N solves start together; the step lasts as long as the slowest of them
        return max(t for t, _ in pool.starmap(run_cpu_solver, [(which, dtype, n_s)] * inst))

    for _ in range(max(args.warmup, 1 if pool else 0)):
        one_step()
    times = [one_step() for _ in range(args.steps)]
    if pool is not None:
        pool.close()
        pool.join()
    ms = 1e3 * float(np.mean(times))
    value = inst * synth.nominal_cells(n_s, cfg["T"]) / (ms * 1e-3)
    sample = (f"{'unmodified reference DPSolver' if kind == 'reference' else 'oracle port'}"
              f"<{'double' if dtype == np.float64 else 'float'}> constructor at n={n_s}, T={cfg['T']} "
              f"(same generator as the headline workload; n=100000 needs {3 * 1e10 * np.dtype(dtype).itemsize / 1e9:.0f} GB of tables)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "C3 sample: Poisson multiple clustering, T=16", "n": n_s, "T": cfg["T"],
                       "host_cores": os.cpu_count(), "concurrent_instances": inst,
                       "same_config_as_gpu_arm": False,
                       "why_not": "the reference needs 3 (n+1)^2 tables: 240 GB (f64) / 120 GB (f32) at n=100000; one f32 solve at "
                                  "full size is timed by the GPU arm's cpu_baseline leg where the box has the memory"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": min(os.cpu_count() or 1, inst * cpu_threads_used(kind)),
                             "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------- our arm
def measured_peaks():
    """FP32/FP64/MUFU pipe peaks measured on this device by tools/peaks.cu (built by build())."""
    exe = os.path.join(ROOT, "partitionsolvers_b200", "_lib", "ps_peaks")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=120).stdout
        return json.loads(out.strip().splitlines()[-1])
    except Exception:
        return None


_RECORD_FD = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_RECORD_FD if _RECORD_FD is not None else 1, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--sum-mode", default="running", choices=["running", "prefix"])
    ap.add_argument("--no-extras", action="store_true", help="skip variants, replicas and the CPU baseline")
    ap.add_argument("--skip-peaks", action="store_true", help="do not run the pipe-peak microbenchmark (ncu runs)")
    ap.add_argument("--exhaustive", action="store_true",
                    help="headline on the exhaustive level kernel (every cell evaluated with the reference's arithmetic) "
                         "instead of the default screened kernel; both produce identical tables")
    ap.add_argument("--chunk-len", type=int, default=0, help="tuning: k-chunk length of the level tiles (0 = library default)")
    ap.add_argument("--sharded-n", type=int, default=1000000, help="N>1 only: size of the row-sharded instance")
    ap.add_argument("--sharded-T", type=int, default=32)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly ONE line, the JSON record: anything libraries write to file descriptor 1 on the way (NCCL prints
    # its version banner there) is sent to stderr, the record goes to the saved descriptor at the end
    sys.stdout.flush()
    global _RECORD_FD
    _RECORD_FD = os.dup(1)
    os.dup2(2, 1)

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    from partitionsolvers_b200 import capi

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version there)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    cfg = synth.CONFIGS["C3"]
    n, T, dist_id, rp = cfg["n"], cfg["T"], cfg["dist"], cfg["risk_part"]
    sum_mode = capi.PS_SUM_RUNNING if args.sum_mode == "running" else capi.PS_SUM_PREFIX
    ctx = capi.Context(local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def l2_flush():
        flush.add_(1)
        torch.cuda.synchronize()

    def run_variant(dtype_name, steps, warmup, with_e2e=True, sm=sum_mode, screen=0):
        npdt = np.float64 if dtype_name == "f64" else np.float32
        tdt = torch.float64 if dtype_name == "f64" else torch.float32
        # every rank solves its own instance of the workload (rank-dependent seed)
        a_h, b_h = synth.poisson_planted(n, 16, seed=synth.SEED + 3 + 1000 * rank, dtype=npdt)
        a_pin = torch.from_numpy(a_h).pin_memory()
        b_pin = torch.from_numpy(b_h).pin_memory()
        a_d, b_d = a_pin.to(dev), b_pin.to(dev)
        perm_d = torch.empty(n, dtype=torch.int32, device=dev)
        bounds_d = torch.empty(T + 1, dtype=torch.int32, device=dev)
        scores_d = torch.empty(T, dtype=tdt, device=dev)
        perm_pin = torch.empty(n, dtype=torch.int32).pin_memory()
        bounds_pin = torch.empty(T + 1, dtype=torch.int32).pin_memory()
        scores_pin = torch.empty(T, dtype=tdt).pin_memory()

        def step_resident():
            ctx.solve_device(a_d.data_ptr(), b_d.data_ptr(), n, T, dist_id, rp, npdt, sum_mode=sm,
                             d_perm_out=perm_d.data_ptr(), d_bounds=bounds_d.data_ptr(),
                             d_scores=scores_d.data_ptr(), chunk_len=args.chunk_len, screen=screen)

        def step_e2e():
            ctx.solve_into(a_pin.numpy(), b_pin.numpy(), T, dist_id, rp, npdt, perm_out=perm_pin.numpy(),
                           bounds_out=bounds_pin.numpy(), scores_out=scores_pin.numpy(), sum_mode=sm,
                           chunk_len=args.chunk_len, screen=screen)

        for _ in range(warmup):
            step_resident()
        # ---- resident, device-timed
        dev_ms, lev_ms, lev_cells, launches, exact_cells, screened = [], [], [], 0, 0, 0
        stage = {"sort_scan_ms": 0.0, "prepass_ms": 0.0, "levels_ms": 0.0, "backtrace_ms": 0.0}
        barrier()
        w0 = time.perf_counter()
        for _ in range(steps):
            l2_flush()
            step_resident()
            tm = ctx.timings()
            dev_ms.append(tm["total_ms"])
            lev_ms += tm["level_ms"]
            lev_cells += tm["level_cells"]
            launches += tm["kernel_launches"]
            exact_cells += tm["screen_exact_cells"]
            screened = tm["screen_used"]
            for k in stage:
                stage[k] += tm[k] / steps
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - w0) / steps
        ms = allmax(float(np.mean(dev_ms)))
        out = {"ms_per_step": ms, "wall_ms_per_step": allmax(wall_ms), "launches": launches,
               "value": world * synth.nominal_cells(n, T) / (ms * 1e-3),
               "actual_cells_per_s": world * synth.actual_cells(n, T) / (ms * 1e-3),
               "level_ms_sum": float(np.sum(lev_ms)), "level_cells_sum": float(np.sum(lev_cells)),
               "level_launches": len(lev_ms), "stage_ms": stage, "screen_used": int(screened),
               "exact_cells": float(exact_cells), "fallback_level": int(tm["screen_fallback_level"]),
               "bounds": bounds_d.cpu().numpy().tolist()}
        # fraction of (i,k) cells past the A = B frontier (only those evaluate division + log in the mult-clust
        # scores, exactly like the reference's ternary): Monte-Carlo estimate on the sorted inputs
        pm = perm_d.cpu().numpy()
        Pa = np.concatenate([[0.0], np.cumsum(a_h[pm].astype(np.float64))])
        Pb = np.concatenate([[0.0], np.cumsum(b_h[pm].astype(np.float64))])
        rs_ = np.random.RandomState(7)
        ii, kk = rs_.randint(0, n, 1000000), rs_.randint(1, n + 1, 1000000)
        mm = kk > ii
        out["frac_cells_A_gt_B"] = float(np.mean((Pa[kk[mm]] - Pa[ii[mm]]) > (Pb[kk[mm]] - Pb[ii[mm]])))
        if with_e2e:
            for _ in range(2):
                step_e2e()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                l2_flush()
                step_e2e()
            barrier()
            e2e_ms = allmax(1e3 * (time.perf_counter() - t0) / steps)
            tm = ctx.timings()
            out["e2e"] = {"value": world * synth.nominal_cells(n, T) / (e2e_ms * 1e-3), "unit": UNIT,
                          "h2d_bytes_per_step": int(tm["h2d_bytes"]), "d2h_bytes_per_step": int(tm["d2h_bytes"]),
                          "ms_per_step": e2e_ms}
            assert np.array_equal(bounds_pin.numpy(), np.array(out["bounds"], np.int32)), "e2e and resident differ"
        return out

    sampler = ClockSampler(local_rank)
    sampler.start()
    main_res = run_variant(args.dtype, args.steps, args.warmup, screen=1 if args.exhaustive else 0)
    clocks = sampler.stop()

    line = {"metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "C3: n=100000, T=16 Poisson multiple clustering, one instance per GPU",
                       "n": n, "T": T, "score": "Poisson", "objective": "multiple_clustering",
                       "sum_mode": args.sum_mode, "sort": "device radix (stable)",
                       "l2": "flushed between timed steps (256 MB write)",
                       "parallelism": f"{world} independent instance(s), no collective"},
            "actual_cells_per_s": main_res["actual_cells_per_s"],
            "wall_ms_per_step": main_res["wall_ms_per_step"], "stage_ms": main_res["stage_ms"],
            "e2e": main_res["e2e"], "gpu_launches": int(main_res["launches"]), "clocks": clocks}

    peaks = measured_peaks() if (rank == 0 and not args.skip_peaks) else None
    w = 8 if args.dtype == "f64" else 4
    flops_full = ALG_FLOPS_PER_CELL[(dist_id, rp)]
    hbm_peak = 6554.6
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "level_kernel_summary.json")))
    except Exception:
        prof = {}

    try:
        level_instr = json.load(open(os.path.join(ROOT, "profiles", "level_instr_r2.json")))
    except Exception:
        level_instr = {}

    def roofline_of(res, dtype_name):
        """Roofline object(s) of the level kernels a run used (> 90 % of a step).

        Exhaustive kernel (every cell evaluated with the reference's arithmetic): the classical one -- achieved =
        ALGORITHMIC flops of the level launches / their CUDA-event time against the measured non-fused add/mul rate of the
        data type (parity forbids FMA contraction of the score).  The count charges every (i,k) cell: 10 flops for a Poisson
        mult-clust cell with A > B, 5 for a cell with A <= B whose score is 0 by definition (score_impl.hpp:385-388).

        Screened path (exact branch-and-bound): the same algorithmic count divided by its time EXCEEDS the pipe peak --
        it evaluates ~1 % of the cells and dismisses the rest in ranges, so that ratio measures pruning, not the kernel.
        What is reported as `frac` for it is instruction-issue utilisation (executed warp instructions of the four level
        launches, counted per level by ncu on this very workload: profiles/level_instr_r2.json) with the algorithmic ratio
        kept beside it under `algorithmic` and labelled for what it is."""
        screened = bool(res["screen_used"])
        p_gt = res["frac_cells_A_gt_B"] if (not rp and dist_id != 2) else 1.0
        flops_cell = 5.0 + (flops_full - 5.0) * p_gt
        lev_s = res["level_ms_sum"] * 1e-3
        achieved = res["level_cells_sum"] * flops_cell / lev_s / 1e12
        pipe64 = (dtype_name == "f64")
        if peaks:
            peak = (peaks["dadd_ops"] if pipe64 else peaks["fadd_ops"]) / 1e12
            peak_src = ("measured by tools/peaks.cu on this device: non-fused FP add/mul issue rate "
                        "(parity forbids FMA contraction of the score)")
        else:
            peak = (148 * 64 if pipe64 else 148 * 128) * 1.965e9 / 1e12
            peak_src = "fallback: lanes x 1.965 GHz (ps_peaks binary missing)"
        pk = prof.get(("live_" if screened else "") + dtype_name, {})
        launches = max(1, res["level_launches"])
        avg_launch_ms = res["level_ms_sum"] / launches
        alg_bytes_per_launch = ALG_HBM_BYTES_PER_ROW(8 if dtype_name == "f64" else 4) * n
        mufu_peak = peaks["mufu_lg2_ops"] if peaks else 148 * 16 * 1.965e9
        pipes = {k: pk.get("metrics", {}).get(k, {}).get("value") for k in (
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active")}
        classical = {"kernel": "level_tile_kernel (exhaustive)", "bound": "fp64_pipe" if pipe64 else "fp32_issue",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": pk.get("dram_bytes_per_launch"),
                     "alg_flops_per_cell": flops_cell, "alg_flops_per_evaluated_cell": flops_full,
                     "frac_cells_A_gt_B": p_gt, "cells_per_launch_avg": res["level_cells_sum"] / launches,
                     "avg_launch_ms": avg_launch_ms, "peak_source": peak_src,
                     "share_of_step": res["level_ms_sum"] / (res["steps"] * res["ms_per_step"]),
                     "ncu_pipe_pct": pipes,
                     "hbm": {"alg_bytes_per_launch": alg_bytes_per_launch,
                             "achieved_gbs": alg_bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                             "note": "not HBM-bound: algorithmic traffic is (4w+4)n bytes per level"}}
        if not screened:
            classical["sfu"] = {"ops_per_cell": 1,
                                "what": "one MUFU reciprocal per evaluated cell (division seed); log is table-driven, no MUFU",
                                "achieved_tops": res["level_cells_sum"] / lev_s / 1e12, "peak_tops": mufu_peak / 1e12,
                                "frac": (res["level_cells_sum"] / lev_s) / mufu_peak}
            return classical
        r = {"kernel": "level_live_screen_kernel (+ screen_lb / screen_select / screen_fold of the same level)",
             "bound": "issue", "unit": "T warp-instructions/s",
             "cells_per_launch_avg": res["level_cells_sum"] / launches, "avg_launch_ms": avg_launch_ms,
             "share_of_step": res["level_ms_sum"] / (res["steps"] * res["ms_per_step"]),
             "traffic": pk.get("dram_bytes_per_launch"), "ncu_pipe_pct": pipes,
             "hbm": classical["hbm"],
             "algorithmic": {"achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                             "peak_source": peak_src,
                             "note": "all (i,k) cells charged although ~99 % are dismissed by bounds: a measure of pruning "
                                     "(how far beyond ANY kernel that evaluates every cell), not a pipe utilisation"}}
        li = level_instr.get(dtype_name)
        issue_peak = 148 * 4 * 1.965e9 / 1e12
        if li and li.get("n") == n and li.get("T") == T:
            instr_per_solve = float(sum(sum(v) for v in li["per_level"].values()))
            issue_ach = instr_per_solve * res["steps"] / lev_s / 1e12
            r.update({"achieved": issue_ach, "peak": issue_peak, "frac": issue_ach / issue_peak,
                      "issue_utilisation": issue_ach / issue_peak,
                      "warp_instructions_per_cell": instr_per_solve * res["steps"] / (res["level_cells_sum"] / 32.0),
                      "warp_instructions_per_solve": instr_per_solve,
                      "peak_source": "148 SMs x 4 sub-partitions x 1.965 GHz, one warp instruction per cycle each; executed "
                                     "instructions per level from ncu on this workload (profiles/level_instr_r2.json), time from "
                                     "this run's CUDA events"})
        else:
            r.update({"achieved": None, "peak": issue_peak, "frac": None, "issue_utilisation": None,
                      "peak_source": "profiles/level_instr_r2.json has no entry for this workload"})
        r["screen"] = {"exact_cell_fraction": res["exact_cells"] / max(1.0, res["level_cells_sum"]),
                       "fallback_level": res.get("fallback_level", 0),
                       "note": "exact branch-and-bound: cells are bounded in float, singly or 8 / 128 / a tile at a time where "
                               "the score is monotone in k; only cells that can still be the row maximum are evaluated with "
                               "the reference's arithmetic (identical tables).  Data-dependent: see `pruning`"}
        return r

    main_res["steps"] = args.steps
    line["roofline"] = roofline_of(main_res, args.dtype)
    line["roofline"]["measured_pipe_peaks"] = peaks
    line["config"]["level_kernel"] = ("screened (float bounds, exact arithmetic for surviving cells; identical tables)"
                                      if main_res["screen_used"] else "exhaustive")

    if not args.no_extras:
        # the same workload on the other level kernel (identical tables): exhaustive = every cell evaluated with the
        # reference's arithmetic, the kernel whose pipe roofline is the classical one
        ex_steps = max(2, args.steps // 2)
        vx = run_variant(args.dtype, ex_steps, 3, with_e2e=True, screen=0 if args.exhaustive else 1)
        vx["steps"] = ex_steps
        line["other_level_kernel"] = {"value": vx["value"], "ms_per_step": vx["ms_per_step"], "e2e": vx["e2e"],
                                      "same_partition": vx["bounds"] == main_res["bounds"],
                                      "roofline": roofline_of(vx, args.dtype)}
        # the classical roofline of the path: the kernel that evaluates EVERY cell with the reference's arithmetic
        ex_res = main_res if args.exhaustive else vx
        line["roofline_exhaustive"] = dict(roofline_of(ex_res, args.dtype), ms_per_step=ex_res["ms_per_step"],
                                           value=ex_res["value"])
        other = "f32" if args.dtype == "f64" else "f64"
        v = run_variant(other, max(2, args.steps // 2), 3, with_e2e=True)
        line["variants"] = {other: {"value": v["value"], "ms_per_step": v["ms_per_step"], "e2e": v["e2e"],
                                    "actual_cells_per_s": v["actual_cells_per_s"]}}
        vp = run_variant(args.dtype, 2, 3, with_e2e=False,
                         sm=capi.PS_SUM_PREFIX if sum_mode == capi.PS_SUM_RUNNING else capi.PS_SUM_RUNNING)
        line["variants"]["other_sum_mode"] = {"value": vp["value"], "ms_per_step": vp["ms_per_step"],
                                              "same_partition": vp["bounds"] == main_res["bounds"]}
        # ---- replica workload (BASELINE configs[3]): 4096 x (n=5000, T=8) split over the ranks
        c4 = synth.CONFIGS["C4"]
        R_total = c4["replicas"]
        R_loc = R_total // world + (1 if rank < R_total % world else 0)
        a4, b4 = synth.poisson_null(c4["n"], seed=synth.SEED + 4 + 1000 * rank, dtype=np.float32, replicas=R_loc)
        a4p, b4p = torch.from_numpy(a4).pin_memory(), torch.from_numpy(b4).pin_memory()
        bo = torch.empty((R_loc, c4["T"] + 1), dtype=torch.int32).pin_memory()
        so = torch.empty((R_loc, c4["T"]), dtype=torch.float32).pin_memory()

        def step_rep():
            ctx.solve_into(a4p.numpy(), b4p.numpy(), c4["T"], c4["dist"], c4["risk_part"], np.float32,
                           perm_out=None, bounds_out=bo.numpy(), scores_out=so.numpy())
        step_rep()
        barrier()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            step_rep()
        barrier()
        rep_s = allmax((time.perf_counter() - t0) / reps)
        line["mc_replicas"] = {"value": R_total / rep_s, "unit": "replicas/s", "replicas": R_total,
                               "n": c4["n"], "T": c4["T"], "dtype": "f32", "scaling": "strong",
                               "seconds_per_pass": rep_s, "includes": "H2D of inputs, D2H of boundaries+scores",
                               "cells_per_s_nominal": R_total * synth.nominal_cells(c4["n"], c4["T"]) / rep_s}
        # ---- paper-scale null table (SimulatedExperiments.py:339-345: n=5000, 1000 replicas, T in {2,3,5,10}, both
        # objectives): replicas generated on the device, two batched sweep solves (randomization.py)
        if rank == 0:
            from partitionsolvers_b200 import randomization as rz
            rz.create_null_scores(64, 5000, 100.0, (2, 3, 5, 10), ctx=ctx, seed=1)  # warm-up (allocations)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            tab = rz.create_null_scores(1000, 5000, 100.0, (2, 3, 5, 10), ctx=ctx, seed=1869)
            torch.cuda.synchronize()
            line["null_table_paper_scale"] = {
                "seconds": time.perf_counter() - t0, "replicas": 1000, "n": 5000, "cluster_list": [2, 3, 5, 10],
                "columns": sorted(tab.keys()), "dtype": "f32",
                "what": "device-side replica generation + rp and mcd scores for every T (8 reference solves per replica)"}
        # ---- the reference's own small configurations (BASELINE configs[0], [1]): latency of one host-buffer call
        # (repeated solves of a shape are replayed from a CUDA graph; a context created with PS_NO_GRAPH=1 shows the eager call)
        small = {}
        os.environ["PS_NO_GRAPH"] = "1"
        try:
            from partitionsolvers_b200 import capi as _capi
            ctx_eager = _capi.Context(local_rank)
        finally:
            del os.environ["PS_NO_GRAPH"]
        for nm, dt_small in (("C1", np.float32), ("C2", np.float64)):
            c = synth.CONFIGS[nm]
            a_c, b_c = c["gen"](dt_small)
            per = {}
            for key, cx in (("ms_per_call_host_buffers", ctx), ("ms_per_call_eager_no_graph", ctx_eager)):
                for _ in range(3):
                    cx.solve_raw(a_c, b_c, c["T"], c["dist"], c["risk_part"], dtype=dt_small)
                t0 = time.perf_counter()
                for _ in range(20):
                    cx.solve_raw(a_c, b_c, c["T"], c["dist"], c["risk_part"], dtype=dt_small)
                per[key] = 1e3 * (time.perf_counter() - t0) / 20
            small[nm] = {"n": c["n"], "T": c["T"], "dtype": np.dtype(dt_small).name, **per}
        line["small_configs"] = small
        # ---- OLS partition-size sweep (BASELINE configs[3], second half): one pass to level 64 serves every S = 1..64
        # (DP_impl.hpp:200-230), then the elbow regression (find_optimal_t); planted 4-level mixture as in
        # solver_tests.py:181-201
        if rank == 0:
            from partitionsolvers_b200 import host as H
            rs_o = np.random.RandomState(synth.SEED + 6)
            a_o = (rs_o.choice([-20., -10., 10., 20.], size=5000) + rs_o.normal(0., 1., size=5000)).astype(np.float32)
            b_o = np.ones(5000, np.float32)
            H.DPSolver(5000, 64, a_o, b_o, 0, True, True, sweep_down=True, find_optimal_t=True, dtype=np.float32, ctx=ctx)
            t0 = time.perf_counter()
            s_o = H.DPSolver(5000, 64, a_o, b_o, 0, True, True, sweep_down=True, find_optimal_t=True, dtype=np.float32,
                             ctx=ctx)
            line["ols_sweep"] = {"n": 5000, "T_max": 64, "dtype": "f32", "score": "Gaussian risk-part",
                                 "seconds": time.perf_counter() - t0, "device_ms": ctx.timings()["total_ms"],
                                 "optimal_num_clusters_OLS": int(s_o.get_optimal_num_clusters_OLS_extern()),
                                 "planted_levels": 4,
                                 "what": "one solve to level 64 + 64 backtraces + host OLS (reference: find_optimal_t)"}
        # ---- every score of BASELINE configs[2] at full size (n=100000, T=16): Poisson on planted counts, Gaussian and
        # Rational on the real-valued mixture of solver_tests.py:145-146; device-timed host-buffer calls, best of 3
        if rank == 0:
            names = {0: "gaussian", 1: "poisson", 2: "rational"}
            per_score = {}
            for dt_s in (np.float64, np.float32):
                a_p, b_p = synth.CONFIGS["C3"]["gen"](dt_s)
                a_g, b_g = synth.CONFIGS["C3R"]["gen"](dt_s)
                for d_s, rp_s in ((1, False), (1, True), (0, False), (0, True), (2, False)):
                    aa, bb = (a_p, b_p) if d_s == 1 else (a_g, b_g)
                    best = 1e30
                    for _ in range(3):
                        ctx.solve_raw(aa, bb, T, d_s, rp_s, dtype=dt_s, want_perm=False)
                        best = min(best, ctx.timings()["total_ms"])
                    key = f"{names[d_s]}_{'rp' if rp_s else 'mc'}_{'f64' if dt_s == np.float64 else 'f32'}"
                    per_score[key] = {"ms": best, "value": synth.nominal_cells(n, T) / (best * 1e-3),
                                      "level_kernel": {0: "exhaustive", 1: "screened, exact sums",
                                                       2: "screened, running sums"}[ctx.timings()["screen_used"]]}
                # Poisson on REAL-VALUED data as well (sums round: running-sum screens; float risk-part is the one case left
                # on the exhaustive kernel)
                a_r = (np.abs(a_g) + 0.25).astype(dt_s)
                for rp_s in (False, True):
                    best = 1e30
                    for _ in range(3):
                        ctx.solve_raw(a_r, b_g, T, 1, rp_s, dtype=dt_s, want_perm=False)
                        best = min(best, ctx.timings()["total_ms"])
                    key = f"poisson_{'rp' if rp_s else 'mc'}_realvalued_{'f64' if dt_s == np.float64 else 'f32'}"
                    per_score[key] = {"ms": best, "value": synth.nominal_cells(n, T) / (best * 1e-3),
                                      "level_kernel": {0: "exhaustive", 1: "screened, exact sums",
                                                       2: "screened, running sums"}[ctx.timings()["screen_used"]]}
            line["c3_all_scores"] = per_score
        # ---- exact branch-and-bound is data dependent: the inputs on which it cannot prune, at the headline size, on both
        # level kernels (same tables).  constant a/b: every cell of a row ties; a in {1,2}, b = 1 (gtest_all.cpp:924-958);
        # uniform counts: no planted structure.  The engine hands the levels to the exhaustive kernel once a screened level
        # evaluates > 30 % of its cells exactly (ps_timings.screen_fallback_level)
        if rank == 0:
            npdt_a = np.float64 if args.dtype == "f64" else np.float32
            adv = {}
            for kind in ("constant_ratio", "one_two", "uniform_counts"):
                a_h, b_h = hostile_inputs(kind, n, npdt_a)
                rec = {}
                for label, scr in (("screened_default", 0), ("exhaustive", 1)):
                    best = 1e30
                    for _ in range(2):
                        ctx.solve_raw(a_h, b_h, T, dist_id, rp, dtype=npdt_a, want_perm=False, screen=scr)
                        tmh = ctx.timings()
                        best = min(best, tmh["total_ms"])
                    rec[label + "_ms"] = best
                    if scr == 0:
                        cells_h = float(sum(tmh["level_cells"][:-1]))
                        rec["screen_used"] = int(tmh["screen_used"])
                        rec["exact_cell_fraction"] = tmh["screen_exact_cells"] / max(1.0, cells_h)
                        rec["fallback_level"] = int(tmh["screen_fallback_level"])
                rec["default_over_exhaustive"] = rec["screened_default_ms"] / rec["exhaustive_ms"]
                adv[kind] = rec
            line["adversarial"] = {"n": n, "T": T, "dtype": args.dtype, "score": "Poisson multiple clustering", "cases": adv,
                                   "worst_default_over_exhaustive": max(v["default_over_exhaustive"] for v in adv.values())}
            ex_ms = line["other_level_kernel"]["ms_per_step"] if not args.exhaustive else main_res["ms_per_step"]
            sc_ms = main_res["ms_per_step"] if not args.exhaustive else line["other_level_kernel"]["ms_per_step"]
            # typical case: null-model replicas (flat objectives): screened vs exhaustive on 256 replicas of configs[3]
            a_t, b_t = synth.poisson_null(c4["n"], seed=synth.SEED + 4, dtype=np.float32, replicas=256)
            typ = {}
            for label, scr in (("screened", 0), ("exhaustive", 1)):
                best = 1e30
                for _ in range(2):
                    ctx.solve_raw(a_t, b_t, c4["T"], c4["dist"], c4["risk_part"], dtype=np.float32, want_perm=False, screen=scr)
                    tmt = ctx.timings()
                    best = min(best, tmt["total_ms"])
                typ[label + "_ms"] = best
                if scr == 0:
                    typ["exact_cell_fraction"] = tmt["screen_exact_cells"] / max(1.0, float(sum(tmt["level_cells"][:-1])))
            line["pruning"] = {
                "what": "the default level kernels are an EXACT branch-and-bound (identical maxScore_/nextStart_ tables); their "
                        "speed over the exhaustive kernel depends on how pronounced the row maxima are",
                "best_case": {"workload": "C3 planted 16-level counts (the headline)", "screened_ms": sc_ms,
                              "exhaustive_ms": ex_ms, "speedup": ex_ms / sc_ms,
                              "exact_cell_fraction": line["roofline"].get("screen", {}).get("exact_cell_fraction")},
                "typical_case": dict(typ, workload="256 null-model replicas (configs[3]: n=5000, T=8, f32), flat objectives",
                                     speedup=typ["exhaustive_ms"] / typ["screened_ms"]),
                "worst_case": {"workload": "adversarial inputs at the headline size (see `adversarial`)",
                               "default_over_exhaustive": line["adversarial"]["worst_default_over_exhaustive"]}}
        if world > 1:
            # ---- one large instance (BASELINE configs[4]: n = 1e6, T = 32, f64), rows of every level sharded block-cyclically
            # over the ranks INSIDE the library (ps_solve_sharded_f64): owned-row kernels, one ncclAllGather of the packed
            # {maxScore_, nextStart_} chunk per level on the engine's stream, scatter, replicated last level and backtrace
            n_s, T_s = args.sharded_n, args.sharded_T
            a_s, b_s = synth.poisson_planted(n_s, T_s, seed=synth.SEED + 5, dtype=np.float64)
            ids = [ctx.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            comm = ctx.comm_create(ids[0], rank, world)
            ctx.solve_sharded(comm, a_s, b_s, T_s, 1, False, dtype=np.float64)  # warm-up: NCCL channels, workspaces
            secs = []
            for _ in range(3):
                barrier()
                t0 = time.perf_counter()
                sh = ctx.solve_sharded(comm, a_s, b_s, T_s, 1, False, dtype=np.float64)
                torch.cuda.synchronize()
                secs.append(allmax(time.perf_counter() - t0))
            tms = ctx.timings()
            sh_s = float(np.mean(secs))
            line["sharded"] = {"n": n_s, "T": T_s, "dtype": "f64", "seconds": sh_s, "seconds_best": float(np.min(secs)),
                               "scaling": "strong", "ranks": world,
                               "device_ms_rank0": {k: tms[k] for k in ("total_ms", "sort_scan_ms", "prepass_ms", "levels_ms",
                                                                        "backtrace_ms", "h2d_ms", "d2h_ms")},
                               "levels_seconds_max_over_ranks": allmax(tms["levels_ms"] * 1e-3),
                               "comm_bytes_per_rank": int(tms["comm_bytes"]), "level_kernel": int(tms["screen_used"]),
                               "value": synth.nominal_cells(n_s, T_s) / sh_s, "unit": UNIT,
                               "collective": "ncclAllGather of the packed {value, split} level chunk, issued by the library on "
                                             "its own stream (C ABI ps_solve_sharded_f64); no host sync between levels",
                               "includes": "H2D of a, b on every rank, D2H of sort index + boundaries + subset scores"}
            ctx.comm_destroy(comm)
            if rank == 0:
                ctx.solve_raw(a_s, b_s, T_s, 1, False, dtype=np.float64)
                t0 = time.perf_counter()
                one = ctx.solve_raw(a_s, b_s, T_s, 1, False, dtype=np.float64)
                line["sharded"]["single_gpu_seconds"] = time.perf_counter() - t0
                line["sharded"]["single_gpu_levels_seconds"] = ctx.timings()["levels_ms"] * 1e-3
                line["sharded"]["speedup_vs_single_gpu"] = line["sharded"]["single_gpu_seconds"] / sh_s
                line["sharded"]["same_partition_as_single_gpu"] = bool(
                    np.array_equal(one["bounds"].ravel(), sh["bounds"].ravel()) and np.array_equal(one["perm"], sh["perm"])
                    and np.array_equal(one["subset_scores"].ravel(), sh["subset_scores"].ravel()))
        if rank == 0 and world == 1:
            dtype = np.float64 if args.dtype == "f64" else np.float32
            # (1) SAME CONFIGURATION where the box allows it: the unmodified reference DPSolver<float> at n = 100000, T = 16
            full = cpu_reference_full_size()
            line["cpu_reference_full_size"] = full
            # (2) bounded sample in the headline precision (the f64 tables, 240 GB, never fit)
            which, kind, n_s = cpu_reference_sample(dtype, 10.0)
            t, _ = run_cpu_solver(which, dtype, n_s)
            sample_rec = {
                "value": synth.nominal_cells(n_s, T) / t, "unit": UNIT, "cores": cpu_threads_used(kind),
                "kind": kind, "host_cores": os.cpu_count(), "same_config": False,
                "sample": f"one {'reference DPSolver' if kind == 'reference' else 'oracle port'} solve at n={n_s}, "
                          f"T={T}, {args.dtype}, same generator ({t:.2f} s); the reference cannot allocate n=100000 in f64"}
            if full.get("ran"):
                f32_ms = (line.get("variants", {}).get("f32", {}) or {}).get("e2e", {}).get("ms_per_step") if args.dtype == "f64" \
                    else line["e2e"]["ms_per_step"]
                # the reference's own objective at full size next to ours (device sort: stable order)
                from partitionsolvers_b200 import host as H
                a32, b32 = cfg["gen"](np.float32)
                ours = H.DPSolver(n, T, a32, b32, dist_id, rp, dtype=np.float32, ctx=ctx)
                ours_score = float(ours.get_optimal_score_extern())
                line["cpu_baseline"] = {
                    "value": full["value"], "unit": UNIT, "cores": cpu_threads_used("reference"), "kind": "reference",
                    "host_cores": os.cpu_count(), "same_config": True, "dtype": "f32", "seconds": full["seconds"],
                    "sample": f"the unmodified reference DPSolver<float> constructor on the headline workload itself: n={n}, "
                              f"T={T}, Poisson multiple clustering, {full['table_bytes'] / 1e9:.0f} GB of tables, "
                              f"{full['seconds']:.1f} s (f64 needs 240 GB and does not fit)",
                    "gpu_same_config": {"dtype": "f32", "e2e_ms_per_step": f32_ms,
                                        "speedup_e2e": (full["seconds"] * 1e3 / f32_ms) if f32_ms else None},
                    "objective": {"reference": full["score"], "this_library_default_sort": ours_score,
                                  "rel_diff": abs(ours_score - full["score"]) / max(abs(full["score"]), 1e-300)},
                    "bounded_sample_in_headline_precision": sample_rec}
            else:
                line["cpu_baseline"] = dict(sample_rec, full_size_leg=full)
            try:
                # the reference's native-threaded variant on the replica workload: independent DPSolver<float>
                # constructions dispatched on its own ThreadPool (threadpool.hpp:105-113, the call pattern of
                # sweep_parallel__DP, python_dpsolver.cpp:263-316); hardware_concurrency()-1 workers
                from oracle import oracle_py as O
                c4r = synth.CONFIGS["C4"]
                r_cpu = 2 * max(1, (os.cpu_count() or 2) - 1)
                a_r, b_r = synth.poisson_null(c4r["n"], seed=synth.SEED + 4, dtype=np.float32, replicas=r_cpu)
                t_pool, _, workers = O.time_pool_ref(a_r, b_r, c4r["T"], c4r["dist"], c4r["risk_part"])
                line["cpu_replicas_baseline"] = {
                    "value": r_cpu / t_pool, "unit": "replicas/s", "cores": workers, "kind": "reference",
                    "host_cores": os.cpu_count(),
                    "sample": f"{r_cpu} null-model replicas (n={c4r['n']}, T={c4r['T']}, f32) on the reference's own "
                              f"ThreadPool, {workers} workers ({t_pool:.2f} s)"}
            except Exception as e:
                line["cpu_replicas_baseline"] = {"error": str(e)}
            try:
                from oracle import oracle_py as O
                n_p = 60000  # ~6 s on 16 cores
                ap_, bp_ = cfg["gen"](dtype, n_p)
                tp, _ = O.time_solve("port", ap_, bp_, T, dist_id, rp, dtype=dtype, nthreads=0)
                line["cpu_baseline_all_cores"] = {
                    "value": synth.nominal_cells(n_p, T) / tp, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                    "sample": f"streaming oracle port, all host threads, n={n_p}, T={T} ({tp:.2f} s)"}
            except Exception as e:  # the port is optional extra context
                line["cpu_baseline_all_cores"] = {"error": str(e)}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
