"""ctypes binding of the C ABI in include/ps_b200.h (libps_b200.so).

This module is plumbing for tests and bench.py: it loads the in-tree CUDA library, mirrors the ABI
structs, and exposes :class:`Context`.  There is no fallback: if the library is missing or no CUDA
device is usable, it raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libps_b200.so")

PS_OK = 0
PS_ERR_DIST = 2
PS_ERR_NODEVICE = 5
PS_SUM_RUNNING, PS_SUM_PREFIX = 0, 1
PS_SORT_DEVICE, PS_SORT_GIVEN = 0, 1
PS_MAX_TIMED_LEVELS = 130

# every symbol include/ps_b200.h declares (tests/test_abi.py checks the .so exports all of them)
ABI_SYMBOLS = [
    "ps_abi_version", "ps_status_string", "ps_create", "ps_destroy", "ps_last_error", "ps_device", "ps_device_count",
    "ps_stream", "ps_solve_f32", "ps_solve_f64", "ps_get_timings", "ps_range_score_f32",
    "ps_range_score_f64", "ps_shard_create_f64", "ps_shard_create_f32", "ps_shard_chunk_rows",
    "ps_shard_level1", "ps_shard_set_prev", "ps_shard_level", "ps_shard_next_start", "ps_shard_owner",
    "ps_shard_destroy", "ps_selftest_math", "ps_eval_log_f64", "ps_eval_log_f32",
    "ps_ltss_f32", "ps_selftest_mufu", "ps_shard_screened", "ps_shard_subset_scores",
    "ps_comm_unique_id", "ps_comm_create", "ps_comm_destroy", "ps_comm_rank", "ps_comm_world", "ps_nccl_version",
    "ps_solve_sharded_f32", "ps_solve_sharded_f64", "ps_stream_wait",
]
PS_COMM_ID_BYTES = 128


class PsProblem(C.Structure):
    _fields_ = [("n", C.c_int), ("T", C.c_int), ("dist", C.c_int), ("risk_part", C.c_int),
                ("sum_mode", C.c_int), ("sort_mode", C.c_int), ("perm", C.c_void_p), ("sweep", C.c_int),
                ("batch", C.c_int), ("on_device", C.c_int), ("no_fast_math", C.c_int), ("chunk_len", C.c_int),
                ("screen", C.c_int)]


class PsResult(C.Structure):
    _fields_ = [("perm", C.c_void_p), ("bounds", C.c_void_p), ("subset_scores", C.c_void_p),
                ("max_score", C.c_void_p), ("next_start", C.c_void_p)]


class PsTimings(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("h2d_ms", C.c_float), ("sort_scan_ms", C.c_float),
                ("prepass_ms", C.c_float), ("levels_ms", C.c_float), ("backtrace_ms", C.c_float),
                ("d2h_ms", C.c_float), ("n_levels", C.c_int), ("level_ms", C.c_float * PS_MAX_TIMED_LEVELS),
                ("level_cells", C.c_longlong * PS_MAX_TIMED_LEVELS), ("kernel_launches", C.c_longlong),
                ("h2d_bytes", C.c_longlong), ("d2h_bytes", C.c_longlong), ("screen_used", C.c_int),
                ("screen_exact_cells", C.c_longlong), ("screen_violations", C.c_longlong),
                ("screen_tiers", C.c_longlong * 6), ("comm_bytes", C.c_longlong), ("screen_fallback_level", C.c_int)]


class PsError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"ps_b200 status {status}: {msg}")
        self.status = status


class DistributionError(PsError):
    """PS_ERR_DIST: the reference throws distributionException here (LTSS.hpp:21-25)."""


_lib = None


def load():
    """Load libps_b200.so (built in-tree by __graft_entry__.build()).  Fails loudly when absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.ps_status_string.restype = C.c_char_p
        L.ps_last_error.restype = C.c_char_p
        L.ps_last_error.argtypes = [C.c_void_p]
        L.ps_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.ps_destroy.argtypes = [C.c_void_p]
        L.ps_destroy.restype = None
        L.ps_device.argtypes = [C.c_void_p]
        L.ps_stream.argtypes = [C.c_void_p]
        L.ps_stream.restype = C.c_size_t
        for nm in ("ps_solve_f32", "ps_solve_f64"):
            getattr(L, nm).argtypes = [C.c_void_p, C.POINTER(PsProblem), C.c_void_p, C.c_void_p,
                                       C.POINTER(PsResult)]
        L.ps_get_timings.argtypes = [C.c_void_p, C.POINTER(PsTimings)]
        for nm in ("ps_range_score_f32", "ps_range_score_f64"):
            getattr(L, nm).argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.c_void_p]
        for nm in ("ps_shard_create_f32", "ps_shard_create_f64"):
            getattr(L, nm).argtypes = [C.c_void_p, C.POINTER(PsProblem), C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_int, C.POINTER(C.c_void_p)]
        L.ps_shard_chunk_rows.argtypes = [C.c_void_p]
        L.ps_shard_level1.argtypes = [C.c_void_p, C.c_void_p]
        L.ps_shard_set_prev.argtypes = [C.c_void_p, C.c_void_p]
        L.ps_shard_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ps_shard_next_start.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.ps_shard_owner.argtypes = [C.c_void_p, C.c_int]
        L.ps_shard_subset_scores.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ps_shard_screened.argtypes = [C.c_void_p]
        L.ps_shard_destroy.argtypes = [C.c_void_p]
        L.ps_shard_destroy.restype = None
        L.ps_selftest_math.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_ulonglong, C.POINTER(C.c_longlong)]
        for nm in ("ps_eval_log_f64", "ps_eval_log_f32"):
            getattr(L, nm).argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ps_selftest_mufu.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.ps_stream_wait.argtypes = [C.c_void_p, C.c_size_t]
        L.ps_comm_unique_id.argtypes = [C.c_void_p]
        L.ps_comm_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.ps_comm_destroy.argtypes = [C.c_void_p]
        L.ps_comm_destroy.restype = None
        for nm in ("ps_solve_sharded_f32", "ps_solve_sharded_f64"):
            getattr(L, nm).argtypes = [C.c_void_p, C.c_void_p, C.POINTER(PsProblem), C.c_void_p, C.c_void_p,
                                       C.POINTER(PsResult)]
        L.ps_ltss_f32.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                  C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_float)]
        _lib = L
    return _lib


def _vp(arr):
    return arr.ctypes.data_as(C.c_void_p) if arr is not None else None


class Context:
    """One ps_ctx: a device, a stream and grow-only workspaces.  Use from one thread at a time."""

    def __init__(self, device=-1):
        self.lib = load()
        self._h = C.c_void_p()
        st = self.lib.ps_create(C.byref(self._h), int(device))
        if st != PS_OK:
            raise PsError(st, self.lib.ps_status_string(st).decode())

    def close(self):
        if self._h:
            self.lib.ps_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != PS_OK:
            msg = self.lib.ps_last_error(self._h).decode() or self.lib.ps_status_string(st).decode()
            raise (DistributionError if st == PS_ERR_DIST else PsError)(st, msg)

    @property
    def device(self):
        return self.lib.ps_device(self._h)

    @property
    def stream(self):
        return self.lib.ps_stream(self._h)

    def wait_stream(self, other_stream):
        """Order the engine's stream after the work enqueued so far on `other_stream` (a cudaStream_t as an integer, e.g.
        torch.cuda.current_stream().cuda_stream): device buffers produced there may then be passed to solve_device."""
        self._check(self.lib.ps_stream_wait(self._h, int(other_stream)))

    def wait_torch(self, device=None):
        """wait_stream on torch's current stream of `device` (torch is imported only here)."""
        import torch
        self.wait_stream(torch.cuda.current_stream(device).cuda_stream)

    def timings(self):
        t = PsTimings()
        self._check(self.lib.ps_get_timings(self._h, C.byref(t)))
        nl = t.n_levels
        return dict(total_ms=t.total_ms, h2d_ms=t.h2d_ms, sort_scan_ms=t.sort_scan_ms,
                    prepass_ms=t.prepass_ms, levels_ms=t.levels_ms, backtrace_ms=t.backtrace_ms,
                    d2h_ms=t.d2h_ms, level_ms=[t.level_ms[i] for i in range(nl)],
                    level_cells=[t.level_cells[i] for i in range(nl)],
                    kernel_launches=t.kernel_launches, h2d_bytes=t.h2d_bytes, d2h_bytes=t.d2h_bytes,
                    screen_used=t.screen_used, screen_exact_cells=t.screen_exact_cells,
                    screen_violations=t.screen_violations, screen_tiers=[t.screen_tiers[i] for i in range(6)],
                    screen_fallback_level=t.screen_fallback_level, comm_bytes=t.comm_bytes)

    def solve_raw(self, a, b, T, dist, risk_part, dtype=np.float32, sum_mode=PS_SUM_RUNNING, perm=None,
                  sweep=False, levels=False, chunk_len=0, want_perm=True, no_fast=False, screen=0):
        """Host-buffer call.  a, b: [n] or [R, n].  Returns the raw ABI outputs as numpy arrays."""
        a = np.ascontiguousarray(a, dtype=dtype)
        b = np.ascontiguousarray(b, dtype=dtype)
        assert a.shape == b.shape
        batched = a.ndim == 2
        R = a.shape[0] if batched else 1
        n = a.shape[-1]
        Tc = min(T, n)
        nS = Tc + 1 if sweep else 1
        p = PsProblem(n=n, T=T, dist=dist, risk_part=int(risk_part), sum_mode=sum_mode,
                      sort_mode=PS_SORT_GIVEN if perm is not None else PS_SORT_DEVICE, perm=None,
                      sweep=int(sweep), batch=R, on_device=0, no_fast_math=int(no_fast), chunk_len=chunk_len,
                      screen=int(screen))
        if perm is not None:
            perm = np.ascontiguousarray(perm, dtype=np.int32).reshape(R, n)
            p.perm = perm.ctypes.data
        out_perm = np.empty((R, n), np.int32) if want_perm else None
        bounds = np.empty((R, nS, Tc + 1), np.int32)
        sscores = np.empty((R, nS, Tc), dtype)
        ms = np.empty((R, Tc + 1, n), dtype) if levels else None
        ns = np.empty((R, Tc + 1, n), np.int32) if levels else None
        r = PsResult(perm=_vp(out_perm), bounds=_vp(bounds), subset_scores=_vp(sscores), max_score=_vp(ms),
                     next_start=_vp(ns))
        fn = self.lib.ps_solve_f32 if dtype == np.float32 else self.lib.ps_solve_f64
        self._check(fn(self._h, C.byref(p), _vp(a), _vp(b), C.byref(r)))
        out = dict(n=n, T=Tc, R=R, perm=out_perm, bounds=bounds, subset_scores=sscores, max_score=ms,
                   next_start=ns)
        if not batched:
            for k in ("perm", "bounds", "subset_scores", "max_score", "next_start"):
                if out[k] is not None:
                    out[k] = out[k][0]
        return out

    def solve_into(self, a, b, T, dist, risk_part, dtype, perm_out=None, bounds_out=None, scores_out=None,
                   sum_mode=PS_SUM_RUNNING, sweep=False, chunk_len=0, screen=0):
        """Host-buffer call that writes into caller-provided numpy arrays (e.g. views of pinned memory);
        nothing is allocated here, so the call is exactly the ABI call plus its copies."""
        batched = a.ndim == 2
        R = a.shape[0] if batched else 1
        n = a.shape[-1]
        p = PsProblem(n=n, T=T, dist=dist, risk_part=int(risk_part), sum_mode=sum_mode, sort_mode=PS_SORT_DEVICE,
                      perm=None, sweep=int(sweep), batch=R, on_device=0, no_fast_math=0, chunk_len=chunk_len,
                      screen=int(screen))
        r = PsResult(perm=_vp(perm_out), bounds=_vp(bounds_out), subset_scores=_vp(scores_out), max_score=None,
                     next_start=None)
        fn = self.lib.ps_solve_f32 if dtype == np.float32 else self.lib.ps_solve_f64
        self._check(fn(self._h, C.byref(p), _vp(a), _vp(b), C.byref(r)))

    def solve_device(self, d_a, d_b, n, T, dist, risk_part, dtype, R=1, sum_mode=PS_SUM_RUNNING, sweep=False,
                     d_perm_out=0, d_bounds=0, d_scores=0, d_perm_in=0, chunk_len=0, screen=0):
        """Device-pointer call (inputs resident in HBM; integers are raw device addresses)."""
        p = PsProblem(n=n, T=T, dist=dist, risk_part=int(risk_part), sum_mode=sum_mode,
                      sort_mode=PS_SORT_GIVEN if d_perm_in else PS_SORT_DEVICE, perm=d_perm_in or None,
                      sweep=int(sweep), batch=R, on_device=1, no_fast_math=0, chunk_len=chunk_len,
                      screen=int(screen))
        r = PsResult(perm=d_perm_out or None, bounds=d_bounds or None, subset_scores=d_scores or None,
                     max_score=None, next_start=None)
        fn = self.lib.ps_solve_f32 if dtype == np.float32 else self.lib.ps_solve_f64
        self._check(fn(self._h, C.byref(p), C.c_void_p(d_a), C.c_void_p(d_b), C.byref(r)))

    # ---- row-sharded solve of one instance across the GPUs of a communicator (include/ps_b200.h: ps_solve_sharded_*)
    def comm_unique_id(self):
        """Rank 0: the bytes every rank must pass to :meth:`comm_create` (send them by any means)."""
        buf = C.create_string_buffer(PS_COMM_ID_BYTES)
        self._check(self.lib.ps_comm_unique_id(buf))
        return buf.raw

    def comm_create(self, id_bytes, rank, world):
        h = C.c_void_p()
        self._check(self.lib.ps_comm_create(self._h, C.c_char_p(id_bytes), int(rank), int(world), C.byref(h)))
        return h

    def comm_destroy(self, comm):
        self.lib.ps_comm_destroy(comm)

    def solve_sharded(self, comm, a, b, T, dist, risk_part, dtype=np.float64, sweep=False, levels=False, screen=0,
                      sum_mode=PS_SUM_RUNNING, perm=None):
        """Collective over the ranks of `comm`: same a, b on every rank; every rank gets the full result."""
        a = np.ascontiguousarray(a, dtype=dtype)
        b = np.ascontiguousarray(b, dtype=dtype)
        n = a.shape[0]
        Tc = min(T, n)
        nS = Tc + 1 if sweep else 1
        p = PsProblem(n=n, T=T, dist=dist, risk_part=int(risk_part), sum_mode=sum_mode,
                      sort_mode=PS_SORT_GIVEN if perm is not None else PS_SORT_DEVICE, perm=None, sweep=int(sweep),
                      batch=1, on_device=0, no_fast_math=0, chunk_len=0, screen=int(screen))
        if perm is not None:
            perm = np.ascontiguousarray(perm, dtype=np.int32)
            p.perm = perm.ctypes.data
        out_perm = np.empty(n, np.int32)
        bounds = np.empty((nS, Tc + 1), np.int32)
        sscores = np.empty((nS, Tc), dtype)
        ms = np.empty((Tc + 1, n), dtype) if levels else None
        ns = np.empty((Tc + 1, n), np.int32) if levels else None
        r = PsResult(perm=_vp(out_perm), bounds=_vp(bounds), subset_scores=_vp(sscores), max_score=_vp(ms),
                     next_start=_vp(ns))
        fn = self.lib.ps_solve_sharded_f32 if dtype == np.float32 else self.lib.ps_solve_sharded_f64
        self._check(fn(self._h, comm, C.byref(p), _vp(a), _vp(b), C.byref(r)))
        return dict(n=n, T=Tc, perm=out_perm, bounds=bounds, subset_scores=sscores, max_score=ms, next_start=ns)

    def selftest_math(self, which, samples, seed=1):
        out = C.c_longlong(-1)
        self._check(self.lib.ps_selftest_math(self._h, which, samples, seed, C.byref(out)))
        return out.value

    def selftest_mufu(self):
        """(max |x rcp(x) - 1|, max lg2 error) in units of 2^-24 over every float of the safe range."""
        out = (C.c_double * 2)()
        self._check(self.lib.ps_selftest_mufu(self._h, out))
        return float(out[0]), float(out[1])

    def ltss(self, a, b, dist):
        """LTSSSolver<float>: returns dict(perm, subset, score)."""
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        n = a.shape[0]
        perm = np.empty(n, np.int32)
        lo, hi, sc = C.c_int(0), C.c_int(0), C.c_float(0)
        self._check(self.lib.ps_ltss_f32(self._h, n, _vp(a), _vp(b), int(dist), _vp(perm), C.byref(lo), C.byref(hi),
                                         C.byref(sc)))
        return dict(perm=perm, subset=perm[lo.value:hi.value].copy(), score=np.float32(sc.value))

    def eval_log(self, x, dtype=np.float64):
        """Device log/logf on host arguments -> (general entry, safe-range entry)."""
        x = np.ascontiguousarray(x, dtype=dtype)
        yg, yf = np.empty_like(x), np.empty_like(x)
        fn = self.lib.ps_eval_log_f64 if dtype == np.float64 else self.lib.ps_eval_log_f32
        self._check(fn(self._h, x.shape[0], _vp(x), _vp(yg), _vp(yf)))
        return yg, yf

    def range_score(self, a, b, dist, risk_part, dtype=np.float32):
        a = np.ascontiguousarray(a, dtype=dtype)
        b = np.ascontiguousarray(b, dtype=dtype)
        out = np.zeros(1, dtype)
        fn = self.lib.ps_range_score_f32 if dtype == np.float32 else self.lib.ps_range_score_f64
        self._check(fn(self._h, a.shape[0], _vp(a), _vp(b), dist, int(risk_part), _vp(out)))
        return out[0]
