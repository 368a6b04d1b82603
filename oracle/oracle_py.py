"""ctypes loaders for the two CPU checkers.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

* ``port``  -> oracle/_build/libps_oracle.so  (dp_oracle.cpp, our streaming restatement)
* ``ref``   -> oracle/_ref/libps_ref.so       (the unmodified reference headers behind ref_shim.cpp)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs
may import this module.  Both libraries expose the same solve signature (the port adds ``nthreads``,
``perm_in`` and ``ols_t``), so :func:`solve` returns the same dict for either.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "_build", "libps_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libps_ref.so")

_libs = {}


def build(which=("port", "ref")):
    """Run oracle/Makefile (the ref target is a no-op where /root/reference is absent)."""
    for w in which:
        subprocess.run(["make", "-s", "-C", _HERE, w], check=True)


def have(which):
    return os.path.exists(PORT_SO if which == "port" else REF_SO)


def lib(which):
    if which not in _libs:
        path = PORT_SO if which == "port" else REF_SO
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle {which}`")
        _libs[which] = C.CDLL(path)
    return _libs[which]


def _ptr(arr, ct):
    return arr.ctypes.data_as(C.POINTER(ct)) if arr is not None else None


def solve(which, a, b, T, dist, risk_part, gamma=0.0, reg_power=1, sweep=False, dtype=np.float32,
          levels=True, nthreads=0, perm_in=None, ols=False):
    """Run one solve on the chosen checker and return every observable as numpy arrays."""
    a = np.ascontiguousarray(a, dtype=dtype)
    b = np.ascontiguousarray(b, dtype=dtype)
    n = a.shape[0]
    Tc = min(T, n)
    ct = C.c_float if dtype == np.float32 else C.c_double
    sfx = "f32" if dtype == np.float32 else "f64"
    perm = np.empty(n, np.int32)
    ms = np.empty((Tc + 1, n), dtype) if levels else None
    ns = np.empty((Tc + 1, n), np.int32) if levels else None
    nsub = C.c_int(0)
    sizes = np.zeros(Tc, np.int32)
    elems = np.zeros(n, np.int32)
    sbs = np.zeros(Tc, dtype)
    getter = ct(0)
    member = ct(0)
    all_sizes = np.zeros((Tc + 1, Tc), np.int32) if sweep else None
    all_elems = np.zeros((Tc + 1, n), np.int32) if sweep else None
    all_scores = np.zeros(Tc + 1, dtype) if sweep else None
    L = lib(which)
    if which == "ref":
        fn = getattr(L, f"ref_dp_solve_{sfx}")
        fn.restype = C.c_int
        rc = fn(C.c_int(n), C.c_int(T), _ptr(a, ct), _ptr(b, ct), C.c_int(dist), C.c_int(int(risk_part)),
                ct(gamma), C.c_int(reg_power), C.c_int(int(sweep)), _ptr(perm, C.c_int), _ptr(ms, ct),
                _ptr(ns, C.c_int), C.byref(nsub), _ptr(sizes, C.c_int), _ptr(elems, C.c_int),
                _ptr(sbs, ct), C.byref(getter), C.byref(member), _ptr(all_sizes, C.c_int),
                _ptr(all_elems, C.c_int), _ptr(all_scores, ct))
        ols_t = None
    else:
        fn = getattr(L, f"oracle_dp_solve_{sfx}")
        fn.restype = C.c_int
        pin = np.ascontiguousarray(perm_in, np.int32) if perm_in is not None else None
        ols_c = C.c_int(0)
        rc = fn(C.c_int(n), C.c_int(T), _ptr(a, ct), _ptr(b, ct), C.c_int(dist), C.c_int(int(risk_part)),
                ct(gamma), C.c_int(reg_power), C.c_int(int(sweep)), C.c_int(nthreads),
                _ptr(pin, C.c_int), _ptr(perm, C.c_int), _ptr(ms, ct), _ptr(ns, C.c_int),
                C.byref(nsub), _ptr(sizes, C.c_int), _ptr(elems, C.c_int), _ptr(sbs, ct),
                C.byref(getter), C.byref(member), _ptr(all_sizes, C.c_int), _ptr(all_elems, C.c_int),
                _ptr(all_scores, ct), C.byref(ols_c) if ols else None)
        ols_t = ols_c.value if ols else None
    if rc != 0:
        raise RuntimeError(f"{which} oracle failed rc={rc}")
    k = nsub.value
    subsets, off = [], 0
    for s in range(k):
        subsets.append(elems[off:off + sizes[s]].copy())
        off += sizes[s]
    out = dict(n=n, T=Tc, perm=perm, max_score=ms, next_start=ns, subsets=subsets,
               score_by_subset=sbs, score=dtype(getter.value), optimal_score_member=dtype(member.value),
               ols_t=ols_t)
    if sweep:
        allp = []
        for S in range(Tc + 1):
            off, ss = 0, []
            for q in range(S):
                ss.append(all_elems[S, off:off + all_sizes[S, q]].copy())
                off += all_sizes[S, q]
            allp.append((ss, dtype(all_scores[S])))
        out["all"] = allp
    return out


def time_solve(which, a, b, T, dist, risk_part, dtype=np.float32, nthreads=0):
    """Wall seconds of one full CPU solve (reference: constructor only, as solver_timer.cpp:71-73)."""
    a = np.ascontiguousarray(a, dtype=dtype)
    b = np.ascontiguousarray(b, dtype=dtype)
    n = a.shape[0]
    ct = C.c_float if dtype == np.float32 else C.c_double
    sfx = "f32" if dtype == np.float32 else "f64"
    score = ct(0)
    L = lib(which)
    if which == "ref":
        fn = getattr(L, f"ref_dp_time_{sfx}")
        fn.restype = C.c_double
        t = fn(C.c_int(n), C.c_int(T), _ptr(a, ct), _ptr(b, ct), C.c_int(dist), C.c_int(int(risk_part)),
               C.byref(score))
    else:
        fn = getattr(L, f"oracle_dp_time_{sfx}")
        fn.restype = C.c_double
        t = fn(C.c_int(n), C.c_int(T), _ptr(a, ct), _ptr(b, ct), C.c_int(dist), C.c_int(int(risk_part)),
               C.c_int(nthreads), C.byref(score))
    return float(t), float(score.value)


def time_pool_ref(a, b, T, dist, risk_part):
    """R independent f32 solves on the reference's own ThreadPool; returns (seconds, scores, workers)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    R, n = a.shape
    scores = np.zeros(R, np.float32)
    nw = C.c_int(0)
    fn = lib("ref").ref_dp_time_pool_f32
    fn.restype = C.c_double
    t = fn(C.c_int(R), C.c_int(n), C.c_int(T), _ptr(a, C.c_float), _ptr(b, C.c_float), C.c_int(dist),
           C.c_int(int(risk_part)), _ptr(scores, C.c_float), C.byref(nw))
    return float(t), scores, nw.value


def range_score(a, b, dist, risk_part, i, j, dtype=np.float32):
    a = np.ascontiguousarray(a, dtype=dtype)
    b = np.ascontiguousarray(b, dtype=dtype)
    ct = C.c_float if dtype == np.float32 else C.c_double
    sfx = "f32" if dtype == np.float32 else "f64"
    fn = getattr(lib("port"), f"oracle_range_score_{sfx}")
    fn.restype = ct
    return dtype(fn(C.c_int(a.shape[0]), _ptr(a, ct), _ptr(b, ct), C.c_int(dist), C.c_int(int(risk_part)),
                    C.c_int(i), C.c_int(j)))


def ref_score_pairs(a, b, dist, risk_part, ii, jj, dtype=np.float32):
    a = np.ascontiguousarray(a, dtype=dtype)
    b = np.ascontiguousarray(b, dtype=dtype)
    ii = np.ascontiguousarray(ii, np.int32)
    jj = np.ascontiguousarray(jj, np.int32)
    ct = C.c_float if dtype == np.float32 else C.c_double
    sfx = "f32" if dtype == np.float32 else "f64"
    out = np.zeros(ii.shape[0], dtype)
    fn = getattr(lib("ref"), f"ref_score_pairs_{sfx}")
    fn.restype = C.c_int
    rc = fn(C.c_int(a.shape[0]), _ptr(a, ct), _ptr(b, ct), C.c_int(dist), C.c_int(int(risk_part)),
            C.c_int(ii.shape[0]), _ptr(ii, C.c_int), _ptr(jj, C.c_int), _ptr(out, ct))
    if rc:
        raise RuntimeError("ref_score_pairs failed")
    return out


def dp_rows(a_sorted, b_sorted, prev, kmax, dist, risk_part, rows, dtype=np.float32):
    """(best, argmax) of single DP rows given the previous level vector (port only)."""
    a = np.ascontiguousarray(a_sorted, dtype=dtype)
    b = np.ascontiguousarray(b_sorted, dtype=dtype)
    prev = np.ascontiguousarray(prev, dtype=dtype)
    rows = np.ascontiguousarray(rows, np.int32)
    ct = C.c_float if dtype == np.float32 else C.c_double
    best = np.zeros(rows.shape[0], dtype)
    arg = np.zeros(rows.shape[0], np.int32)
    fn = lib("port").oracle_dp_rows_f32 if dtype == np.float32 else lib("port").oracle_dp_rows_f64
    fn.restype = None
    fn(C.c_int(a.shape[0]), C.c_int(kmax), C.c_int(dist), C.c_int(int(risk_part)), _ptr(a, ct), _ptr(b, ct),
       _ptr(prev, ct), C.c_int(rows.shape[0]), _ptr(rows, C.c_int), _ptr(best, ct), _ptr(arg, C.c_int))
    return best, arg


def ref_solve_truncated(n, a, b, T, dist, risk_part):
    """The reference constructor with n < len(a) (f32): returns (subsets, getter score)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    Tc = min(T, n)
    nsub, sizes, elems, sc = C.c_int(0), np.zeros(Tc, np.int32), np.zeros(n, np.int32), C.c_float(0)
    fn = lib("ref").ref_dp_solve_trunc_f32
    fn.restype = C.c_int
    rc = fn(C.c_int(n), C.c_int(a.shape[0]), C.c_int(T), _ptr(a, C.c_float), _ptr(b, C.c_float), C.c_int(dist),
            C.c_int(int(risk_part)), C.byref(nsub), _ptr(sizes, C.c_int), _ptr(elems, C.c_int), C.byref(sc))
    if rc:
        raise RuntimeError("ref_dp_solve_trunc_f32 failed")
    subsets, off = [], 0
    for s in range(nsub.value):
        subsets.append(elems[off:off + sizes[s]].tolist())
        off += sizes[s]
    return subsets, np.float32(sc.value)


def ltss(which, a, b, dist):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    n = a.shape[0]
    perm = np.empty(n, np.int32)
    sub = np.empty(n, np.int32)
    sz = C.c_int(0)
    sc = C.c_float(0)
    fn = lib(which).ref_ltss_f32 if which == "ref" else lib(which).oracle_ltss_f32
    fn.restype = C.c_int
    rc = fn(C.c_int(n), _ptr(a, C.c_float), _ptr(b, C.c_float), C.c_int(dist), _ptr(perm, C.c_int),
            C.byref(sz), _ptr(sub, C.c_int), C.byref(sc))
    if rc:
        raise RuntimeError("ltss failed")
    return dict(perm=perm, subset=sub[:sz.value].copy(), score=np.float32(sc.value))


def ols(scores, dtype=np.float32):
    s = np.ascontiguousarray(scores, dtype=dtype)
    ct = C.c_float if dtype == np.float32 else C.c_double
    fn = lib("port").oracle_ols_f32 if dtype == np.float32 else lib("port").oracle_ols_f64
    fn.restype = C.c_int
    return int(fn(C.c_int(s.shape[0]), _ptr(s, ct)))


def host_log(x, dtype=np.float64):
    """libm log/logf of the host (the function the reference's Poisson scores call)."""
    x = np.ascontiguousarray(x, dtype=dtype)
    y = np.empty_like(x)
    ct = C.c_float if dtype == np.float32 else C.c_double
    fn = lib("port").oracle_log_f32 if dtype == np.float32 else lib("port").oracle_log_f64
    fn.restype = None
    fn(C.c_longlong(x.shape[0]), _ptr(x, ct), _ptr(y, ct))
    return y


def hardware_threads():
    fn = lib("port").oracle_hardware_threads
    fn.restype = C.c_int
    return int(fn())
