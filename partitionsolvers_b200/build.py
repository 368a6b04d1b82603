"""In-tree builds (nvcc / g++ invoked directly; nothing is JIT-cached outside the repository).

    python -m partitionsolvers_b200.build [--force]

Outputs (git-ignored, shipped to the GPU box by gpurun):
    partitionsolvers_b200/_lib/libps_b200.so     CUDA kernels + C ABI, sm_100a only
    partitionsolvers_b200/_lib/proto*.so         pybind11 module `proto` (reference proto.i surface)
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOSTSRC = os.path.join(HERE, "cxx")
LIBDIR = os.path.join(HERE, "_lib")
INCLUDE = os.path.join(ROOT, "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(d, exts):
    out = []
    for dp, _, files in os.walk(d):
        out += [os.path.join(dp, f) for f in files if f.endswith(exts)]
    return out


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def build_cuda(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    target = os.path.join(LIBDIR, "libps_b200.so")
    srcs = _sources(CSRC, (".cu", ".cuh")) + [os.path.join(INCLUDE, "ps_b200.h")]
    if force or _newer(target, srcs):
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
            "-I", INCLUDE, "-o", target, os.path.join(CSRC, "ps_b200.cu"), "-ldl"]
        subprocess.run(cmd, check=True)
    return target


def build_proto(force=False):
    """pybind11 twin of the reference's SWIG module (src/cpp/proto.i): same name, functions, shapes."""
    src = os.path.join(HOSTSRC, "proto_pybind.cpp")
    if not os.path.exists(src):
        return None
    import pybind11
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    target = os.path.join(LIBDIR, "proto" + ext)
    srcs = _sources(HOSTSRC, (".cpp", ".hpp")) + _sources(INCLUDE, (".h", ".hpp"))
    if force or _newer(target, srcs):
        cxx = os.environ.get("CXX", "g++")
        cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-I", INCLUDE, "-I",
               pybind11.get_include(), "-I", sysconfig.get_paths()["include"], src,
               os.path.join(HOSTSRC, "python_dpsolver.cpp"), os.path.join(HOSTSRC, "python_ltsssolver.cpp"),
               "-L", LIBDIR, "-lps_b200", "-Wl,-rpath,$ORIGIN", "-pthread", "-o", target]
        subprocess.run(cmd, check=True)
    return target


def build_peaks(force=False):
    """tools/peaks.cu: pipe-rate microbenchmarks bench.py uses as roofline denominators."""
    src = os.path.join(ROOT, "tools", "peaks.cu")
    target = os.path.join(LIBDIR, "ps_peaks")
    if force or _newer(target, [src]):
        subprocess.run([nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                        "-fmad=false", "-w", "-o", target, src], check=True)
    return target


def build_facade_test(force=False):
    """tests/cpp/test_facade.cpp: the reference's gtest cases against the C++ facade (run by pytest -m gpu)."""
    src = os.path.join(ROOT, "tests", "cpp", "test_facade.cpp")
    if not os.path.exists(src):
        return None
    target = os.path.join(LIBDIR, "test_facade")
    srcs = [src] + _sources(HOSTSRC, (".cpp", ".hpp")) + _sources(INCLUDE, (".h", ".hpp"))
    if force or _newer(target, srcs):
        cxx = os.environ.get("CXX", "g++")
        subprocess.run([cxx, "-O2", "-std=c++17", "-I", INCLUDE, src, os.path.join(HOSTSRC, "python_dpsolver.cpp"),
                        os.path.join(HOSTSRC, "python_ltsssolver.cpp"), "-L", LIBDIR, "-lps_b200", "-Wl,-rpath,$ORIGIN", "-pthread", "-o", target], check=True)
    return target


def build_reference_programs(force=False, ref="/root/reference/src/cpp"):
    """Drop-in demonstration: the reference's OWN demo / timer programs (DP_solver_demo.cpp, LTSS_solver_demo.cpp,
    solver_timer.cpp), compiled UNCHANGED from where they lie against the facade headers and linked to the engine.
    Only where the reference checkout exists (the binaries travel to the GPU box with the snapshot; no reference
    source enters the repository).  The six solver headers are shadowed through a temporary include directory."""
    import tempfile
    if not os.path.isdir(os.path.join(ref, "include")):
        return []
    outs = []
    with tempfile.TemporaryDirectory() as tmp:
        replaced = {"DP.hpp", "DP_impl.hpp", "score.hpp", "score_impl.hpp", "LTSS.hpp", "LTSS_impl.hpp"}
        for f in os.listdir(os.path.join(ref, "include")):
            if f not in replaced:
                os.symlink(os.path.join(ref, "include", f), os.path.join(tmp, f))
        compat = os.path.join(INCLUDE, "partitionsolvers", "compat")
        for f in os.listdir(compat):
            os.symlink(os.path.join(compat, f), os.path.join(tmp, f))
        deps = _sources(INCLUDE, (".h", ".hpp")) + [os.path.join(LIBDIR, "libps_b200.so")]
        for prog in ("DP_solver_demo", "LTSS_solver_demo", "solver_timer"):
            src = os.path.join(ref, prog + ".cpp")
            target = os.path.join(LIBDIR, "ref_" + prog)
            if force or _newer(target, deps + [src]):
                subprocess.run([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-w", "-I", tmp, "-I", INCLUDE, src,
                                "-L", LIBDIR, "-lps_b200", "-Wl,-rpath,$ORIGIN", "-pthread", "-o", target], check=True)
            outs.append(target)
    return outs


def build_all(force=False, verbose=False):
    out = [build_cuda(force, verbose), build_peaks(force)]
    out += build_reference_programs(force)
    for fn in (build_proto, build_facade_test):
        p = fn(force)
        if p:
            out.append(p)
    return out


if __name__ == "__main__":
    print("\n".join(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)))
