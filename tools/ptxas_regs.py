"""Register / spill summary of the kernels in libps_b200 (ptxas -v), filtered by a substring of the kernel name.

    python tools/ptxas_regs.py [substring ...]
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-fmad=false", "-Xptxas", "-v",
       "-I", os.path.join(ROOT, "include"), "-cubin", "-o", "/dev/null",
       os.path.join(ROOT, "partitionsolvers_b200", "csrc", "ps_b200.cu")]
out = subprocess.run(cmd, capture_output=True, text=True).stderr
want = sys.argv[1:]
cur = None
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        continue
    m = re.search(r"Used (\d+) registers", line)
    if m and cur and (not want or any(w in cur for w in want)):
        sp = re.search(r"(\d+) bytes spill stores", prev) if (prev := globals().get("_prev")) else None
        print(f"{int(m.group(1)):4d} regs  {cur}  {line.split('Used')[1].strip()}" + (f"  [{sp.group(0)}]" if sp and sp.group(1) != '0' else ""))
    if "spill" in line:
        _prev = line
