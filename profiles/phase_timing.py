"""Per-phase cycle breakdown of the eigenfunction step kernels (profiling build with -DCVF_PHASE_TIMERS).

    python profiles/phase_timing.py [frames]

Builds a SEPARATE library (profiles/_build/libcvf_prof.so; the product library has no timers), points the package at
it, runs pass 1 and pass 2 on the C3 workload and prints the share of SM cycles (thread 0 of every CTA) per phase.
"""
import ctypes as C
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "colvars-finder_b200")]
out = os.path.join(ROOT, "profiles", "_build", "libcvf_prof.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
srcs = sorted(glob.glob(os.path.join(ROOT, "colvars-finder_b200", "csrc", "*.cu")))
subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-DCVF_PHASE_TIMERS",
                       "-Xcompiler", "-fPIC", "-shared", "-o", out] + srcs)
from colvarsfinder import _lib  # noqa: E402
_lib.LIB_PATH = out
import torch  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
wl = sys.argv[2] if len(sys.argv) > 2 else "c3"
dev = torch.device("cuda", 0)
step, X, w, task = bench.build_workload(wl, n, dev, seed=1)
lib = _lib.lib()
lib.cvf_debug_phase_cycles.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_ulonglong * 16)()
names = ["load", "preprocess", "forward", "reverse", "J phase", "tangent", "outer(tangent)", "2nd reverse", "outer(all)", "stats",
         "setup"]
ctx = task._ctx
for label, fn in (("pass 1 (stats)", lambda: ctx.stats(X, w)),):
    fn(); torch.cuda.synchronize(); lib.cvf_debug_phase_cycles(buf, 1)
    y, stats = fn(); torch.cuda.synchronize(); lib.cvf_debug_phase_cycles(buf, 1)
    tot = sum(buf[:11])
    print(label, "total Mcycles/CTA %.2f" % (tot / 148 / 1e6))
    for i, nm in enumerate(names):
        if buf[i]:
            print("  %-16s %5.1f%%" % (nm, 100.0 * buf[i] / tot))
comb = ctx.combine(stats)
ctx.grads(X, w, y, comb); torch.cuda.synchronize(); lib.cvf_debug_phase_cycles(buf, 1)
ctx.grads(X, w, y, comb); torch.cuda.synchronize(); lib.cvf_debug_phase_cycles(buf, 1)
tot = sum(buf[:11])
print("pass 2 (grad) total Mcycles/CTA %.2f" % (tot / 148 / 1e6))
for i, nm in enumerate(names):
    if buf[i]:
        print("  %-16s %5.1f%%" % (nm, 100.0 * buf[i] / tot))
