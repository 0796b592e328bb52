"""Profiling driver: a few launches of each line-solve sweep at the bench size (no setup
beyond astr_gpu_init), for `ncu -k regex:sweep_kernel`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astr_b200 import lib as _L
if len(sys.argv) > 3:      # experiment builds of the same sources (make skel / nofma)
    _L.use_debug_library(os.path.abspath(sys.argv[3]))
from astr_b200 import RhsEngine, decompose, refcal

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
homo = (True, True, True)
b = decompose((n, n, n), (1, 1, 1), homo)[0]
eng = RhsEngine(b, (n, n, n), homo, refcal(1600.0, 0.1), device=0)
for op in ((0, 1, 10, 11) if os.environ.get('PROF_SWAP') else (0, 1)):
    for d in (0, 1, 2):
        ms = eng.bench_sweep(op, d, 5, max(iters, 1)) if iters > 0 else eng.bench_sweep(op, d, 5, 0)
        gbs = 5 * 16.0 * (n + 1) ** 3 / (ms * 1e-3) / 1e9 if ms > 0 and ms == ms and ms != float('inf') else 0.0
        print(f"op={('deriv','filter')[op % 10]}{' (placement swapped)' if op >= 10 else ''} dir={'ijk'[d]} {ms:.3f} ms/launch  {gbs:.0f} GB/s", flush=True)
eng.close()
