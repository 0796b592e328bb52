"""Host<->device transfer rates on the bench box: plain contiguous pinned copies (torch) against the
pitched per-field copies of astr_gpu_upload_state / download_state at 512^3."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astr_b200 import RhsEngine, decompose, refcal

n = 512
m = n + 11
pin = torch.empty((5, m, m, m), dtype=torch.float64, pin_memory=True)
pin.zero_()
dev = torch.empty((m, m, m), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
for name, fn in (("H2D 1-D", lambda k: dev.copy_(pin[k], non_blocking=True)), ("D2H 1-D", lambda k: pin[k].copy_(dev, non_blocking=True))):
    fn(0); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(5):
        fn(k)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name}: {5 * m ** 3 * 8 / dt / 1e9:.1f} GB/s")
del dev
block = decompose((n, n, n), (1, 1, 1), (True,) * 3)[0]
eng = RhsEngine(block, (n, n, n), (True,) * 3, refcal(1600.0, 0.1), device=0)
q = pin.numpy().reshape(-1).reshape((m, m, m, 5), order="F")
eng.upload_state(q=q); eng.synchronize()
for name, fn in (("upload_state(q)", lambda: eng.upload_state(q=q)), ("download_state(q)", lambda: eng.download_state(q=q))):
    t0 = time.perf_counter()
    fn(); eng.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name}: {5 * m ** 3 * 8 / dt / 1e9:.1f} GB/s  ({dt * 1e3:.1f} ms)")
eng.close()
