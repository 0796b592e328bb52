#!/usr/bin/env python
"""bench.py -- TGV fp64 RK-stage throughput of the B200 RHS engine (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm
                                                           # (oracle port, all host threads)

A "step" is one RK3 time step = 3 RK stages (filter, halo, gradient, RHS, update, primitives)
of a periodic Taylor-Green-vortex block.  N=1: 512^3 (configs[1]); N>1: weak scaling, one
512^3 block per GPU on the block grid of astr_b200.parallel.mpisizedis, NCCL halo exchange.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

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

METRIC = "tgv_fp64_rk_stage_throughput"
UNIT = "Mpts/s"
STAGES = 3
# a derivative / filter field-sweep must move 16 B per grid point per field
# (1 read + 1 write); with an accumulate epilogue (qrhs += d/dxi) 24 B.
BYTES_STORE, BYTES_ADD = 16.0, 24.0


def ncu_traffic(kernel: str, n: int):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/ncu_traffic.json),
    or None when that kernel / block size has not been captured."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if int(d.get("n", 0)) == n:
            v = d["bytes_per_launch"].get(kernel)
            return float(v) if v is not None else None
    except Exception:
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            if "hbm_gbs" in d:
                return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].startswith("Active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU algorithm
# --------------------------------------------------------------------------------------
def cpu_reference(n: int, steps: int, warmup: int):
    """Times the C++ restatement of the reference (oracle/, OpenMP over pencils, every host
    core) on an n^3 periodic TGV block.  The Fortran/MPI reference itself cannot be built in
    this image (no gfortran, no MPI), so kind = "port"."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    pyoracle.build()
    c = pyoracle.Case(n, n, n)
    c.gridgeom(); c.tgvini()
    if warmup:
        c.run(warmup)
    t0 = time.perf_counter()
    c.run(steps)
    dt = time.perf_counter() - t0
    cores = pyoracle.num_threads()
    c.close()
    pts = float(n + 1) ** 3
    return pts * STAGES * steps / dt / 1e6, dt / steps * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_n
    value, ms, cores = cpu_reference(n, args.steps, args.warmup)
    sample = f"{n}^3 periodic TGV block ({n + 1}^3 nodes), {args.steps} RK3 steps after {args.warmup} warm-up"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"TGV {n}^3 periodic fp64 RK3 (643c + filter 0.49), bounded CPU sample of the "
                               "512^3 workload: throughput per point is size independent",
                   "flush": "working set (58 fields) larger than any cache"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import astr_b200
    from astr_b200 import RhsEngine, cases, decompose, mpisizedis, refcal
    from astr_b200.lib import PROFILE_CATEGORIES

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # NCCL (and anything else native) may write banners to fd 1: stdout must carry ONE JSON line only, so
    # fd 1 points at stderr for the whole run and the JSON line goes to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    n = args.n
    size = tuple(int(v) for v in args.layout.split(",")) if args.layout else mpisizedis(world, (n, n, n))
    assert size[0] * size[1] * size[2] == world, "--layout must multiply to the number of ranks"
    gdims = tuple(n * s for s in size)                      # weak scaling: n^3 per GPU
    homo = (True, True, True)
    block = decompose(gdims, size, homo)[rank]
    th = refcal(1600.0, 0.1)
    up_kw = {}
    if args.conschm == 543:     # the upwind compact / shock-capturing convection path (config 4's scheme) on the TGV block
        up_kw = dict(conschm=543, lchardecomp=True, bfacmpld=0.3, shkcrt=0.01)
    eng = RhsEngine(block, gdims, homo, th, deltat=1e-3 * 128 / n, device=local_rank, **up_kw)
    if world > 1:
        def bcast(b):
            t = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                t.copy_(torch.frombuffer(bytearray(b), dtype=torch.uint8))
            dist.broadcast(t, 0)
            return bytes(t.cpu().numpy().tobytes())
        eng.comm_init(world, rank, bcast)

    # synthetic input: uniform cube grid (gridcube) + Taylor-Green vortex (tgvini), metrics on device
    lengths = tuple(2 * np.pi for _ in range(3))
    x = cases.gridcube(block, gdims, lengths)
    eng.gridgeom(x)
    q, rho, vel, prs, tmp = cases.tgvini(x, th)
    del x
    shp = eng.shape
    qpin = torch.empty((5,) + shp[::-1], dtype=torch.float64, pin_memory=True)     # Fortran (i fastest)
    qhost = qpin.numpy().reshape(-1).reshape(shp + (5,), order="F")
    qhost[...] = q
    del q
    eng.upload_state(qhost, rho, vel, prs, tmp)
    del rho, vel, prs, tmp
    eng.synchronize()
    pts = float(block.dims[0] + 1) * (block.dims[1] + 1) * (block.dims[2] + 1)

    def barrier():
        eng.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- device-resident timing --------------------------------------------------------------
    eng.steploop(args.warmup)
    barrier()
    eng.set_profile(True)
    l0 = eng.kernel_launches()
    with ClockSampler(local_rank) as clk:
        barrier()
        ms = eng.steploop_timed(args.steps)
        barrier()
    launches = eng.kernel_launches() - l0
    prof = eng.get_profile()
    eng.set_profile(False)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = pts * world * STAGES * args.steps / (ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel family: the batched line-solve sweep ---------------
    peak, peak_src = measured_peaks()
    nst = STAGES * args.steps
    alg = {  # algorithmic bytes per launch: fields x points x 16 B (1 read + 1 write)
        "filter_i": 5 * BYTES_STORE, "filter_j": 5 * BYTES_STORE, "filter_k": 5 * BYTES_STORE,
        "grad_i": 4 * BYTES_STORE, "grad_j": 4 * BYTES_STORE, "grad_k": 4 * BYTES_STORE,
        "div_i": 5 * BYTES_STORE, "div_j": 5 * BYTES_STORE, "div_k": 5 * BYTES_STORE,
    }
    kern = {}
    tot_b = tot_ms = 0.0
    for k, bpp in alg.items():
        t_ms, cnt = prof[k]
        if cnt:
            gbs = bpp * pts * cnt / (t_ms * 1e-3) / 1e9
            kern[k] = {"ms": t_ms / cnt, "GBps": gbs, "frac": gbs / peak}
            tot_b += bpp * pts * cnt
            tot_ms += t_ms
    other = {k: prof[k][0] / max(prof[k][1], 1) for k in PROFILE_CATEGORIES if k not in alg and prof[k][1]}
    # the single dominant kernel: the one with the largest share of the step
    dom = max(kern, key=lambda k: kern[k]["ms"] * prof[k][1]) if kern else None
    roofline = None
    if dom:
        roofline = {"bound": "hbm", "kernel": f"sweep_kernel<{dom}>", "achieved": kern[dom]["GBps"], "peak": peak,
                    "unit": "GB/s", "frac": kern[dom]["frac"], "traffic": ncu_traffic(dom, n),
                    "algorithmic_bytes": alg[dom] * pts, "peak_source": peak_src,
                    "all_sweeps": {"achieved": tot_b / (tot_ms * 1e-3) / 1e9, "frac": tot_b / (tot_ms * 1e-3) / 1e9 / peak,
                                   "share_of_step": tot_ms / (ms * 1.0)},
                    "per_kernel": kern, "other_ms_per_launch": other}

    # ---- end-to-end through the C ABI with HOST buffers --------------------------------------
    # one step = upload q from pinned host memory, rebuild primitives, 3 RK stages, download q.
    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, args.e2e_steps))
    h2d = d2h = 5 * int(np.prod(shp)) * 8
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.upload_state(q=qhost)
        eng.updatefvar()
        eng.steploop(1)
        eng.download_state(q=qhost)
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = pts * world * STAGES * e2e_steps / e2e_s / 1e6 if e2e_steps else None
    ke = eng_stats = None
    try:
        eng.filterq(); eng.qswap(); eng.gradcal()
        eng_stats = eng.reduce_tgv()
    except Exception:
        pass
    eng.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cms, cores = cpu_reference(args.cpu_n, args.cpu_steps, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_n}^3 periodic TGV block, {args.cpu_steps} RK3 steps after 1 warm-up "
                         f"({cms:.0f} ms/step), C++ restatement of the reference with OpenMP over pencils"}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"TGV {n}^3 per GPU ({n + 1}^3 nodes + 5-deep halos), periodic, fp64, "
                               + ("543c upwind compact convection (Steger-Warming, characteristic MP5, Ducros sensor) + "
                                  if args.conschm == 543 else "") +
                               f"643c compact "
                               f"derivative + compact filter alfa=0.49, RK3, Re=1600 M=0.1; block grid "
                               f"{size[0]}x{size[1]}x{size[2]}",
                   "points_per_gpu": pts, "stages_per_step": STAGES,
                   "flush": "inputs larger than L2 (67 resident fields x 1.19 GB)",
                   "parallelism": f"blocks{size[0]}x{size[1]}x{size[2]}"},
        "clocks": clk.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": e2e_s / max(e2e_steps, 1) * 1e3,
                "what": "astr_gpu_upload_state(q) from pinned host + updatefvar + 3 x astr_gpu_rk_stage + "
                        "astr_gpu_download_state(q)"},
        "gpu_launches": launches,
        "tgv_sums": list(eng_stats) if eng_stats else None,
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=512, help="grid intervals per GPU and direction")
    ap.add_argument("--cpu-n", type=int, default=128, help="size of the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--conschm", type=int, default=643, choices=[643, 543],
                    help="643: central compact convection (the BASELINE metric); 543: upwind compact path")
    ap.add_argument("--layout", default="", help="block grid isize,jsize,ksize (default: mpisizedis)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
