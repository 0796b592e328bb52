#!/usr/bin/env python
"""bench.py -- fp64 RK-stage throughput of the B200 RHS engine (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm
                                                           # (oracle port, all host threads)
  python bench.py --case channel [--n 128]                 # turbulent-channel option set (walls, src_chan)
  python bench.py --conschm 543                            # upwind compact convection on the TGV block

A "step" is one RK3 time step = 3 RK stages (filter, halo, gradient, RHS, update, primitives).
--case tgv (default, the BASELINE metric): periodic Taylor-Green vortex, N=1: 512^3 (configs[1]);
N>1: weak scaling, one 512^3 block per GPU on the block grid of astr_b200.parallel.mpisizedis.
--case channel: examples/Channel/datin/input.chl (walls 41 at jmin/jmax, grichan stretching, src_chan forcing,
Re=3000, M=0.3) at --n^3 per GPU.
Before the timed region every run checks itself: a small block per rank on the SAME block grid runs one RK3
step on the GPU(s) and on the CPU oracle and the largest relative difference goes into the line ("parity").
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

UNIT = "Mpts/s"
STAGES = 3
HM = 5
# a derivative / filter field-sweep must move 16 B per grid point per field (1 read + 1 write).
BYTES_STORE = 16.0
# algorithmic doubles per point of the two large pointwise kernels: k_visc_flux reads 12 raw
# derivatives + 9 dxi + 3 vel + T + q5 + p + J (28) and writes the 15 directional fluxes + sigma/qflux shells (~19
# amortised -> 47 with the shell planes); k_rk_update reads qsave5 + q5 + 15 G + J (26) and writes q5 + prims6 (11).
DOUBLES_VISC_FLUX, DOUBLES_RK = 47.0, 37.0


def metric_name(case):
    return "tgv_fp64_rk_stage_throughput" if case == "tgv" else "channel_fp64_rk_stage_throughput"


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


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


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
# case set-up shared by the GPU arm, the CPU arm and the parity check
# --------------------------------------------------------------------------------------
def case_params(case: str):
    if case == "channel":    # examples/Channel/datin/input.chl
        return dict(reynolds=3000.0, mach=0.3, homo=(True, False, True), lengths=(2 * np.pi, 2.0, np.pi),
                    bctype=(1, 1, 41, 41, 1, 1), twall=(0.0, 0.0, 1.0, 1.0, 0.0, 0.0), flowtype=1,
                    force=(2.5e-3, 0.0, 1e-4))
    return dict(reynolds=1600.0, mach=0.1, homo=(True, True, True), lengths=(2 * np.pi,) * 3, bctype=(1,) * 6,
                twall=(0.0,) * 6, flowtype=0, force=(0.0, 0.0, 0.0))


def oracle_case(pyoracle, case, gdims, size, conschm, deltat):
    """The CPU oracle set up like the GPU engine: same grid generator formulas, same initial state."""
    from astr_b200 import cases, decompose, refcal
    cp = case_params(case)
    th = refcal(cp["reynolds"], cp["mach"])
    c = pyoracle.Case(*gdims, blocks=size, homo=cp["homo"], reynolds=cp["reynolds"], mach=cp["mach"],
                      lengths=cp["lengths"] if case == "channel" else None, deltat=deltat)
    blocks = decompose(gdims, size, cp["homo"])
    if case == "channel":
        c.set_bc(cp["bctype"], cp["twall"])
        c.set_flow(1, cp["force"])
        for ib, b in enumerate(blocks):
            x = cases.grichan(b, gdims, cp["lengths"])
            c.set_x(np.asfortranarray(x[HM:-HM, HM:-HM, HM:-HM, :]), ib)
    c.gridgeom()
    if case == "channel":
        names = ["q1", "q2", "q3", "q4", "q5", "rho", "u", "v", "w", "prs", "tmp"]
        for ib, b in enumerate(blocks):
            x = np.stack([c.get(f"x{d + 1}", ib) for d in range(3)], axis=-1)
            q, rho, vel, prs, tmp = cases.chanini(np.asfortranarray(x), th)
            vals = [q[..., m] for m in range(5)] + [rho] + [vel[..., m] for m in range(3)] + [prs, tmp]
            for nm, v in zip(names, vals):
                c.set(nm, np.asfortranarray(v), ib)
    else:
        c.tgvini()
    if conschm == 543:
        c.set_upwind(543, True, 0.3, 0.01)
    return c, th, blocks


def gpu_engine(case, gdims, size, rank, local_rank, conschm, deltat, engine_kw=None):
    """RhsEngine of this rank's block with grid, metrics (device gridgeom) and initial state; returns
    (engine, thermo, block, (q, rho, vel, prs, tmp))."""
    from astr_b200 import RhsEngine, cases, decompose, refcal
    cp = case_params(case)
    th = refcal(cp["reynolds"], cp["mach"])
    block = decompose(gdims, size, cp["homo"])[rank]
    up_kw = {}
    if conschm == 543:     # the upwind compact / shock-capturing convection path (config 4's scheme)
        up_kw = dict(conschm=543, lchardecomp=True, bfacmpld=0.3, shkcrt=0.01)
    eng = RhsEngine(block, gdims, cp["homo"], th, deltat=deltat, device=local_rank, flowtype=cp["flowtype"],
                    bctype=cp["bctype"], twall=cp["twall"], **up_kw, **(engine_kw or {}))
    eng.set_force(cp["force"])
    return eng, th, block, cp


def init_fields(eng, case, block, gdims, th, cp):
    from astr_b200 import cases
    x = cases.grichan(block, gdims, cp["lengths"]) if case == "channel" else cases.gridcube(block, gdims, cp["lengths"])
    eng.gridgeom(x)
    state = cases.chanini(x, th) if case == "channel" else cases.tgvini(x, th)
    del x
    return state


def make_bcast(dist, torch, rank):
    def bcast(b):
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(b), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())
    return bcast


# --------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU algorithm
# --------------------------------------------------------------------------------------
def cpu_reference(case: str, n: int, steps: int, warmup: int, conschm: int):
    """Times the C++ restatement of the reference (oracle/, OpenMP over pencils, every host core this process
    may use) on an n^3 block of the workload.  The Fortran/MPI reference itself cannot be built in this image
    (no gfortran, no MPI), so kind = "port"."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    pyoracle.build()
    cores = pyoracle.set_num_threads(host_cores())      # launchers export OMP_NUM_THREADS=1: set it explicitly
    c, _, _ = oracle_case(pyoracle, case, (n, n, n), (1, 1, 1), conschm, 1e-3 * 128 / n)
    if warmup:
        c.run(warmup)
    t0 = time.perf_counter()
    c.run(steps)
    dt = time.perf_counter() - t0
    c.close()
    pts = float(n + 1) ** 3
    return pts * STAGES * steps / dt / 1e6, dt / steps * 1e3, cores


def workload_text(case, n, conschm, size):
    up = ("543c upwind compact convection (Steger-Warming, characteristic MP5, Ducros sensor) + " if conschm == 543 else "")
    if case == "channel":
        return (f"Channel {n}^3 per GPU ({n + 1}^3 nodes + 5-deep halos), walls 41 at jmin/jmax, grichan stretching, "
                f"src_chan forcing, fp64, {up}643c compact derivative + compact filter alfa=0.49, RK3, Re=3000 M=0.3; "
                f"block grid {size[0]}x{size[1]}x{size[2]}")
    return (f"TGV {n}^3 per GPU ({n + 1}^3 nodes + 5-deep halos), periodic, fp64, {up}643c compact derivative + "
            f"compact filter alfa=0.49, RK3, Re=1600 M=0.1; block grid {size[0]}x{size[1]}x{size[2]}")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_n
    # bounded sample: the whole --steps/--warmup run must end within a few minutes
    value, ms, cores = cpu_reference(args.case, n, args.steps, args.warmup, args.conschm)
    sample = (f"{n}^3 block of the workload ({n + 1}^3 nodes), {args.steps} RK3 steps after {args.warmup} warm-up, "
              f"{cores} OpenMP threads")
    line = {
        "impl": "reference", "metric": metric_name(args.case), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(args.case, args.n, args.conschm, (1, 1, 1)) +
                               f" -- CPU arm: bounded {n}^3 sample of it (throughput per point is size independent: "
                               "the working set of 58 fields exceeds every cache at either size)",
                   "flush": "working set (58 fields) larger than any cache"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# parity check on the run's own block grid (before the timed region)
# --------------------------------------------------------------------------------------
def parity_check(args, torch, dist, world, rank, local_rank, size):
    """One RK3 step of a small block per rank on the SAME block grid, GPU(s) against the CPU oracle (results are
    layout dependent: interface closures).  The oracle is the checker here, never the thing measured."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    if rank == 0:
        pyoracle.build()
    if dist is not None:
        dist.barrier()
    pyoracle.set_num_threads(max(1, host_cores() // max(world, 1)))
    pn = args.parity_n
    gdims = tuple(pn * s for s in size)
    deltat = 1e-3
    ekw = {} if args.overlap_visc < 0 else {"overlap_visc": bool(args.overlap_visc)}
    eng, th, block, cp = gpu_engine(args.case, gdims, size, rank, local_rank, args.conschm, deltat, ekw)
    if world > 1:
        eng.comm_init(world, rank, make_bcast(dist, torch, rank))
    c, _, _ = oracle_case(pyoracle, args.case, gdims, size, args.conschm, deltat)
    names = ["q1", "q2", "q3", "q4", "q5", "rho", "u", "v", "w", "prs", "tmp"]
    # identical inputs: the oracle's grid, metrics and state go to the device through the C ABI
    x = eng.empty(3)
    for d in range(3):
        x[..., d] = c.get(f"x{d + 1}", rank)
    eng.set_grid(x)
    dxi = eng.empty(9).reshape(eng.shape + (3, 3), order="F")
    for a in range(3):
        for b in range(3):
            dxi[..., a, b] = c.get(f"dxi{a + 1}{b + 1}", rank)
    eng.set_metrics(dxi, c.get("jacob", rank))
    for nm in names:
        eng.set(nm, c.get(nm, rank))
    eng.steploop(args.parity_steps)
    c.run(args.parity_steps)
    core = (slice(HM, -HM),) * 3
    groups = [["q2", "q3", "q4"], ["u", "v", "w"]]
    refs = {nm: c.get(nm, rank)[core] for nm in names}
    scale = {nm: max(float(np.abs(refs[nm]).max()), 1e-300) for nm in names}
    for grp in groups:
        s = max(scale[nm] for nm in grp)
        for nm in grp:
            scale[nm] = s
    worst = max(float(np.abs(eng.get(nm)[core] - refs[nm]).max()) / scale[nm] for nm in names)
    eng.close(); c.close()
    if dist is not None:
        t = torch.tensor([worst], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst = float(t.item())
    return {"max_rel": worst, "tolerance": 1e-12, "ok": bool(worst <= 1e-12),
            "what": f"{args.parity_steps} RK3 step(s), {pn}^3 per rank on the block grid {size[0]}x{size[1]}x{size[2]}, "
                    "q(5) + rho, vel, prs, tmp of every rank against the CPU oracle run with the same block grid "
                    "(max-norm relative per field)"}


_ORIGINAL_AFFINITY = set()     # this process's CPU set before bind_to_gpu_numa_node narrowed it


def unbind_from_numa_node():
    """Give the process its original CPU set back (rank 0, before the CPU baseline leg: the other ranks idle at a
    barrier by then, and the baseline is quoted on every host core the process may use -- at every N)."""
    if _ORIGINAL_AFFINITY:
        try:
            os.sched_setaffinity(0, sorted(_ORIGINAL_AFFINITY))
        except OSError as e:
            sys.stderr.write(f"bench.py: could not restore the CPU affinity ({e})\n")


def bind_to_gpu_numa_node(torch, local_rank):
    """Pin this rank's threads (and therefore the first-touch placement of its pinned staging buffers) to the
    NUMA node its GPU hangs off: with eight ranks staging 2 x 5.7 GB per step, buffers that all land on one node
    share that node's memory controllers and its PCIe root (round 1: e2e efficiency 0.37 at 8 GPUs)."""
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0"
        node = int(open(path + "/numa_node").read())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            _ORIGINAL_AFFINITY.update(os.sched_getaffinity(0))
            os.sched_setaffinity(0, allowed)
            return node
    except Exception as e:       # no sysfs / no permission: leave the affinity alone
        sys.stderr.write(f"bench.py: NUMA binding skipped ({e})\n")
    return None


# --------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import astr_b200
    from astr_b200 import mpisizedis
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
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    n = args.n
    size = tuple(int(v) for v in args.layout.split(",")) if args.layout else mpisizedis(world, (n, n, n))
    assert size[0] * size[1] * size[2] == world, "--layout must multiply to the number of ranks"
    parity = None
    if args.parity_n > 0:
        parity = parity_check(args, torch, dist, world, rank, local_rank, size)

    gdims = tuple(n * s for s in size)                      # weak scaling: n^3 per GPU
    ekw = {} if args.overlap_visc < 0 else {"overlap_visc": bool(args.overlap_visc)}
    eng, th, block, cp = gpu_engine(args.case, gdims, size, rank, local_rank, args.conschm, 1e-3 * 128 / n, ekw)
    if world > 1:
        eng.comm_init(world, rank, make_bcast(dist, torch, rank))
    # synthetic input: grid (gridcube / grichan) + initial state (tgvini / chanini), metrics on the device
    q, rho, vel, prs, tmp = init_fields(eng, args.case, block, gdims, th, cp)
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
    barrier()        # the ranks finish their uploads seconds apart: line them up before the first exchange
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

    # ---- roofline: every kernel family with its algorithmic bytes; the dominant one is named ------------
    peak, peak_src = measured_peaks()
    alg = {  # algorithmic bytes per launch and point: fields x 16 B (1 read + 1 write)
        "filter_i": 5 * BYTES_STORE, "filter_j": 5 * BYTES_STORE, "filter_k": 5 * BYTES_STORE,
        "grad_i": 4 * BYTES_STORE, "grad_j": 4 * BYTES_STORE, "grad_k": 4 * BYTES_STORE,
        "div_i": 5 * BYTES_STORE, "div_j": 5 * BYTES_STORE, "div_k": 5 * BYTES_STORE,
    }
    sweeps = list(alg)
    if args.conschm != 543:      # with the upwind path the div_* spans also hold the split / interface / difference kernels
        alg["visc"] = 8.0 * DOUBLES_VISC_FLUX
    alg["rk"] = 8.0 * DOUBLES_RK
    kern = {}
    tot_b = tot_ms = 0.0
    for k, bpp in alg.items():
        t_ms, cnt = prof[k]
        if cnt:
            gbs = bpp * pts * cnt / (t_ms * 1e-3) / 1e9
            kern[k] = {"ms": t_ms / cnt, "GBps": gbs, "frac": gbs / peak, "launches_per_step": cnt / args.steps,
                       "share_of_step": t_ms / ms}
            if k in sweeps:
                tot_b += bpp * pts * cnt
                tot_ms += t_ms
    if args.conschm == 543:
        for k in ("div_i", "div_j", "div_k"):
            kern.pop(k, None)
    other = {k: prof[k][0] / max(prof[k][1], 1) for k in PROFILE_CATEGORIES if k not in alg and prof[k][1]}
    # the single dominant kernel: the one with the largest share of the step
    dom = max(kern, key=lambda k: kern[k]["share_of_step"]) if kern else None
    names = {"visc": "k_visc_flux", "rk": "k_rk_update"}
    roofline = None
    if dom:
        low = min((k for k in kern if k in sweeps), key=lambda k: kern[k]["frac"], default=None)
        roofline = {"bound": "hbm", "kernel": names.get(dom, f"sweep2_kernel<{dom}>"), "achieved": kern[dom]["GBps"],
                    "peak": peak, "unit": "GB/s", "frac": kern[dom]["frac"], "traffic": ncu_traffic(dom, n),
                    "algorithmic_bytes": alg[dom] * pts, "peak_source": peak_src,
                    "all_sweeps": {"achieved": tot_b / (tot_ms * 1e-3) / 1e9 if tot_ms else None,
                                   "frac": tot_b / (tot_ms * 1e-3) / 1e9 / peak if tot_ms else None,
                                   "share_of_step": tot_ms / ms,
                                   "lowest": {"kernel": low, "frac": kern[low]["frac"]} if low else None},
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
    eng_stats = None
    try:
        eng.filterq(); eng.qswap(); eng.gradcal()
        eng_stats = eng.reduce_tgv()
    except Exception:
        pass
    eng.close()

    cpu = None
    if rank == 0 and not args.no_cpu:      # rank 0, every N: the same bounded sample on all host cores
        unbind_from_numa_node()
        v, cms, cores = cpu_reference(args.case, args.cpu_n, args.cpu_steps, 1, args.conschm)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_n}^3 block of the workload, {args.cpu_steps} RK3 steps after 1 warm-up "
                         f"({cms:.0f} ms/step), C++ restatement of the reference with OpenMP over pencils, "
                         f"{cores} threads"}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": metric_name(args.case), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(args.case, n, args.conschm, size),
                   "points_per_gpu": pts, "stages_per_step": STAGES,
                   "flush": "inputs larger than L2 (67 resident fields x 1.19 GB at 512^3)" if n >= 256 else
                            f"inputs larger than L2 only in total (67 fields x {pts * 8 / 1e6:.0f} MB)",
                   "parallelism": f"blocks{size[0]}x{size[1]}x{size[2]}"},
        "clocks": clk.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": e2e_s / max(e2e_steps, 1) * 1e3, "numa_node": numa,
                "what": "astr_gpu_upload_state(q) from pinned host + updatefvar + 3 x astr_gpu_rk_stage + "
                        "astr_gpu_download_state(q)"},
        "gpu_launches": launches,
        "parity": parity,
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
    ap.add_argument("--case", default="tgv", choices=["tgv", "channel"],
                    help="tgv: the BASELINE metric; channel: examples/Channel option set (walls 41, src_chan, grichan)")
    ap.add_argument("--n", type=int, default=512, help="grid intervals per GPU and direction")
    ap.add_argument("--cpu-n", type=int, default=256, help="size of the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--parity-n", type=int, default=48, help="per-rank size of the parity pre-check (0: skip)")
    ap.add_argument("--parity-steps", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--conschm", type=int, default=643, choices=[643, 543],
                    help="643: central compact convection (the BASELINE metric); 543: upwind compact path")
    ap.add_argument("--layout", default="", help="block grid isize,jsize,ksize (default: mpisizedis)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--overlap-visc", type=int, default=-1,
                    help="multi-block runs: sigma/qflux exchange on a side stream behind the interior stress+flux pass "
                         "(cfg.overlap_visc); -1: the library default")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
