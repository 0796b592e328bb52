"""One rank = one GPU = one block: runs the CUDA engine on its block of a multi-block case
(NCCL halo exchange inside libastr_gpu.so) and checks it against the oracle run with the
SAME block layout (results are layout dependent)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyoracle  # noqa: E402
from astr_b200 import RhsEngine, decompose, refcal  # noqa: E402
from gpu_common import PRIMS, QS, GROUPS, auto_shkcrt, channel_state, channel_x, clean_metrics  # noqa: E402

HM = 5


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", rank))
    layout = tuple(int(v) for v in sys.argv[1].split(","))
    homo = tuple(bool(int(v)) for v in sys.argv[2].split(","))
    n = tuple(int(v) for v in sys.argv[3].split(","))
    nsteps = int(sys.argv[4])
    modes = set(sys.argv[5:])
    device_metrics = "devgeom" in modes
    channel, upwind = "channel" in modes, "upwind" in modes
    blocks = decompose(n, layout, homo)
    blk = blocks[rank]
    reynolds, mach = (3000.0, 0.3) if channel else (1600.0, 0.1)
    th = refcal(reynolds, mach)
    lengths = (2 * np.pi, 2.0, np.pi) if channel else None
    c = pyoracle.Case(*n, blocks=layout, homo=homo, reynolds=reynolds, mach=mach, lengths=lengths)
    kw = {}
    if channel:      # examples/Channel/datin/input.chl at reduced size, split over the ranks
        kw = dict(flowtype=1, bctype=(1, 1, 41, 41, 1, 1), twall=(0, 0, 1.0, 1.0, 0, 0))
        force = (2.5e-3, 0.0, 1e-4)
        c.set_bc(kw["bctype"], kw["twall"]); c.set_flow(1, force)
        xg = channel_x(n, lengths)
        for ib in range(world):
            b = c.block_info(ib)
            g0, dims = b["g0"], (b["im"], b["jm"], b["km"])
            c.set_x(xg[tuple(slice(g0[d], g0[d] + dims[d] + 1) for d in range(3))], ib)
    c.gridgeom()
    if channel:
        for ib in range(world):
            channel_state(c, th, ib=ib)
    else:
        c.tgvini()
    rng = np.random.default_rng(99)
    for ib in range(world):
        for name in QS:
            a = c.get(name, ib)
            a *= 1.0 + (1e-2 if upwind else 1e-3) * rng.standard_normal(a.shape)
            c.set(name, a, ib)
    # shared interface nodes must hold one value: let the oracle's qswap average them first
    c.qswap(); c.updatefvar()
    if upwind:       # conschm 543c with characteristic decomposition and the Ducros sensor
        clean_metrics(c)
        shk = auto_shkcrt(c)
        c.set_upwind(543, True, 0.3, shk)
        kw.update(conschm=543, lchardecomp=True, bfacmpld=0.3, shkcrt=shk)
    if "overlap" in modes:      # cfg.overlap_visc: sigma/qflux exchange on the side stream behind the interior pass
        kw.update(overlap_visc=True)
    eng = RhsEngine(blk, n, homo, th, deltat=1e-3, device=local, **kw)
    if channel:
        eng.set_force(force)

    def bcast(b):
        obj = [b]
        dist.broadcast_object_list(obj, 0)
        return obj[0]
    eng.comm_init(world, rank, bcast)
    if device_metrics:
        x = eng.empty(3)
        for d in range(3):
            x[..., d] = c.get(f"x{d + 1}", rank)
        eng.gridgeom(x)
    else:
        dxi = eng.empty(9).reshape(eng.shape + (3, 3), order="F")
        for a in range(3):
            for b in range(3):
                dxi[..., a, b] = c.get(f"dxi{a + 1}{b + 1}", rank)
        eng.set_metrics(dxi, c.get("jacob", rank))
        if channel:
            x = eng.empty(3)
            for d in range(3):
                x[..., d] = c.get(f"x{d + 1}", rank)
            eng.set_grid(x)
    for name in QS + PRIMS:
        eng.set(name, c.get(name, rank))
    hist = []
    for step in range(nsteps):
        for rk in (1, 2, 3):
            if rk == 1:
                eng.filterq(); eng.boucon(); eng.qswap(); eng.gradcal()
                ke, en = eng.reduce_tgv()
                t = torch.tensor([ke, en], dtype=torch.float64)
                dist.all_reduce(t)                       # psum
                hist.append(t.numpy().copy())
                eng.rhscal(); eng.rk_update(1); eng.updatefvar()
            else:
                eng.rk_stage(rk)
    h = c.run(nsteps)
    core = (slice(HM, -HM),) * 3
    names = QS + PRIMS + (["jacob", "dxi11", "dxi22", "dxi33", "dxi12"] if device_metrics else [])
    refs = {nm: c.get(nm, rank) for nm in names}
    scales = {nm: max(np.abs(refs[nm][core]).max(), 1e-300) for nm in names}
    for grp in GROUPS:
        pres = [x for x in grp if x in scales]
        if pres:
            s = max(scales[x] for x in pres)
            for x in pres:
                scales[x] = s
    worst = {nm: float(np.abs(eng.get(nm)[core] - refs[nm][core]).max() / scales[nm]) for nm in names}
    bad = {k: v for k, v in worst.items() if not v <= (1e-11 if k.startswith(("dxi", "jacob")) else 1e-12)}
    assert not bad, f"rank {rank}: {bad}"
    # statistics history (psum of block sums, normalised as statistic.F90)
    cnt = float(n[0] * n[1] * n[2])
    ke = np.array([0.5 * x[0] / cnt for x in hist])
    en = np.array([0.5 * x[1] / cnt for x in hist])
    assert np.abs(ke - h[:, 2]).max() < 1e-12 * h[0, 2], (ke, h[:, 2])
    assert np.abs(en - h[:, 3]).max() < 1e-12 * h[0, 3], (en, h[:, 3])
    eng.close(); c.close()
    dist.barrier()
    if rank == 0:
        print("GPU_MULTIBLOCK_OK", max(worst.values()))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
