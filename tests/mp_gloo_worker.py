"""world_size-2+ gloo worker (CPU): executes astr_b200.parallel.halo_plan with numpy pack/unpack
over torch.distributed send/recv and checks every block's halos against the oracle's
multi-block exchange (oracle/solver.cpp exchange_dir)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle  # noqa: E402
from astr_b200 import decompose, halo_plan  # noqa: E402

HM = 5


def planes(a, d, idx):
    sl = [slice(HM, -HM)] * 3
    sl[d] = idx
    return a[tuple(sl)]


def exchange(a, block, d, mode):
    """numpy restatement of pw_pack / grouped send-recv / pw_unpack for ONE field."""
    dm = block.dims[d]
    l0, l1 = {"swap": (1, HM), "qswap": (0, HM), "sync": (0, 0)}[mode]
    sends, recvs = halo_plan(block, d, mode, 1)
    reqs, rbufs = [], []
    for m in sends:
        ls = [HM + (dm - l if m.send_side else l) for l in range(l0, l1 + 1)]
        buf = np.ascontiguousarray(np.stack([planes(a, d, i) for i in ls]))
        assert buf.size == m.count
        reqs.append(dist.isend(torch.from_numpy(buf), m.peer))
    for m in recvs:
        o = [x for x in range(3) if x != d]
        buf = np.empty((l1 - l0 + 1, block.dims[o[0]] + 1, block.dims[o[1]] + 1))
        reqs.append(dist.irecv(torch.from_numpy(buf), m.peer))
        rbufs.append((m, buf))
    for r in reqs:
        r.wait()
    for m, buf in rbufs:
        for k, l in enumerate(range(l0, l1 + 1)):
            node = dm + l if m.recv_side else -l
            tgt = planes(a, d, HM + node)
            if l == 0:
                tgt[...] = 0.5 * (tgt + buf[k])
            else:
                tgt[...] = buf[k]


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    layout = tuple(int(v) for v in sys.argv[1].split(","))
    homo = tuple(bool(int(v)) for v in sys.argv[2].split(","))
    n = (24, 20, 28)
    assert layout[0] * layout[1] * layout[2] == world
    blocks = decompose(n, layout, homo)
    blk = blocks[rank]
    c = pyoracle.Case(*n, blocks=layout, homo=homo)
    c.gridgeom(); c.tgvini()
    rng = np.random.default_rng(5)
    for ib in range(world):          # same random fields in every process
        for name in ("q1", "q2"):
            a = c.get(name, ib)
            a[...] = rng.standard_normal(a.shape)
            c.set(name, a, ib)
    mine = {name: c.get(name, rank).copy() for name in ("q1", "q2")}
    # oracle: filterq's dataswap is exercised through qswap (QSWAP mode) on q; SWAP mode through
    # the oracle's filterq would also filter, so check QSWAP here and SWAP via a second case
    c.qswap()
    for d in range(3):
        if layout[d] > 1:
            for name in mine:
                exchange(mine[name], blk, d, "qswap")
        elif homo[d]:
            for name in mine:       # single block periodic: local wrap + average
                a, dm = mine[name], blk.dims[d]
                for l in range(1, HM + 1):
                    planes(a, d, HM - l)[...] = planes(a, d, HM + dm - l)
                    planes(a, d, HM + dm + l)[...] = planes(a, d, HM + l)
                v = 0.5 * (planes(a, d, HM) + planes(a, d, HM + dm))
                planes(a, d, HM)[...] = v
                planes(a, d, HM + dm)[...] = v
    for name in mine:
        ref = c.get(name, rank)
        assert np.array_equal(mine[name], ref), f"rank {rank} {name} differs from the oracle exchange"
    c.close()
    dist.barrier()
    if rank == 0:
        print("GLOO_EXCHANGE_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
