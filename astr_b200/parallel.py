"""Block decomposition of the reference: mpisizedis / parapp / parallelini
(src/parallel.F90:199-484, :919-1248).  Host logic only; one Block per GPU."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple


@dataclass
class Block:
    rank: int
    rk: Tuple[int, int, int]          # irk, jrk, krk
    size: Tuple[int, int, int]        # isize, jsize, ksize
    dims: Tuple[int, int, int]        # im, jm, km
    g0: Tuple[int, int, int]          # ig0, jg0, kg0 (global index of local node 0)
    npdc: Tuple[int, int, int]
    s: Tuple[int, int, int]           # is, js, ks
    e: Tuple[int, int, int]           # ie, je, ke
    nbr: List[int] = field(default_factory=list)   # i-,i+,j-,j+,k-,k+ ; -1 = MPI_PROC_NULL


def mpisizedis(nranks: int, dims: Sequence[int], strict3d: bool = False) -> Tuple[int, int, int]:
    """(isize,jsize,ksize) as src/parallel.F90:199-323: among the factorisations of nranks,
    minimise ja*ka*isize + ia*ka*jsize + ia*ja*ksize (first minimum in the reference's loop
    order wins).

    strict3d=True reproduces the reference's extra rule that a 3-D run needs all three
    factors > 1 (:289-297), which is why it cannot run 3-D on 2 or 4 ranks.  The default
    lifts that rule (the operators never needed it) so 2 GPUs -> 1x1x2 slabs, 4 -> 1x2x2;
    on ties the cut goes to the slowest index (contiguous halo planes).
    """
    ia, ja, ka = dims
    if nranks == 1:
        return (1, 1, 1)
    ndims = 1 if (ja == 0 and ka == 0) else (2 if ka == 0 else 3)
    kaa = ka + 1 if ka == 0 else ka
    factors = [nranks // i for i in range(1, nranks + 1) if nranks % i == 0]   # descending
    best, best_cost = None, 2 ** 62
    for f1 in factors:
        for f2 in factors:
            for f3 in factors:
                if f1 * f2 * f3 != nranks:
                    continue
                if ndims == 1 and (f2 != 1 or f3 != 1):
                    continue
                if ndims == 2 and (f3 != 1 or f2 == 1):
                    continue
                if ndims == 3 and strict3d and not (f1 > 1 and f2 > 1 and f3 > 1):
                    continue
                cost = ja * kaa * f1 + ia * kaa * f2 + ia * ja * f3
                if strict3d:
                    better = cost < best_cost
                else:   # ties: prefer cutting k, then j
                    better = (cost, f1, f2) < (best_cost, *(best[:2] if best else (0, 0)))
                if better:
                    best, best_cost = (f1, f2, f3), cost
    if best is None:
        raise ValueError(f"size of ranks can not be allocated: {dims} over {nranks} ranks")
    return best


def decompose(dims: Sequence[int], size: Sequence[int], homo: Sequence[bool]) -> List[Block]:
    """parapp (:356-484) + parallelini (:919-1248).  rank = krk*isize*jsize + jrk*isize + irk."""
    n = list(dims)
    mp, off = [], []
    for d in range(3):
        sz = size[d]
        m = [n[d] // sz] * sz
        rem = n[d] % sz
        for r in range(sz - 1, sz - 1 - rem, -1):   # the last `rem` ranks get one more (:399-405)
            m[r] += 1
        o = [0] * sz
        for r in range(1, sz):
            o[r] = o[r - 1] + m[r - 1]
        mp.append(m)
        off.append(o)
    isize, jsize, ksize = size
    blocks = []
    for krk in range(ksize):
        for jrk in range(jsize):
            for irk in range(isize):
                rk = (irk, jrk, krk)
                dm = tuple(mp[d][rk[d]] for d in range(3))
                g0 = tuple(off[d][rk[d]] for d in range(3))
                npdc, s, e, nbr = [], [], [], []

                def rank_of(d, rr, rk=rk):
                    cc = list(rk)
                    cc[d] = rr
                    return cc[2] * isize * jsize + cc[1] * isize + cc[0]

                for d in range(3):
                    sz, r, rm = size[d], rk[d], size[d] - 1
                    if homo[d]:                       # :1042-1066
                        npdc.append(3); s.append(0); e.append(dm[d])
                        if sz == 1:
                            nbr += [-1, -1]
                        else:
                            nbr += [rank_of(d, rm if r == 0 else r - 1), rank_of(d, 0 if r == rm else r + 1)]
                    elif sz == 1:                     # :1071-1077 (is..ie unset in the reference)
                        npdc.append(4); s.append(1); e.append(dm[d] - 1); nbr += [-1, -1]
                    elif r == 0:
                        npdc.append(1); s.append(1); e.append(dm[d]); nbr += [-1, rank_of(d, r + 1)]
                    elif r == rm:
                        npdc.append(2); s.append(0); e.append(dm[d] - 1); nbr += [rank_of(d, r - 1), -1]
                    else:
                        npdc.append(3); s.append(0); e.append(dm[d]); nbr += [rank_of(d, r - 1), rank_of(d, r + 1)]
                blocks.append(Block(rank=rank_of(0, irk), rk=rk, size=tuple(size), dims=dm, g0=g0,
                                    npdc=tuple(npdc), s=tuple(s), e=tuple(e), nbr=nbr))
    return blocks


# --------------------------------------------------------------------------------------
# Halo-exchange plan: which messages one block posts for one direction.  This is the
# schedule libastr_gpu.so executes with grouped ncclSend/ncclRecv (csrc/api.cu
# exchange_dir); it is stated here so that host code (and the gloo tests) can reason about
# it without a GPU.
# --------------------------------------------------------------------------------------
@dataclass
class HaloMessage:
    peer: int            # rank of the neighbour
    send_side: int       # 0: my planes l0..l1 (low side), 1: my planes dm-l (high side)
    recv_side: int       # 1: lands in my high halo dm+l, 0: lands in my low halo -l
    count: int           # doubles


def halo_plan(block: Block, d: int, mode: str, nfields: int) -> Tuple[List[HaloMessage], List[HaloMessage]]:
    """(sends, recvs) of direction d in posting order.

    mode: "swap"  -> planes 1..hm          (dataswap, src/parallel.F90:4132-4370)
          "qswap" -> planes 0..hm, plane 0 averaged with the neighbour's (qswap, :4848-5318)
          "sync"  -> plane 0 only, averaged (array3d_sync, :3725-3934)
    Low-side planes go to the low neighbour (which puts them in its HIGH halo) and vice
    versa.  When both neighbours are the same rank (two blocks, periodic) the peer's first
    message is its low-side buffer, so the receive for MY HIGH halo is posted first.
    """
    hm = 5
    l0, l1 = {"swap": (1, hm), "qswap": (0, hm), "sync": (0, 0)}[mode]
    o = [x for x in range(3) if x != d]
    n1, n2 = block.dims[o[0]] + 1, block.dims[o[1]] + 1
    cnt = (l1 - l0 + 1) * n1 * n2 * nfields
    lo, hi = block.nbr[2 * d], block.nbr[2 * d + 1]
    sends, recvs = [], []
    if lo >= 0:
        sends.append(HaloMessage(lo, 0, 1, cnt))
    if hi >= 0:
        sends.append(HaloMessage(hi, 1, 0, cnt))
    if hi >= 0:
        recvs.append(HaloMessage(hi, 0, 1, cnt))   # the high neighbour's low-side planes -> my high halo
    if lo >= 0:
        recvs.append(HaloMessage(lo, 1, 0, cnt))   # the low neighbour's high-side planes -> my low halo
    return sends, recvs
