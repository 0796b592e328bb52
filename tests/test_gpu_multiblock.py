"""Multi-GPU parity (needs >= 2 GPUs on the box: `gpurun --gpus 2|4|8`): one process per GPU,
NCCL halo exchange inside libastr_gpu.so, compared with the oracle on the same block grid."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _run(world, layout, homo, n, nsteps, port, extra=()):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mp_gpu_worker.py"), layout, homo, n, str(nsteps), *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "GPU_MULTIBLOCK_OK" in r.stdout


def test_two_k_slabs_periodic(oracle):
    _run(2, "1,1,2", "1,1,1", "32,32,48", 2, 29711)


def test_two_blocks_in_i_with_walls(oracle):
    _run(2, "2,1,1", "0,1,1", "48,32,32", 2, 29712)


def test_two_blocks_device_gridgeom(oracle):
    _run(2, "1,2,1", "1,1,1", "32,48,32", 1, 29713, extra=("devgeom",))


def test_four_blocks(oracle):
    _run(4, "1,2,2", "1,1,1", "32,48,48", 2, 29714)


def test_eight_blocks_as_the_reference_decomposes(oracle):
    _run(8, "2,2,2", "1,1,1", "48,48,48", 2, 29715)


def test_channel_split_in_y_and_x(oracle):
    # walls owned by different ranks, src_chan bulk integrals summed over ranks (psum -> ncclAllReduce)
    _run(4, "2,2,1", "1,0,1", "48,48,24", 2, 29716, extra=("channel",))


def test_channel_two_blocks(oracle):
    _run(2, "1,2,1", "1,0,1", "32,48,24", 2, 29717, extra=("channel",))


def test_upwind_two_blocks(oracle):
    # conschm 543c across an interface: ssf halo exchange, interface closures of the compact flux
    _run(2, "2,1,1", "1,1,1", "48,32,32", 2, 29718, extra=("upwind",))


def test_upwind_eight_blocks_with_walls(oracle):
    _run(8, "2,2,2", "0,1,0", "48,48,48", 1, 29719, extra=("upwind",))


def test_two_blocks_register_engine_walls_in_i(oracle):
    # blocks of 56 x 48 x 56 intervals: every line runs on the register engine (sweep2.cu); the wall direction is
    # split, so block 0 has ntype 1 (wall at node 0, interface at im) and block 1 ntype 2 -- the two closures a
    # single block cannot have
    _run(2, "2,1,1", "0,1,1", "112,48,56", 1, 29720)


def test_two_blocks_register_engine_walls_in_k(oracle):
    _run(2, "1,1,2", "1,1,0", "48,56,112", 1, 29721)


def test_two_blocks_overlapped_stress_exchange(oracle):
    # cfg.overlap_visc = 1: shell pass -> sigma/qflux exchange on the side stream -> interior stress+flux pass -> join
    _run(2, "1,1,2", "1,1,1", "32,32,48", 2, 29722, extra=("overlap",))


def test_eight_blocks_overlapped_stress_exchange_with_walls(oracle):
    _run(8, "2,2,2", "1,0,1", "48,48,48", 2, 29723, extra=("overlap",))
