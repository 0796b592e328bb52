"""N>1 host logic on CPU: world_size-2 (and 4) gloo runs of the halo-exchange plan."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, layout, homo, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mp_gloo_worker.py"), layout, homo]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "GLOO_EXCHANGE_OK" in r.stdout


@pytest.mark.parametrize("world,layout,homo,port", [
    (2, "1,1,2", "1,1,1", 29611),     # two k-slabs, periodic: both neighbours are the same rank
    (2, "2,1,1", "0,1,1", 29612),     # cut in i, physical boundaries at both outer ends
    (4, "1,2,2", "1,1,0", 29613),
])
def test_halo_plan_over_gloo(oracle, world, layout, homo, port):
    _run(world, layout, homo, port)
