"""Multi-GPU path of the tile driver: tiles sharded round-robin over ranks, no inter-step communication,
ONE all_gather_into_tensor of the decoded tiles, blend kernel — launched with torchrun at world size 2 (and 4 when the box
has them) and compared with the single-process run: the stitched image must be bit-identical for every world size."""
import os
import re
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world: int) -> str:
    env = dict(os.environ)
    env.pop("RANK", None), env.pop("WORLD_SIZE", None), env.pop("LOCAL_RANK", None)
    if world == 1:
        cmd = [sys.executable, WORKER]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), WORKER]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert res.returncode == 0, f"world {world} failed:\n{res.stdout[-2000:]}\n{res.stderr[-4000:]}"
    m = re.search(r"CRC ([0-9a-f]{8}) world (\d+)", res.stdout)
    assert m and int(m.group(2)) == world, res.stdout[-2000:]
    return m.group(1)


def test_restore_image_bit_identical_across_world_sizes(cuda_lib):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (tile sharding + NCCL all-gather)")
    crc1 = _run(1)
    crc2 = _run(2)
    assert crc1 == crc2, f"stitched image depends on the world size: {crc1} (1 rank) vs {crc2} (2 ranks)"
    if n >= 4:
        assert _run(4) == crc1
