"""The real multi-GPU path on GPUs: needs >= 2 CUDA devices (`gpurun --gpus 2`), skipped otherwise.

One stream sharded over N ranks with tail replication, encoded on the GPUs, gathered with NCCL; the container must equal
the single-GPU container AND the unmodified reference's streams, and decode back (tests/sharded_worker.py)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.timeout(900)
@pytest.mark.parametrize("kind,total,block,ext", [
    ("text", (96 << 20) + 12345, 262144, 0),        # cfg-2 shape; ragged last block
    ("rep8", 128 << 20, 1 << 20, 0),                # cfg-4 shape: 1 MiB blocks of the 8-byte period
    ("text", (8 << 20) + 7, 4 << 20, 1),            # 3 blocks of 4 MiB over the ranks (uneven split), extension format
    ("random", (16 << 20) + 1, 65536, 0),
])
def test_sharded_encode_gather_equals_single_gpu_and_reference(kind, total, block, ext):
    n = _gpus()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs (found {n}); run under `gpurun --gpus 2`")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "sharded_worker.py"), kind, str(total), str(block), str(ext)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=800, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, (r.returncode, r.stdout[-1500:], r.stderr[-3000:])
    res = json.loads(lines[-1])
    assert res["ok"] and res["equals_single_gpu_container"] and res["equals_reference_streams"] and res["sharded_decode_round_trip"], res
