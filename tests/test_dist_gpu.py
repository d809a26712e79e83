"""Multi-GPU parity on real kernels (skipped on a 1-GPU box): 2 NCCL ranks shard one stream set, run the engine and gather
the 24-byte records; the gathered result must equal the single-GPU result bit for bit (SURVEY.md 4 iv, 8e)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('config', ['c2', 'c5'])
def test_sharded_engine_gather_equals_single_gpu(tmp_path, config):
    world = 2
    out = tmp_path / 'result.txt'
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(ROOT, 'tests', 'dist_worker.py'), str(out), config]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    text = out.read_text()
    assert text.startswith('OK'), text
    assert f'world={world}' in text
