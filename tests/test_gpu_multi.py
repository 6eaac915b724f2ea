"""Two GPUs of one node (opt-in: `DAE_TEST_MULTIMEM=1 gpurun --gpus 2 -- pytest tests/test_gpu_multi.py -m gpu`): the in-switch gradient exchange
dae_allreduce_multimem against the NCCL all-reduce -- same sums, same training trajectory, identical replicas, and the whole
data-parallel step captured in ONE graph."""
import glob
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2 or os.environ.get('DAE_TEST_MULTIMEM') != '1',
                    reason='needs two GPUs on one node and DAE_TEST_MULTIMEM=1 (the in-switch exchange has not run on hardware yet)')
def test_multimem_exchange_matches_nccl(tmp_path):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', '29561', os.path.join(ROOT, 'tests', 'dp_worker.py'), str(tmp_path)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    outs = [json.load(open(f)) for f in sorted(glob.glob(str(tmp_path / 'dp_rank*.json')))]
    assert len(outs) == 2
    for o in outs:
        assert o['nccl']['raw_err'] < 1e-6 and o['multimem']['raw_err'] < 1e-6
        assert o['nccl']['replicas_equal'] and o['multimem']['replicas_equal']
        assert o['nccl']['two_graphs'] and not o['multimem']['two_graphs']          # the exchange kernel is inside the step's graph
        assert o['w_rel_diff'] < 1e-6                                                # a sum of two floats is order-independent
        assert all(abs(a - b) <= 1e-9 * abs(a) for a, b in zip(o['nccl']['cost'], o['multimem']['cost']))
