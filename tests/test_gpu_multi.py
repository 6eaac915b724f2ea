"""Two GPUs of one node (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`; skipped on a one-GPU box): the data-parallel
TrainEngine step under every gradient exchange -- eager NCCL between two graphs, NCCL captured in the step's graph, the in-switch
kernel dae_allreduce_multimem -- against the oracle's "P batches, mean of gradients" step (mode A): same sums, same
training trajectory, identical replicas."""
import glob
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs on one node')
def test_dp_engine_matches_oracle_mean_of_gradients(tmp_path):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', '29561', os.path.join(ROOT, 'tests', 'dp_worker.py'), str(tmp_path)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    outs = [json.load(open(f)) for f in sorted(glob.glob(str(tmp_path / 'dp_rank*.json')))]
    assert len(outs) == 2
    keep = os.environ.get('DAE_DP_REPORT')   # file name under <repo>/gpurun_out/ (the tests run from a temporary cwd)
    if keep:
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        json.dump(outs, open(os.path.join(ROOT, 'gpurun_out', os.path.basename(keep)), 'w'), indent=1)
    for o in outs:
        for mode, r in o['modes'].items():
            assert 'error' not in r, (mode, r['error'])
            assert r['raw_err'] < 1e-6, (mode, r)
            assert r['replicas_equal'], mode
            assert r['two_graphs'] == (mode == 'nccl'), mode        # the in-graph exchanges leave ONE graph per step
    r0 = outs[0]['modes']
    for mode, r in r0.items():                                       # rank 0 ran the oracle
        assert r['w_rel_err_vs_oracle'] < 1e-4, (mode, r)
        assert r['cost_rel_err_vs_oracle'] < 1e-4, (mode, r)
