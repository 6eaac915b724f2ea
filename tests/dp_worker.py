"""Worker of tests/test_gpu_multi.py, one process per GPU (torchrun).  Compares the in-switch gradient exchange
(dae_allreduce_multimem, DAE_ALLREDUCE=multimem) with the NCCL all-reduce on the same seeded data-parallel steps."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import torch.distributed as dist


def run(mode, rank, world, dev):
    from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR
    from helpers import random_csr, mask_csr, xavier
    os.environ['DAE_ALLREDUCE'] = mode
    F, H, B, steps = 2000, 100, 200, 4
    eng = TrainEngine(F, H, triplet_strategy='batch_all', opt='ada_grad', device=dev)
    assert eng.allreduce_mode == mode
    eng.set_parameters(xavier(F, H, 0))
    x = random_csr(B * steps, F, 30, seed=10 + rank)          # every rank trains on its own rows
    xc, _ = mask_csr(x, 0.3, seed=20 + rank)
    labels = np.random.default_rng(30 + rank).integers(0, 4, B * steps).astype(np.float32)
    eng.set_data(DeviceCSR(x, dev), torch.from_numpy(xc.data.astype(np.float32)).to(dev), torch.from_numpy(labels).to(dev))
    # raw exchange on a known buffer
    eng.grad.copy_(torch.arange(eng.n_params, device=dev, dtype=torch.float32) * (rank + 1) * 1e-3)
    eng._allreduce_grad()
    torch.cuda.synchronize()
    want = torch.arange(eng.n_params, device=dev, dtype=torch.float32) * 1e-3 * sum(r + 1 for r in range(world))
    raw_err = float(((eng.grad - want).abs() / want.abs().clamp_min(1e-6)).max())
    # eager steps, then graph-replayed steps
    perm = torch.arange(B * steps, device=dev, dtype=torch.int32)
    log = torch.zeros(steps, 16, dtype=torch.float64, device=dev)
    eng.step(perm, 0, B, log[0])
    eng.capture_step_graph(perm, B, log, row_stride=B)
    eng.set_step_cursor(B, 1)
    for _ in range(steps - 1):
        eng.replay_step()
    torch.cuda.synchronize()
    p = eng.get_parameters()
    return {'raw_err': raw_err, 'cost': log[:, 0].cpu().tolist(), 'w_sum': float(np.abs(p['enc_w']).sum()), 'w': p['enc_w'],
            'two_graphs': eng._graph2 is not None}


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    a = run('nccl', rank, world, dev)
    b = run('multimem', rank, world, dev)
    # all ranks hold the same parameters after the exchange, in both modes
    for r in (a, b):
        t = torch.from_numpy(r['w']).to(dev)
        ref = t.clone()
        dist.broadcast(ref, src=0)
        r['replicas_equal'] = bool(torch.equal(t, ref))
    out = {'rank': rank, 'world': world, 'nccl': {k: v for k, v in a.items() if k != 'w'}, 'multimem': {k: v for k, v in b.items() if k != 'w'},
           'w_rel_diff': float(np.abs(a['w'] - b['w']).max() / np.abs(a['w']).max())}
    with open(os.path.join(sys.argv[1], 'dp_rank%d.json' % rank), 'w') as f:
        json.dump(out, f)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
