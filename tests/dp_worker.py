"""Worker of tests/test_gpu_multi.py, one process per GPU (torchrun).  Runs the same seeded data-parallel steps with every gradient
exchange of TrainEngine -- eager NCCL between two graphs ('nccl'), NCCL captured in the step's graph ('nccl_graph'), the in-switch
kernel dae_allreduce_multimem ('multimem') -- and, on rank 0, the oracle's "P batches, mean of gradients" step (SURVEY 8e mode A)."""
import json
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import torch.distributed as dist

F, H, B, STEPS = 2000, 100, 200, 4


def rank_data(rank):
    from helpers import random_csr, mask_csr
    x = random_csr(B * STEPS, F, 30, seed=10 + rank)          # every rank trains on its own rows
    xc, _ = mask_csr(x, 0.3, seed=20 + rank)
    labels = np.random.default_rng(30 + rank).integers(0, 4, B * STEPS).astype(np.float32)
    return x, xc, labels


def run(mode, rank, world, dev):
    from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR
    from helpers import xavier
    eng = TrainEngine(F, H, triplet_strategy='batch_all', opt='ada_grad', device=dev, allreduce=mode)
    assert eng.allreduce_mode == mode
    eng.set_parameters(xavier(F, H, 0))
    x, xc, labels = rank_data(rank)
    eng.set_data(DeviceCSR(x, dev), torch.from_numpy(xc.data.astype(np.float32)).to(dev), torch.from_numpy(labels).to(dev))
    # raw exchange on a known buffer
    eng.grad.copy_(torch.arange(eng.n_params, device=dev, dtype=torch.float32) * (rank + 1) * 1e-3)
    eng._allreduce_grad()
    torch.cuda.synchronize()
    want = torch.arange(eng.n_params, device=dev, dtype=torch.float32) * 1e-3 * sum(r + 1 for r in range(world))
    raw_err = float(((eng.grad - want).abs() / want.abs().clamp_min(1e-6)).max())
    # eager steps, then graph-replayed steps
    perm = torch.arange(B * STEPS, device=dev, dtype=torch.int32)
    log = torch.zeros(STEPS, 16, dtype=torch.float64, device=dev)
    eng.step(perm, 0, B, log[0])
    torch.cuda.synchronize()
    cost0 = float(log[0, 0])                     # (the capture's warm-up steps reuse the first log rows)
    eng.capture_step_graph(perm, B, log, row_stride=B)
    eng.set_step_cursor(B, 1)
    for _ in range(STEPS - 1):
        eng.replay_step()
    torch.cuda.synchronize()
    p = eng.get_parameters()
    return {'raw_err': raw_err, 'cost': [cost0] + log[1:, 0].cpu().tolist(), 'w_sum': float(np.abs(p['enc_w']).sum()), 'w': p['enc_w'],
            'two_graphs': eng._graph2 is not None}


def time_exchange(mode, dev, n_iter=40):
    """Average time of one exchange of the C2-sized flat gradient (20.04 MB), back to back on one stream."""
    from dae_rnn_news_recommendation_b200.engine import TrainEngine
    eng = TrainEngine(10000, 500, triplet_strategy='none', device=dev, allreduce=mode)
    eng.grad.fill_(1e-3)
    for _ in range(5):
        eng._allreduce_grad()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n_iter):
        eng._allreduce_grad()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n_iter * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del eng
    return float(t.item())


def oracle_mean_of_gradients(world):
    """Single-process restatement of the P-rank step: per global step the mean of the P local gradients (oracle/dae_oracle.py)."""
    from oracle.dae_oracle import OracleDAE
    from helpers import xavier
    model = OracleDAE(xavier(F, H, 0), triplet_strategy='batch_all', opt='ada_grad')
    data = [rank_data(r) for r in range(world)]
    costs = [[] for _ in range(world)]
    for s in range(STEPS):
        sl = slice(s * B, (s + 1) * B)
        grads = []
        for r, (x, xc, lab) in enumerate(data):
            out = model.forward(x[sl], xc[sl], lab[sl])
            costs[r].append(float(out['cost']))
            grads.append(model.grads(out))
        model.apply_gradients([sum(gs) / world for gs in zip(*grads)])
    return model.get_parameters()['enc_w'], costs


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    modes = os.environ.get('DAE_DP_MODES', 'nccl,nccl_graph,multimem').split(',')
    res = {}
    for mode in modes:
        try:
            r = run(mode, rank, world, dev)
            t = torch.from_numpy(r['w']).to(dev)     # all ranks hold the same parameters after the exchange
            ref = t.clone()
            dist.broadcast(ref, src=0)
            r['replicas_equal'] = bool(torch.equal(t, ref))
            r['exchange_us_20MB'] = time_exchange(mode, dev)
            res[mode] = r
        except Exception:   # noqa: BLE001 -- report which exchange failed instead of hanging the peers
            res[mode] = {'error': traceback.format_exc()[-1500:]}
            break
    out = {'rank': rank, 'world': world}
    ok = [m for m in modes if 'error' not in res.get(m, {'error': 1})]
    if rank == 0 and ok:
        w_or, costs = oracle_mean_of_gradients(world)
        out['oracle_cost_rank0'] = costs[0]
        for m in ok:
            w = res[m]['w']
            res[m]['w_rel_err_vs_oracle'] = float(np.abs(w - w_or).max() / np.abs(w_or).max())
            res[m]['cost_rel_err_vs_oracle'] = float(max(abs(a - b) / abs(b) for a, b in zip(res[m]['cost'], costs[0])))
    for m in ok:
        if m != ok[0]:
            res[m]['w_rel_diff_vs_' + ok[0]] = float(np.abs(res[m]['w'] - res[ok[0]]['w']).max() / np.abs(res[ok[0]]['w']).max())
    for m in res:
        res[m].pop('w', None)
    out['modes'] = res
    with open(os.path.join(sys.argv[1], 'dp_rank%d.json' % rank), 'w') as f:
        json.dump(out, f)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
