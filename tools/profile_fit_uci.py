"""Where an epoch of `fit` on the real UCI data (C1) goes: per-kernel CUDA-event times of eager steps, the graph-replayed epoch,
and the per-epoch host phases (corruption, permutation, cursor, replays, sync).  python tools/profile_fit_uci.py [strategy]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from helpers import load_uci_c1, xavier
from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR

strategy = sys.argv[1] if len(sys.argv) > 1 else 'batch_all'
d = load_uci_c1()
x, lab = d['train'], d['train_label_category_publish_name'].astype(np.float32)
n, F = x.shape
H, B = 500, 800
dev = torch.device('cuda:0')
eng = TrainEngine(F, H, enc_act_func='sigmoid', dec_act_func='sigmoid', loss_func='cross_entropy', opt='gradient_descent',
                  learning_rate=0.1, alpha=1.0, triplet_strategy=strategy, device=dev)
eng.set_parameters(xavier(F, H, 0))
eng.set_data(DeviceCSR(x, dev), None, torch.from_numpy(lab).to(dev))
perm = torch.randperm(n, device=dev, dtype=torch.int32)
eng.corrupt_masking(0.3, seed=0, epoch=0)
for s in range(3):
    eng.step(perm, s * B, B)
tags = ['gemm_decode_fwd', 'gemm_decode_dW', 'gemm_decode_dE', 'gemm_gram', 'gemm_dE_tri', 'dae_encode_csr_fwd', 'dae_encode_csr_bwd',
        'dae_triplet_batch_all', 'dae_batch_prepare', 'dae_step_finalize', 'dae_optimizer_step']
eng.time_kernels(tags)
for s in range(10):
    eng.step(perm, s * B, B)
kt = {k: round(float(np.mean(v)), 4) for k, v in eng.kernel_times_ms().items() if v}
eng.time_kernels(None)
log = torch.zeros(10, 16, dtype=torch.float64, device=dev)
eng.capture_step_graph(perm, B, log, row_stride=B)
res = {'strategy': strategy, 'kernels_ms': kt, 'epochs': []}
for ep in range(4):
    torch.cuda.synchronize(); t = [time.time()]
    eng.corrupt_masking(0.3, seed=0, epoch=ep); torch.cuda.synchronize(); t.append(time.time())
    perm.copy_(torch.randperm(n, device=dev, dtype=torch.int32)); torch.cuda.synchronize(); t.append(time.time())
    eng.set_step_cursor(0, 0); torch.cuda.synchronize(); t.append(time.time())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        eng.replay_step()
    b.record(); t.append(time.time())
    torch.cuda.synchronize(); t.append(time.time())
    res['epochs'].append({'corrupt_ms': (t[1] - t[0]) * 1e3, 'randperm_ms': (t[2] - t[1]) * 1e3, 'cursor_ms': (t[3] - t[2]) * 1e3,
                          'replay_issue_ms': (t[4] - t[3]) * 1e3, 'drain_ms': (t[5] - t[4]) * 1e3, 'replays_gpu_ms': a.elapsed_time(b)})
print(json.dumps(res))
