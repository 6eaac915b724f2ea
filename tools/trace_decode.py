"""Diagnostic: clock64 trace of CTA 0 of the fused decode kernel inside a real training step (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from dae_rnn_news_recommendation_b200 import _cabi
from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR
w = bench.WORKLOAD
dev = torch.device('cuda:0')
x, labels = bench.make_data(8000, 1)
eng = TrainEngine(w['F'], w['H'], enc_act_func=w['enc'], dec_act_func=w['dec'], loss_func=w['loss'], opt=w['opt'], learning_rate=w['lr'],
                  alpha=w['alpha'], triplet_strategy=w['strategy'], device=dev)
eng.set_parameters(bench.xavier(w['F'], w['H'], 0))
eng.set_data(DeviceCSR(x, dev), None, torch.from_numpy(labels).to(dev))
eng.corrupt_masking(0.3, seed=1)
for s in range(3):
    eng.step(None, s * 800, 800)
tr = torch.zeros(1000, dtype=torch.int64, device=dev)
_cabi.call('dae_debug_set_trace', tr.data_ptr())
eng.step(None, 3 * 800, 800)
torch.cuda.synchronize()
_cabi.call('dae_debug_set_trace', None)
t = tr.cpu().numpy()
mma = t[:500][t[:500] > 0]; epi = t[500:][t[500:] > 0]
b = mma[0]
print('mma  :', (mma - b).tolist())
print('epi  :', (epi - b).tolist())
