"""`transform`-sized K1 launch (SURVEY 8d secondary metric): E = f(X.W + bh) - f(bh) for N articles in one call.
Prints one JSON line with rows/s and the HBM-algorithmic roofline fraction (GPU box)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
w = bench.WORKLOAD
dev = torch.device('cuda:0')
x, _ = bench.make_data(N, 1)
eng = TrainEngine(w['F'], w['H'], enc_act_func='sigmoid', triplet_strategy='none', device=dev)
eng.set_parameters(bench.xavier(w['F'], w['H'], 0))
csr = DeviceCSR(x, dev)
out = torch.empty(N, w['H'], device=dev)
for _ in range(3):
    eng.encode(csr, in_scale=0.7, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    eng.encode(csr, in_scale=0.7, out=out)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
alg = x.nnz * 8.0 + (N + 1) * 8.0 + w['F'] * w['H'] * 4.0 + w['H'] * 4.0 + N * w['H'] * 4.0
gather = x.nnz * w['H'] * 4.0
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
print(json.dumps({'kernel': 'dae_encode_csr_fwd (transform)', 'rows': N, 'ms': ms, 'rows_per_s': N / ms * 1e3,
                  'algorithmic_bytes': alg, 'achieved_GBs': alg / ms / 1e6, 'hbm_peak_GBs': peaks['hbm_gbs'],
                  'frac_of_hbm': alg / ms / 1e6 / peaks['hbm_gbs'], 'l2_gather_bytes': gather, 'l2_gather_GBs': gather / ms / 1e6,
                  'bytes_per_article': alg / N}))
