"""`transform`-sized K1 launch (SURVEY 8d secondary metric): E = f(X.W + bh) - f(bh) for N articles in one call, on C2
(F = 10 000, H = 500: W = 20 MB, L2 resident) and C4 (F = 50 000, H = 1000: W = 200 MB > L2), for both kernels:
  row  = dae_encode_csr_fwd      (one CTA per article, every W row gathered through L2 -> L1)
  hot  = dae_encode_csr_fwd_hot  (persistent CTAs, the K most frequent W rows staged in shared memory by bulk TMA)
Prints one JSON object: rows/s, the HBM-algorithmic roofline fraction, and the share of stored entries the hot set serves.

    python tools/bench_transform.py [N] [reps]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR
import argparse
ap = argparse.ArgumentParser()
ap.add_argument('N', nargs='?', type=int, default=100000)
ap.add_argument('reps', nargs='?', type=int, default=10)
ap.add_argument('--configs', default='C2,C4')
ap.add_argument('--variants', default='row,hot,hot_g8,hot_100k,hot_64k,hot_40k')
ap.add_argument('--warmup', type=int, default=3)
args = ap.parse_args()
N, reps = args.N, args.reps
dev = torch.device('cuda:0')
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
res = {}
for cfg in args.configs.split(','):
    w = bench.CONFIGS[cfg]
    F, H = w['F'], w['H']
    x, _ = bench.make_data(w, N, 1)
    eng = TrainEngine(F, H, enc_act_func='sigmoid', triplet_strategy='none', device=dev)
    eng.set_parameters(bench.xavier(F, H, 0))
    csr = DeviceCSR(x, dev)
    out = torch.empty(N, H, device=dev)
    cols, slot, K = eng._hot_columns(csr)
    hot_share = float((slot[csr.indices.long()] >= 0).float().mean())
    distinct = int((torch.bincount(csr.indices, minlength=F) > 0).sum())
    alg = x.nnz * 8.0 + (N + 1) * 8.0 + distinct * H * 4.0 + H * 4.0 + N * H * 4.0
    gather = x.nnz * H * 4.0
    # row kernel, then the hot-rows kernel with (row groups per CTA, staged bytes per CTA -> CTAs per SM): 200 KB = 1 CTA/SM,
    # 100 KB = 2, 64 KB = 3, 40 KB = 4
    for kern, min_rows, groups, hot_bytes in (('row', 1 << 30, 4, 200), ('hot', 1, 4, 200), ('hot_g8', 1, 8, 200), ('hot_100k', 1, 4, 100),
                                               ('hot_64k', 1, 4, 64), ('hot_40k', 1, 4, 40)):
        if kern not in args.variants.split(','):
            continue
        eng.HOT_MIN_ROWS, eng.HOT_GROUPS, eng.HOT_BYTES = min_rows, groups, hot_bytes * 1024
        if kern != 'row':
            cols, slot, K = eng._hot_columns(csr)
            hot_share = float((slot[csr.indices.long()] >= 0).float().mean())
        for _ in range(args.warmup):
            eng.encode(csr, in_scale=0.7, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            eng.encode(csr, in_scale=0.7, out=out)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        res['%s_%s' % (cfg, kern)] = {
            'kernel': 'dae_encode_csr_fwd' + ('_hot' if kern != 'row' else ''), 'rows': N, 'F': F, 'H': H, 'ms': ms, 'rows_per_s': N / ms * 1e3,
            'algorithmic_bytes': alg, 'achieved_GBs': alg / ms / 1e6, 'hbm_peak_GBs': peaks['hbm_gbs'], 'frac_of_hbm': alg / ms / 1e6 / peaks['hbm_gbs'],
            'w_row_gather_bytes': gather, 'w_row_gather_GBs': gather / ms / 1e6, 'bytes_per_article': alg / N,
            'hot_rows_K': K if kern != 'row' else 0, 'entries_served_from_smem': hot_share if kern != 'row' else 0.0,
            'row_groups_per_cta': groups if kern != 'row' else None}
    del eng, csr, out
print(json.dumps(res))
