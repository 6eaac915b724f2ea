"""Diagnostic: time the tcgen05 bf16x3 GEMM variants on the step's shapes and dump CTA-0 clock traces (GPU box)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dae_rnn_news_recommendation_b200 import _cabi
DEV = 'cuda:0'
st = lambda: torch.cuda.current_stream().cuda_stream


def split(x, ld):
    rows, cols = x.shape
    hi = torch.empty(rows, ld, dtype=torch.bfloat16, device=DEV); lo = torch.empty_like(hi)
    _cabi.call('dae_split_bf16', x.data_ptr(), rows, cols, x.stride(0), hi.data_ptr(), lo.data_ptr(), ld, -1, 1.0, st())
    return hi, lo


def run(variant, trace, M, N, K, A, a_mn, B, b_mn, C, ks=1):
    _cabi.call('dae_gemm_bf16x3_tune', variant, None if trace is None else trace.data_ptr(), M, N, K, 1.0, A[0].data_ptr(), A[1].data_ptr(),
               A[0].stride(0), a_mn, B[0].data_ptr(), B[1].data_ptr(), B[0].stride(0), b_mn, C.data_ptr(), C.stride(0), 0, -1, None, ks, 0, st())


def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


pad = lambda n: (n + 7) // 8 * 8
shapes = {'gram 800x800x500 K/K': (800, 800, 500, 0, 0, 1), 'Z 800x10000x500 K/K': (800, 10000, 500, 0, 0, 1),
          'dW 10000x501x800 MN/MN': (10000, 501, 800, 1, 1, 1), 'dE 800x500x10000 K/MN ks16': (800, 500, 10000, 0, 1, 16),
          'dEtri 800x500x800 K/MN': (800, 500, 800, 0, 1, 1)}
out = {}
import os
CL = int(os.environ.get('DAE_GEMM_CLUSTER', '0'))
print('cluster mode', CL)
for name, (M, N, K, a_mn, b_mn, ks) in shapes.items():
    A = torch.randn(M, K, device=DEV); B = torch.randn(N, K, device=DEV)
    Aop = split(A.t().contiguous(), pad(M)) if a_mn else split(A, pad(K))
    Bop = split(B.t().contiguous(), pad(N)) if b_mn else split(B, pad(K))
    C = torch.empty(M, N, device=DEV)
    for v in (0, 1):
        us = timeit(lambda: run(v, None, M, N, K, Aop, a_mn, Bop, b_mn, C, ks))
        fl = 2.0 * M * N * K
        out['%s v%d' % (name, v)] = {'us': us, 'TFLOPs_alg': fl / us / 1e6}
        print('%-34s variant %d: %8.1f us  %7.1f TFLOP/s (algorithmic)' % (name, v, us, fl / us / 1e6), flush=True)
    tr = torch.zeros(1000, dtype=torch.int64, device=DEV)
    run(0, tr, M, N, K, Aop, a_mn, Bop, b_mn, C, ks)
    torch.cuda.synchronize()
    t = tr.cpu().numpy()
    mma = t[:500][t[:500] > 0]; epi = t[500:][t[500:] > 0]
    base = mma[0]
    print('  mma trace (cycles since start):', (mma - base)[:24].tolist())
    print('  epi trace (wait_start, got_full, done)*:', (epi - base)[:12].tolist())
json.dump(out, open('gpurun_out/tune_gemm.json', 'w'), indent=1)
