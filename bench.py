#!/usr/bin/env python
"""bench.py -- articles/sec of the DAE-with-triplet-loss training hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one training step (corrupt-> encode -> decode -> loss -> triplet mining -> backward -> optimizer) on one batch of
B=800 synthetic articles of BASELINE.json configs[1]: 100k articles (or as many as the run consumes), 10 000-dim sparse
TF-IDF (1 % nnz), 500 hidden units, batch_all triplet loss, sigmoid/sigmoid, cross-entropy, SGD.
N>1: one process per GPU under torchrun (weak scaling: every rank trains B rows per step on its own shard, ONE NCCL
all-reduce of the flat gradient per step).

`--impl reference` times the reference algorithm restated on PyTorch-CPU (oracle/dae_oracle.py; TensorFlow 1.12 cannot
be installed offline) on the host cores, same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name='C2: synthetic tf-idf 10000-dim 1% nnz, H=500, batch_all, B=800', F=10000, H=500, B=800,
                mean_nnz=100, kind='tfidf', n_classes=4, strategy='batch_all', loss='cross_entropy', enc='sigmoid',
                dec='sigmoid', opt='gradient_descent', lr=0.1, corr_frac=0.3, alpha=1.0)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--rows', type=int, default=0, help='synthetic articles per rank (default: what the run consumes, <= 100k)')
    ap.add_argument('--flush-l2', action='store_true', help='write a 256 MB buffer between timed steps (per-step events)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch the step eagerly instead of replaying the captured CUDA graph')
    ap.add_argument('--cpu-steps', type=int, default=3)
    return ap.parse_args()


def make_data(n_rows, seed):
    from dae_rnn_news_recommendation_b200.synth import make_sparse, make_labels
    w = WORKLOAD
    x = make_sparse(n_rows, w['F'], w['mean_nnz'], w['kind'], seed=seed)
    return x, make_labels(n_rows, w['n_classes'], seed=seed)


def xavier(F, H, seed=0):
    b = np.sqrt(6.0 / (F + H))
    return np.random.default_rng(seed).uniform(-b, b, (F, H)).astype(np.float32)


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (reference algorithm on PyTorch-CPU), used for cpu_baseline and --impl reference
# ----------------------------------------------------------------------------------------------------------------------
def cpu_steps(x, labels, n_steps, n_warm, seed=0):
    import torch
    from oracle.dae_oracle import OracleDAE, masking_noise
    w = WORKLOAD
    torch.set_num_threads(os.cpu_count())
    B = w['B']
    model = OracleDAE(xavier(w['F'], w['H'], seed), enc_act_func=w['enc'], dec_act_func=w['dec'], loss_func=w['loss'],
                      opt=w['opt'], learning_rate=w['lr'], alpha=w['alpha'], triplet_strategy=w['strategy'])
    rng = np.random.RandomState(seed)
    need = (n_steps + n_warm) * B
    assert x.shape[0] >= need, (x.shape, need)
    times = []
    for s in range(n_steps + n_warm):
        t0 = time.perf_counter()
        sl = slice(s * B, (s + 1) * B)
        xb = x[sl]
        xc = masking_noise(xb, w['corr_frac'], rng)       # host corruption + batching are inside the reference's window
        model.step(xb, xc, labels[sl])
        if s >= n_warm:
            times.append(time.perf_counter() - t0)
    return times


def cpu_info():
    model = ''
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                model = line.split(':', 1)[1].strip()
                break
    except OSError:
        pass
    return model


def run_reference(args):
    """Reference arm: rank 0 only; bounded so the run ends within a few minutes."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    w = WORKLOAD
    B = w['B']
    warm = min(args.warmup, 1)
    x, labels = make_data((warm + 1) * B, seed=0)
    t_probe = cpu_steps(x, labels, 1, warm)[0]
    budget = 150.0
    k = int(max(1, min(args.steps, budget // max(t_probe, 1e-3))))
    x, labels = make_data((k + 1) * B, seed=1)
    times = cpu_steps(x, labels, k, 1)
    t = float(np.sum(times))
    val = k * B / t
    out = {'impl': 'reference', 'metric': 'articles/sec', 'value': val, 'unit': 'articles/s', 'n_gpus': args.gpus, 'steps': k,
           'steps_requested': args.steps, 'warmup': 1, 'ms_per_step': 1e3 * t / k, 'higher_is_better': True,
           'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': {'workload': w['name'], 'global_batch': B, 'note': 'reference algorithm restated on PyTorch-CPU '
                      '(TF 1.12 unavailable offline); steps capped to fit ~150 s'},
           'cpu_baseline': {'value': val, 'unit': 'articles/s', 'cores': os.cpu_count(), 'kind': 'port',
                            'sample': '%d steps of B=%d (%s)' % (k, B, cpu_info())},
           'e2e': {'value': val, 'unit': 'articles/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------------------------------------------------
class Clocks:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.lines = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '20'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        sel = [l for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l for (_, l) in self.lines]
        for l in sel:
            f = [v.strip() for v in l.split(',')]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': float(np.max(mx)) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# roofline bookkeeping: algorithmic work per launch of each kernel tag
# ----------------------------------------------------------------------------------------------------------------------
def kernel_work(tag, w, nnz_c_batch, nnz_batch):
    B, F, H = w['B'], w['F'], w['H']
    if tag in ('gemm_decode_fwd', 'gemm_decode_dW', 'gemm_decode_dE'):
        return 'tensor', 2.0 * B * F * H
    if tag == 'gemm_gram':
        return 'tensor', 2.0 * B * B * H
    if tag == 'gemm_dE_tri':
        return 'tensor', 2.0 * B * B * H
    if tag == 'dae_encode_csr_fwd':
        return 'hbm', nnz_c_batch * 8.0 + (B + 1) * 8.0 + F * H * 4.0 + H * 4.0 + B * H * 4.0
    if tag == 'dae_encode_csr_bwd':
        return 'hbm', nnz_c_batch * (8.0 + H * 4.0 * 2.0)
    if tag == 'dae_decode_loss_bwd':
        return 'hbm', 2.0 * B * F * 4.0 + nnz_batch * 8.0
    if tag == 'dae_colsum':
        return 'hbm', B * F * 4.0
    if tag == 'dae_optimizer_step':
        return 'hbm', 3.0 * (F * H + F + H) * 4.0
    return None, 0.0


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR, HostFeed
    w = WORKLOAD
    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B, F, H, K, W = w['B'], w['F'], w['H'], args.steps, max(args.warmup, 3)
    n_rows = args.rows or min(100000, (K + W + 8) * B)
    n_rows = max(n_rows, (K + W) * B if (K + W) * B <= 100000 else 100000)
    x, labels = make_data(n_rows, seed=1000 + rank)

    eng = TrainEngine(F, H, enc_act_func=w['enc'], dec_act_func=w['dec'], loss_func=w['loss'], opt=w['opt'],
                      learning_rate=w['lr'], alpha=w['alpha'], triplet_strategy=w['strategy'], device=dev)
    eng.set_parameters(xavier(F, H, 0))
    csr = DeviceCSR(x, dev)
    eng.set_data(csr, None, torch.from_numpy(labels).to(dev))
    steps_per_epoch = n_rows // B

    perm_buf = torch.zeros(n_rows, dtype=torch.int32, device=dev)
    use_graph = not args.no_graph

    def epoch_start(epoch):
        eng.corrupt_masking(w['corr_frac'], seed=1234 + rank, epoch=epoch)           # utils.masking_noise, on device
        perm_buf.copy_(torch.randperm(n_rows, device=dev, dtype=torch.int32))         # utils.gen_batches shuffle

    state = {'epoch': -1}

    def run(n, first_step, log=None, flush=None, evs=None, graph=False):
        """n steps starting at global step `first_step` (epoch boundaries re-corrupt + re-shuffle inside the window)."""
        for i in range(n):
            ep, in_epoch = divmod(first_step + i, steps_per_epoch)
            if ep != state['epoch']:
                epoch_start(ep)
                state['epoch'] = ep
            if graph and (i == 0 or in_epoch == 0):
                eng.set_step_cursor(in_epoch * B, i)
            if flush is not None:
                flush.add_(1.0)
            if evs is not None:
                evs[i][0].record()
            if graph:
                eng.replay_step()
            else:
                eng.step(perm_buf, in_epoch * B, B, None if log is None else log[i])
            if evs is not None:
                evs[i][1].record()

    run(W, 0)  # warm-up (also allocates the workspaces)
    torch.cuda.synchronize()

    # -- per-kernel profile pass (3 steps, every kernel bracketed) to find the dominant kernel
    tags = ['gemm_decode_fwd', 'gemm_decode_dW', 'gemm_decode_dE', 'gemm_gram', 'gemm_dE_tri', 'dae_encode_csr_fwd',
            'dae_encode_csr_bwd', 'dae_decode_loss_bwd', 'dae_colsum', 'dae_triplet_batch_all', 'dae_triplet_batch_hard',
            'dae_batch_prepare', 'dae_step_finalize', 'dae_optimizer_step']
    fork = eng.fork_branches
    eng.fork_branches = False   # per-kernel times are taken with the step serialised on one stream (no overlap with the mining branch)
    eng.time_kernels(tags)
    run(3, W)
    prof = {k: float(np.sum(v)) / 3.0 for k, v in eng.kernel_times_ms().items() if v}
    eng.time_kernels(None)
    eng.fork_branches = fork
    nnz_batch = x.nnz / n_rows * B
    nnz_c_batch = nnz_batch * (1.0 - w['corr_frac'])
    rooflined = {k: v for k, v in prof.items() if kernel_work(k, w, nnz_c_batch, nnz_batch)[0]}
    dominant = max(rooflined, key=rooflined.get)

    # -- timed region: EXACTLY K steps, barrier + synchronize on both sides, CUDA events, max over ranks
    log = torch.zeros(K, 16, dtype=torch.float64, device=dev)
    if use_graph:
        eng.capture_step_graph(perm_buf, B, log)   # one CUDA graph of the whole step; cursors live in device memory
    flush = torch.zeros(64 * 1024 * 1024, device=dev) if args.flush_l2 else None
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)] if args.flush_l2 else None
    eng.time_kernels(None if use_graph else [dominant])
    launches0 = eng.launches
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    run(K, W + 3, log=log, flush=flush, evs=evs, graph=use_graph)
    e1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1) if evs is None else float(sum(a.elapsed_time(b) for a, b in evs))
    gpu_launches = eng.launches - launches0
    if use_graph:  # events cannot bracket nodes inside a graph: time the dominant kernel in 3 extra eager steps right after
        eng.time_kernels([dominant])
        eng.fork_branches = False
        run(3, W + 3 + K)
        eng.fork_branches = fork
    dom_ms = eng.kernel_times_ms()[dominant]
    eng.time_kernels(None)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None
    value = K * B * world / (ms * 1e-3)
    losses = log.cpu().numpy()

    # -- e2e: per-step HOST feed (pinned) -> H2D -> step -> D2H of the step's scalars, through TrainEngine.run_feed
    from dae_rnn_news_recommendation_b200.autoencoder import utils as hostutils
    Ke = min(K, 20)
    rng = np.random.RandomState(7 + rank)
    batches = []
    for i in range(Ke + 2):
        idx = rng.randint(0, n_rows, B)
        xb = x[idx]
        keep = rng.rand(xb.nnz) >= w['corr_frac']
        batches.append((xb, xb.data * keep, labels[idx]))
    cap = max(b[0].nnz for b in batches)   # one device layout for every feed -> the step is captured once and replayed
    feeds = [HostFeed(xb, xc, lb, cap_nnz=None if args.no_graph else cap) for xb, xc, lb in batches]
    for f in feeds[:2]:
        eng.run_feed(f)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for f in feeds[2:]:
        eng.run_feed(f)
    f1.record()
    torch.cuda.synchronize()
    tme = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tme, op=dist.ReduceOp.MAX)
    e2e_val = Ke * B * world / (float(tme.item()) * 1e-3)
    h2d = int(np.mean([f.nbytes for f in feeds[2:]]))

    # -- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    bound, work = kernel_work(dominant, w, nnz_c_batch, nnz_batch)
    dur = float(np.mean(dom_ms)) * 1e-3
    if bound == 'tensor':
        achieved, unit = work / dur / 1e12, 'TFLOP/s'
        peak = peaks.get('bf16_tflops_sustained', 1400.0)
        peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else 'fallback 1.4 PFLOP/s sustained'
    else:
        achieved, unit = work / dur / 1e9, 'GB/s'
        peak = peaks.get('hbm_gbs', 6650.0)
        peak_src = 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6.65 TB/s'
    traffic, extra = None, {}
    try:  # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture (profiles/traffic.json)
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get(dominant, {})
        traffic = tj.get('traffic_bytes')
        extra = {k: v for k, v in tj.items() if k != 'traffic_bytes'}
    except (OSError, ValueError):
        pass
    roofline = {'kernel': dominant, 'bound': bound, 'achieved': achieved, 'peak': peak, 'unit': unit, 'frac': achieved / peak,
                'traffic': traffic, 'algorithmic_work': work, 'ncu': extra, 'peak_source': peak_src,
                'note': ('algorithmic FLOPs = 2*M*N*K; the kernel executes 3x that as bf16 MMAs (hi/lo split) for fp32 parity'
                         if bound == 'tensor' else ''), 'avg_launch_ms': dur * 1e3,
                'share_of_step': float(np.mean(dom_ms)) / (ms / K), 'launches_per_step': float(gpu_launches) / K}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if not args.no_cpu_baseline:
        n = max(1, args.cpu_steps)
        xs, ls = make_data((n + 1) * B, seed=1)
        ts = cpu_steps(xs, ls, n, 1)
        cpu_baseline = {'value': n * B / float(np.sum(ts)), 'unit': 'articles/s', 'cores': os.cpu_count(), 'kind': 'port',
                        'sample': '%d steps of B=%d after 1 warm-up step, reference algorithm restated on PyTorch-CPU (%s)'
                                  % (n, B, cpu_info())}

    out = {
        'metric': 'articles/sec', 'value': value, 'unit': 'articles/s', 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': w['name'], 'global_batch': B * world, 'rows_per_rank': n_rows, 'parallelism': 'dp%d' % world,
                   'l2': ('flushed between steps (256 MB write), per-step events' if args.flush_l2 else
                          'inputs larger than L2: every step reads batch rows not touched since the previous epoch; dataset CSR '
                          '+ per-step state (W, Z, grad ~ 92 MB) exceed the 126 MB L2'),
                   'loss_first_last': [float(losses[0, 0]), float(losses[-1, 0])],
                   'launch': 'cuda graph replay' if use_graph else 'eager', 'grad_exchange': eng.allreduce_mode},
        'clocks': clk,
        'e2e': {'value': e2e_val, 'unit': 'articles/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 128, 'steps': Ke,
                'api': 'TrainEngine.run_feed(HostFeed) per step: pinned host batch -> H2D -> step -> D2H scalars, synchronised'},
        'gpu_launches': gpu_launches,
        'roofline': roofline,
        'kernels_ms_per_step': prof,
        'cpu_baseline': cpu_baseline,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
