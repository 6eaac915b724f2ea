"""Kernel timeline of graph-replayed training steps (CUPTI through torch.profiler): which kernels of the step's parallel branches
actually overlap, and where the step's critical path runs.  Prints one JSON object: per kernel of ONE steady-state step its start
and end (us, relative to the step's first kernel) and stream, plus the busy / idle breakdown.

    python tools/step_timeline.py [--config C2] [--steps 6]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='C2')
ap.add_argument('--steps', type=int, default=6)
ap.add_argument('--rows', type=int, default=20000)
ap.add_argument('--feeds', action='store_true', help='timeline of TrainEngine.run_feeds (streamed pinned host feeds) instead of device-resident replays')
a = ap.parse_args()
w = bench.CONFIGS[a.config]
dev = torch.device('cuda:0')
B, F, H = w['B'], w['F'], w['H']
x, labels = bench.make_data(w, a.rows, 1)
eng = TrainEngine(F, H, enc_act_func=w['enc'], dec_act_func=w['dec'], loss_func=w['loss'], opt=w['opt'], learning_rate=w['lr'],
                  alpha=w['alpha'], triplet_strategy=w['strategy'], device=dev)
eng.set_parameters(bench.xavier(F, H, 0))
eng.set_data(DeviceCSR(x, dev), None, torch.from_numpy(labels).to(dev))
eng.corrupt_masking(w['corr_frac'], seed=1, epoch=0)
perm = torch.randperm(a.rows, device=dev, dtype=torch.int32)
log = torch.zeros(64, 16, dtype=torch.float64, device=dev)
eng.capture_step_graph(perm, B, log)
eng.set_step_cursor(0, 0)
for _ in range(3):
    eng.replay_step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
feeds = None
if a.feeds:
    from dae_rnn_news_recommendation_b200.engine import HostFeed
    rng = np.random.RandomState(7)
    batches = []
    for i in range(a.steps + 2):
        idx = rng.randint(0, a.rows, B)
        xb = x[idx]
        batches.append((xb, xb.data * (rng.rand(xb.nnz) >= w['corr_frac']), labels[idx]))
    cap = max(b[0].nnz for b in batches)
    feeds = [HostFeed(xb, xc, lb, cap_nnz=cap) for xb, xc, lb in batches]
    eng.run_feeds(feeds[:3])
    torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    if feeds:
        eng.run_feeds(feeds[1:])
    else:
        for _ in range(a.steps):
            eng.replay_step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
rows = [{'name': e.name[:60], 't0': e.time_range.start, 't1': e.time_range.end, 'stream': getattr(e, 'device_index', 0)} for e in ev]
# split into steps at the first kernel of a step (the batch commit / prepare kernel)
first = [i for i, r in enumerate(rows) if 'batch_commit' in r['name'] or 'batch_rows' in r['name']]
out = {'config': w['name'], 'n_events': len(rows), 'steps_found': len(first)}
if rows:
    out['step_starts_us'] = [round(rows[i]['t0'] - rows[0]['t0'], 1) for i in first]
    out['span_us'] = round(max(r['t1'] for r in rows) - rows[0]['t0'], 1)
    out['head'] = [{'name': r['name'][:40], 'start_us': round(r['t0'] - rows[0]['t0'], 1), 'dur_us': round(r['t1'] - r['t0'], 1)} for r in rows[:8]]
    cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU]
    if cpu:
        out['first_gpu_event_after_first_cpu_op_us'] = round(rows[0]['t0'] - min(e.time_range.start for e in cpu), 1)
if len(first) >= 3:
    lo, hi = first[-2], first[-1]
    step = rows[lo:hi]
    t0 = step[0]['t0']
    out['step_us'] = rows[hi]['t0'] - t0
    out['kernels'] = [{'name': r['name'], 'start_us': round(r['t0'] - t0, 1), 'end_us': round(r['t1'] - t0, 1), 'dur_us': round(r['t1'] - r['t0'], 1)} for r in step]
    # busy time (union of intervals) and the sum of durations
    iv = sorted((r['t0'], r['t1']) for r in step)
    busy, cur0, cur1 = 0.0, iv[0][0], iv[0][1]
    for s0, s1 in iv[1:]:
        if s0 > cur1:
            busy += cur1 - cur0; cur0, cur1 = s0, s1
        else:
            cur1 = max(cur1, s1)
    busy += cur1 - cur0
    out['busy_us'] = busy
    out['sum_of_kernel_us'] = float(sum(r['t1'] - r['t0'] for r in step))
print(json.dumps(out))
