"""Evaluation tail at the reference's size: cosine similarity of N embeddings + related-vs-unrelated AUROC
over the N(N-1)/2 pairs.  Per-stage CUDA-event times, the HBM-algorithmic roofline of the two kernels, and the reference's own
CPU call sequence (sklearn cosine_similarity + roc_curve/auc, via oracle/eval_oracle.py) timed beside it.  One JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from dae_rnn_news_recommendation_b200 import helpers
from dae_rnn_news_recommendation_b200._cabi import call

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
H = 500
cpu = '--no-cpu' not in sys.argv
rng = np.random.RandomState(0)
labels = rng.randint(0, 4, N)
emb = (rng.randn(4, H)[labels] * 0.15 + rng.randn(N, H)).astype(np.float32)
dev = torch.device('cuda:0')
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


t_sim, sim = timed(lambda: helpers.pairwise_similarity(emb, metric='cosine', to_host=False))
lab_dev = torch.from_numpy(labels.astype(np.int32)).to(dev)
n_rel, n_unrel = helpers._group_sizes(labels)
rel = torch.empty(n_rel, device=dev); unrel = torch.empty(n_unrel, device=dev)
cur = torch.zeros(2, dtype=torch.int64, device=dev)


def part():
    cur.zero_()
    call('dae_pair_partition', sim.data_ptr(), sim.stride(0), N, lab_dev.data_ptr(), rel.data_ptr(), unrel.data_ptr(), cur.data_ptr(), st)


t_part, _ = timed(part)
t_sort, (rs, us) = timed(lambda: (torch.sort(rel)[0], torch.sort(unrel)[0]))
acc = torch.zeros(1, dtype=torch.int64, device=dev)


def count():
    acc.zero_()
    call('dae_auroc_count', rs.data_ptr(), n_rel, us.data_ptr(), n_unrel, 1, acc.data_ptr(), st)


t_cnt, _ = timed(count)
auroc = int(acc.item()) / (2.0 * n_rel * n_unrel)
t0 = time.time()
full = helpers.visualize_pairwise_similarity(labels, sim)
t_api = (time.time() - t0) * 1e3
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
pairs = N * (N - 1) // 2
part_bytes = pairs * 8.0 + N * 4.0          # read the strict lower triangle once, write every pair once
cnt_bytes = (n_rel + n_unrel) * 4.0         # each score read once (the binary-search re-reads are the overhead)
res = {'N': N, 'pairs': pairs, 'n_related': n_rel, 'n_unrelated': n_unrel, 'auroc': auroc, 'auroc_api': full['auroc'],
       'ms': {'similarity_gemm': t_sim, 'pair_partition': t_part, 'sort_both_groups(torch.sort)': t_sort, 'auroc_count': t_cnt,
              'visualize_pairwise_similarity_api_wall': t_api},
       'pair_partition_GBs': part_bytes / t_part / 1e6, 'pair_partition_frac_of_hbm': part_bytes / t_part / 1e6 / peaks['hbm_gbs'],
       'auroc_count_GBs': cnt_bytes / t_cnt / 1e6, 'auroc_count_frac_of_hbm': cnt_bytes / t_cnt / 1e6 / peaks['hbm_gbs'],
       'similarity_tflops': 2.0 * N * N * H / t_sim / 1e9}
if cpu:
    from oracle import eval_oracle
    from sklearn.metrics import pairwise
    t0 = time.time()
    s_cpu = pairwise.cosine_similarity(emb).astype(np.float32); np.fill_diagonal(s_cpu, 0)
    t1 = time.time()
    r, u = eval_oracle.related_unrelated(labels, s_cpu)
    t2 = time.time()
    a_cpu = eval_oracle.auroc_sklearn(r, u)
    t3 = time.time()
    res['cpu_reference_calls_s'] = {'cosine_similarity': t1 - t0, 'mask_and_gather': t2 - t1, 'roc_curve_auc': t3 - t2, 'cores': os.cpu_count()}
    res['auroc_cpu'] = a_cpu
print(json.dumps(res))
