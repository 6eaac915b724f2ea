"""Parity at the BASELINE.json sizes, where the CPU oracle's B^3 tensors are too slow/large for a unit test:
size-independent properties and cross-checks between the two independent GPU implementations of the contractions."""
import numpy as np
import pytest
import torch

from helpers import rel_err, REL_TOL, xavier

pytestmark = pytest.mark.gpu


def _setup(F, H, B, nnz, strategy, gemm, kind='tfidf', n_classes=4, seed=0, **kw):
    from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR
    from dae_rnn_news_recommendation_b200.synth import make_sparse, make_labels
    x = make_sparse(B, F, nnz, kind, seed=seed)
    labels = make_labels(B, n_classes, seed=seed)
    eng = TrainEngine(F, H, triplet_strategy=strategy, gemm=gemm, device='cuda:0', **kw)
    eng.set_parameters(xavier(F, H, seed + 1))
    csr = DeviceCSR(x, eng.device)
    eng.set_data(csr, None, torch.from_numpy(labels).to(eng.device))
    eng.corrupt_masking(0.3, seed=5, epoch=0)
    return eng, x, labels


@pytest.mark.parametrize('strategy', ['batch_all', 'batch_hard'])
def test_c2_full_size_tc_vs_cuda_core_paths(strategy):
    """C2 / C3 shapes (B=800, F=10000, H=500): tcgen05 bf16x3 path == fp32 CUDA-core path on losses and every gradient."""
    out = {}
    for gemm in ('tc', 'ffma'):
        eng, _, _ = _setup(10000, 500, 800, 100, strategy, gemm)
        eng.step(None, 0, 800)
        torch.cuda.synchronize()
        out[gemm] = (eng.read_stats(), eng.grad.cpu().numpy().copy())
    a, b = out['tc'], out['ffma']
    for k in ('cost', 'ae_loss', 'triplet_loss'):
        assert rel_err(a[0][k], b[0][k]) < REL_TOL, k
    assert a[0]['num'] == pytest.approx(b[0]['num'], rel=1e-4)
    F, H = 10000, 500
    assert rel_err(a[1][:F * H], b[1][:F * H]) < REL_TOL
    assert rel_err(a[1][F * H:F * H + H], b[1][F * H:F * H + H]) < REL_TOL
    assert rel_err(a[1][F * H + H:], b[1][F * H + H:]) < REL_TOL


def test_c2_full_size_properties():
    """batch_all at B=800: N_valid and the data weights equal the closed forms, G = dL/dS has zero row sums over
    (positives + negatives) weighted consistently, the loss falls over a few SGD steps, no NaN anywhere."""
    eng, x, labels = _setup(10000, 500, 800 * 6, 100, 'batch_all', 'tc', opt='gradient_descent', learning_rate=0.1)
    costs = []
    for s in range(6):
        eng.step(None, s * 800, 800)
        torch.cuda.synchronize()
        st = eng.read_stats()
        costs.append(st['cost'])
        lab = eng.labels_b.cpu().numpy()
        _, cnt = np.unique(lab, return_counts=True)
        assert st['n_valid'] == float(sum(c * (c - 1) * (800 - c) for c in cnt))
        assert st['sum_w'] == 3.0 * st['n_valid']
        assert 0.0 <= st['fraction'] <= 1.0 and st['num'] <= st['n_valid']
        # every valid triplet adds +sigma to G[i,k] and -sigma to G[i,j]: each row of G sums to zero
        G = eng.G[:800, :800]
        assert float(G.sum(1).abs().max()) < 1e-6
    assert np.isfinite(costs).all() and costs[-1] < costs[0]
    assert torch.isfinite(eng.theta).all()


def test_c4_shapes_against_oracle():
    """C4-like shapes (F=50000, H=1000: W = 200 MB > L2, two float4 column slices per thread in K1, K=1000 GEMMs) at a batch
    the CPU oracle can still do."""
    from oracle.dae_oracle import OracleDAE
    from dae_rnn_news_recommendation_b200.synth import make_sparse, make_labels
    from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR
    F, H, B = 50000, 1000, 192
    x = make_sparse(B, F, 100, 'tfidf', seed=3)
    labels = make_labels(B, 4, seed=3)
    keep = np.random.default_rng(4).random(x.nnz) >= 0.3
    xc = x.copy(); xc.data = (xc.data * keep).astype(np.float32)
    W0 = xavier(F, H, 5)
    kw = dict(enc_act_func='sigmoid', dec_act_func='sigmoid', loss_func='cross_entropy', opt='gradient_descent', learning_rate=0.1,
              alpha=1.0, triplet_strategy='batch_all')
    eng = TrainEngine(F, H, device='cuda:0', **kw)
    eng.set_parameters(W0)
    eng.set_data(DeviceCSR(x, eng.device), torch.from_numpy(xc.data).to(eng.device), torch.from_numpy(labels).to(eng.device))
    eng.step(None, 0, B)
    torch.cuda.synchronize()
    st = eng.read_stats()
    o = OracleDAE(W0, **kw).step(x, xc, labels)
    assert rel_err(st['cost'], o['cost']) < REL_TOL and rel_err(st['triplet_loss'], o['triplet_loss']) < REL_TOL
    g = eng.grad.cpu().numpy()
    assert rel_err(g[:F * H].reshape(F, H), o['grads'][0]) < REL_TOL
    assert rel_err(g[F * H + H:], o['grads'][2]) < REL_TOL
    # dbh = sum_i dA_i - f'(bh) sum_i dE_i cancels almost completely at bh = 0 with small Xavier weights (f'(A) ~ f'(0)): both
    # sides carry fp32 cancellation noise, so compare against the scale of the terms, not of the tiny difference
    scale = float(np.abs(o['grads'][0]).max())
    assert np.abs(g[F * H:F * H + H] - o['grads'][1]).max() < 1e-4 * scale


def test_c5_explicit_triplets_tc_vs_cuda_core_paths():
    """C5 shapes: 800 (anchor, pos, neg) triples of 10000-dim binary rows through DenoisingAutoencoderTriplet's step."""
    import scipy.sparse as sp
    from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR
    from dae_rnn_news_recommendation_b200.synth import make_sparse, perturb_rows
    B, F, H = 800, 10000, 500
    org = make_sparse(B, F, 100, 'binary', seed=7)
    stacked = sp.vstack([org, perturb_rows(org, 0.3, seed=8), make_sparse(B, F, 100, 'binary', seed=9)]).tocsr()
    res = {}
    for gemm in ('tc', 'ffma'):
        eng = TrainEngine(F, H, triplet_strategy='explicit', gemm=gemm, device='cuda:0', alpha=1.0)
        eng.set_parameters(xavier(F, H, 10))
        eng.set_data(DeviceCSR(stacked, eng.device), None, None)
        eng.corrupt_masking(0.3, seed=11)
        eng.step_explicit(None, 0, B, B)
        torch.cuda.synchronize()
        res[gemm] = (eng.read_stats(), eng.grad.cpu().numpy().copy())
    assert rel_err(res['tc'][0]['cost'], res['ffma'][0]['cost']) < REL_TOL
    assert rel_err(res['tc'][0]['triplet_loss'], res['ffma'][0]['triplet_loss']) < REL_TOL
    assert rel_err(res['tc'][1], res['ffma'][1]) < REL_TOL
