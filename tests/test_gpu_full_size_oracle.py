"""Oracle parity AT the BASELINE.json sizes (B=800, F=10 000, H=500): one training step of C2 (tf-idf, batch_all), C3 (binary,
batch_hard) and C5 (explicit triplets) through the tcgen05 path against oracle/dae_oracle.py on the same seeded inputs, the same
W0 and the same corruption mask -- losses, every gradient, the updated parameters (norm-wise 1e-4) and the embeddings (norm-wise and element-wise 1e-4).

The oracle materialises the reference's B x B x B tensors (8 x 2 GB at B=800 for batch_all): it needs ~20 GB of host memory and
a few seconds per step, which the GPU boxes have."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from helpers import REL_TOL, rel_err, elem_err, xavier

pytestmark = pytest.mark.gpu
B, F, H = 800, 10000, 500
KW = dict(enc_act_func='sigmoid', dec_act_func='sigmoid', loss_func='cross_entropy', opt='gradient_descent', learning_rate=0.1,
          alpha=1.0)


def _masked(x, seed):
    keep = np.random.default_rng(seed).random(x.nnz) >= 0.3
    xc = x.copy()
    xc.data = (xc.data * keep).astype(np.float32)
    return xc


def _check(eng, st, o, x_all, orc):
    assert rel_err(st['cost'], o['cost']) < REL_TOL, (st['cost'], o['cost'])
    assert rel_err(st['ae_loss'], o['autoencoder_loss']) < REL_TOL
    assert rel_err(st['triplet_loss'], o['triplet_loss']) < REL_TOL
    gW, gbh, gbv = o['grads']
    g = eng.grad.cpu().numpy()
    assert rel_err(g[:F * H].reshape(F, H), gW) < REL_TOL
    assert rel_err(g[F * H + H:], gbv) < REL_TOL
    # dbh = sum_i dA_i - f'(bh) sum_i dE_i cancels almost completely at bh = 0: compare against the scale of its terms
    assert np.abs(g[F * H:F * H + H] - gbh).max() < REL_TOL * max(float(np.abs(gW).max()), float(np.abs(gbh).max()))
    p, q = eng.get_parameters(), orc.get_parameters()
    assert rel_err(p['enc_w'], q['enc_w']) < REL_TOL
    assert rel_err(p['dec_b'], q['dec_b']) < REL_TOL
    from dae_rnn_news_recommendation_b200.engine import DeviceCSR
    emb = eng.encode(DeviceCSR(x_all, eng.device)).cpu().numpy()
    want = orc.transform(x_all)
    assert rel_err(emb, want) < REL_TOL
    # element-wise: every embedding entry within 1e-4 of its OWN magnitude; entries below 10 % of the largest one are held to 1e-4 of
    # that floor (1e-5 of the scale): E = f(A) - f(bh) is a difference of two numbers near 0.5, so its fp32 value carries ~1e-7 of
    # ABSOLUTE rounding noise in the reference's own arithmetic as well
    assert elem_err(emb, want, floor=0.1) < REL_TOL, elem_err(emb, want, floor=0.1)


@pytest.mark.parametrize('strategy,kind', [('batch_all', 'tfidf'), ('batch_hard', 'binary')])
def test_c2_c3_step_against_oracle(strategy, kind):
    from oracle.dae_oracle import OracleDAE
    from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR
    from dae_rnn_news_recommendation_b200.synth import make_sparse, make_labels
    x = make_sparse(B, F, 100, kind, seed=21)
    labels = make_labels(B, 4, seed=21)
    xc = _masked(x, 22)
    W0 = xavier(F, H, 23)
    eng = TrainEngine(F, H, device='cuda:0', triplet_strategy=strategy, **KW)
    eng.set_parameters(W0)
    eng.set_data(DeviceCSR(x, eng.device), torch.from_numpy(xc.data).to(eng.device), torch.from_numpy(labels).to(eng.device))
    eng.step(None, 0, B)
    torch.cuda.synchronize()
    st = eng.read_stats()
    orc = OracleDAE(W0, triplet_strategy=strategy, **KW)
    o = orc.step(x, xc, labels)
    assert st['num'] == pytest.approx(float(o['num']), rel=1e-3, abs=2.0)
    assert st['fraction'] == pytest.approx(float(o['fraction']), rel=1e-3, abs=1e-5)
    _check(eng, st, o, x, orc)


def test_c5_explicit_step_against_oracle():
    from oracle.dae_oracle import OracleDAE
    from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR
    from dae_rnn_news_recommendation_b200.synth import make_sparse, perturb_rows
    org = make_sparse(B, F, 100, 'binary', seed=31)
    xs = [org, perturb_rows(org, 0.3, seed=32), make_sparse(B, F, 100, 'binary', seed=33)]
    xcs = [_masked(m, 34 + i) for i, m in enumerate(xs)]
    W0 = xavier(F, H, 37)
    eng = TrainEngine(F, H, device='cuda:0', triplet_strategy='explicit', **KW)
    eng.set_parameters(W0)
    stacked, stacked_c = sp.vstack(xs).tocsr(), sp.vstack(xcs).tocsr()
    eng.set_data(DeviceCSR(stacked, eng.device), torch.from_numpy(stacked_c.data.astype(np.float32)).to(eng.device), None)
    eng.step_explicit(None, 0, B, B)
    torch.cuda.synchronize()
    st = eng.read_stats()
    orc = OracleDAE(W0, triplet_strategy='none', **KW)
    o = orc.step_explicit(xs, xcs)
    _check(eng, st, o, org, orc)
