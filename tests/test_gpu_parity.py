"""GPU parity tests: the CUDA path (through the C-ABI, via TrainEngine) against the CPU oracle on identical seeded
inputs.  Tolerance: 1e-4 relative (north_star) on per-step losses, embeddings, gradients and updated parameters."""
import numpy as np
import pytest
import torch

from helpers import REL_TOL, rel_err, random_csr, mask_csr, xavier

pytestmark = pytest.mark.gpu


@pytest.fixture(params=['tc', 'ffma'], autouse=True)
def gemm_mode(request, monkeypatch):
    """Every parity test runs with the tcgen05 bf16x3 contractions (production) and with the CUDA-core fp32 kernels."""
    monkeypatch.setenv('DAE_GEMM', request.param)
    return request.param


def _engine(F, H, **kw):
    from dae_rnn_news_recommendation_b200.engine import TrainEngine
    return TrainEngine(F, H, device='cuda:0', **kw)


def _oracle(W0, **kw):
    from oracle.dae_oracle import OracleDAE
    return OracleDAE(W0, **kw)


def _run_step_pair(F, H, B, strategy, loss, enc, dec, opt, n_classes=4, kind='binary', seed=0, steps=2, alpha=1.0,
                   bh0=None, bv0=None):
    from dae_rnn_news_recommendation_b200.engine import DeviceCSR
    x = random_csr(B * steps, F, max(3, F // 40), kind=kind, seed=seed)
    xc, _ = mask_csr(x, 0.3, seed=seed + 1)
    labels = np.random.default_rng(seed + 2).integers(0, n_classes, B * steps).astype(np.float32)
    W0 = xavier(F, H, seed + 3) * 3.0
    kw = dict(enc_act_func=enc, dec_act_func=dec, loss_func=loss, opt=opt, learning_rate=0.05, momentum=0.5, alpha=alpha,
              triplet_strategy=strategy)
    eng = _engine(F, H, **kw)
    eng.set_parameters(W0, bh0, bv0)
    orc = _oracle(W0, bh0=bh0, bv0=bv0, **kw)
    csr = DeviceCSR(x, eng.device)
    eng.set_data(csr, torch.from_numpy(xc.data.astype(np.float32)).to(eng.device), torch.from_numpy(labels).to(eng.device))
    for s in range(steps):
        sl = slice(s * B, (s + 1) * B)
        o = orc.step(x[sl], xc[sl], labels[sl])
        eng.step(None, s * B, B)
        torch.cuda.synchronize()
        st = eng.read_stats()
        assert rel_err(st['cost'], o['cost']) < REL_TOL, (s, st, o['cost'])
        assert rel_err(st['ae_loss'], o['autoencoder_loss']) < REL_TOL, (s, st['ae_loss'], o['autoencoder_loss'])
        if strategy != 'none':
            assert abs(st['triplet_loss'] - float(o['triplet_loss'])) <= REL_TOL * max(abs(float(o['triplet_loss'])), 1e-3), \
                (s, st['triplet_loss'], o['triplet_loss'])
            assert st['num'] == pytest.approx(float(o['num']), rel=1e-3, abs=2.0), (st['num'], o['num'])
            assert st['fraction'] == pytest.approx(float(o['fraction']), rel=1e-3, abs=1e-5)
        gW, gbh, gbv = o['grads']
        g = eng.grad.cpu().numpy()
        assert rel_err(g[:F * H].reshape(F, H), gW) < REL_TOL, ('dW', s)
        assert rel_err(g[F * H:F * H + H], gbh) < REL_TOL, ('dbh', s)
        assert rel_err(g[F * H + H:], gbv) < REL_TOL, ('dbv', s)
        p = eng.get_parameters()
        q = orc.get_parameters()
        # Adam's first steps apply lr * g / (|g| + 3e-7): entries whose gradient is at fp32 rounding level get O(lr) updates
        # of either sign, so the *parameters* are only comparable loosely (losses and gradients stay at 1e-4).
        ptol = 5e-3 if opt == 'adam' else REL_TOL
        assert rel_err(p['enc_w'], q['enc_w']) < ptol
        assert rel_err(p['enc_b'], q['enc_b']) < ptol
        assert rel_err(p['dec_b'], q['dec_b']) < ptol
    # embeddings of the whole set after the updates
    emb = eng.encode(csr).cpu().numpy()
    assert rel_err(emb, orc.transform(x)) < (5e-3 if opt == 'adam' else REL_TOL)


@pytest.mark.parametrize('strategy', ['none', 'batch_all', 'batch_hard'])
@pytest.mark.parametrize('loss,enc,dec', [('cross_entropy', 'sigmoid', 'sigmoid'), ('mean_squared', 'tanh', 'none'),
                                          ('cosine_proximity', 'sigmoid', 'sigmoid')])
def test_step_parity(strategy, loss, enc, dec):
    _run_step_pair(F=300, H=24, B=96, strategy=strategy, loss=loss, enc=enc, dec=dec, opt='gradient_descent',
                   kind='binary' if loss == 'cross_entropy' else 'tfidf')


@pytest.mark.parametrize('opt', ['ada_grad', 'momentum', 'adam'])
def test_optimizers(opt):
    _run_step_pair(F=200, H=20, B=64, strategy='batch_all', loss='cross_entropy', enc='sigmoid', dec='sigmoid', opt=opt, steps=3)


@pytest.mark.parametrize('H', [50, 7, 500])
def test_odd_hidden_sizes(H):
    """compress_factor=200 gives H=50 (not a multiple of 4): exercises the 64-bit / scalar gather paths."""
    _run_step_pair(F=1000, H=H, B=40, strategy='batch_hard', loss='cross_entropy', enc='sigmoid', dec='sigmoid',
                   opt='gradient_descent', steps=1)


def test_nonzero_biases_and_alpha():
    rng = np.random.default_rng(5)
    _run_step_pair(F=256, H=32, B=80, strategy='batch_all', loss='cross_entropy', enc='tanh', dec='sigmoid',
                   opt='gradient_descent', alpha=10.0, bh0=rng.normal(0, 0.3, 32).astype(np.float32),
                   bv0=rng.normal(0, 0.3, 256).astype(np.float32))


@pytest.mark.parametrize('n_classes', [1, 2, 37])
def test_class_count_edge_cases(n_classes):
    """classes=1: no valid triplet -> loss 0, all weights 0 -> L_ae = 0/1e-16 = 0 (reference test parametrisation)."""
    for strategy in ('batch_all', 'batch_hard'):
        _run_step_pair(F=128, H=16, B=48, strategy=strategy, loss='cross_entropy', enc='sigmoid', dec='sigmoid',
                       opt='gradient_descent', n_classes=n_classes, steps=1)


def test_transform_matches_oracle_with_decay():
    from dae_rnn_news_recommendation_b200.engine import DeviceCSR
    F, H = 2000, 100
    x = random_csr(500, F, 60, kind='tfidf', seed=11)
    W0 = xavier(F, H, 12) * 2
    bh = np.random.default_rng(13).normal(0, 0.2, H).astype(np.float32)
    eng = _engine(F, H, enc_act_func='sigmoid')
    eng.set_parameters(W0, bh, None)
    orc = _oracle(W0, bh0=bh, enc_act_func='sigmoid')
    got = eng.encode(DeviceCSR(x, eng.device), in_scale=0.7).cpu().numpy()
    want = orc.transform(x * 0.7)
    assert rel_err(got, want) < REL_TOL
    assert got.dtype == np.float32 and got.shape == (500, H)


def test_explicit_triplet_step():
    from dae_rnn_news_recommendation_b200.engine import DeviceCSR
    import scipy.sparse as sp
    F, H, B = 300, 24, 64
    xs = [random_csr(B, F, 12, seed=s) for s in (21, 22, 23)]
    xcs = [mask_csr(m, 0.3, seed=30 + i)[0] for i, m in enumerate(xs)]
    W0 = xavier(F, H, 24) * 3
    kw = dict(enc_act_func='sigmoid', dec_act_func='sigmoid', loss_func='cross_entropy', opt='gradient_descent',
              learning_rate=0.05, alpha=2.0)
    eng = _engine(F, H, triplet_strategy='explicit', **kw)
    eng.set_parameters(W0)
    orc = _oracle(W0, triplet_strategy='none', **kw)
    stacked = sp.vstack(xs).tocsr()
    stacked_c = sp.vstack(xcs).tocsr()
    csr = DeviceCSR(stacked, eng.device)
    assert (stacked.indices == stacked_c.indices).all()
    eng.set_data(csr, torch.from_numpy(stacked_c.data.astype(np.float32)).to(eng.device), None)
    o = orc.step_explicit(xs, xcs)
    eng.step_explicit(None, 0, B, B)
    torch.cuda.synchronize()
    st = eng.read_stats()
    assert rel_err(st['cost'], o['cost']) < REL_TOL
    assert rel_err(st['ae_loss'], o['autoencoder_loss']) < REL_TOL
    assert rel_err(st['triplet_loss'], o['triplet_loss']) < REL_TOL
    g = eng.grad.cpu().numpy()
    assert rel_err(g[:F * H].reshape(F, H), o['grads'][0]) < REL_TOL
    assert rel_err(g[F * H:F * H + H], o['grads'][1]) < REL_TOL
    assert rel_err(g[F * H + H:], o['grads'][2]) < REL_TOL


def test_batch_prepare_permutation_and_weights():
    """dae_batch_prepare: label-sorted rows are a permutation of the slice; weights equal the B^3 reductions."""
    from dae_rnn_news_recommendation_b200 import _cabi
    from oracle.dae_oracle import batch_all_triplet_loss
    B, N = 200, 1000
    dev = torch.device('cuda:0')
    rng = np.random.default_rng(3)
    perm = torch.from_numpy(rng.permutation(N).astype(np.int32)).to(dev)
    labels = torch.from_numpy(rng.integers(0, 5, N).astype(np.float32)).to(dev)
    rows = torch.empty(B, dtype=torch.int32, device=dev)
    lab = torch.empty(B, device=dev)
    lo = torch.empty(B, dtype=torch.int32, device=dev)
    hi = torch.empty(B, dtype=torch.int32, device=dev)
    w = torch.empty(B, device=dev)
    stats = torch.zeros(16, dtype=torch.float64, device=dev)
    _cabi.call('dae_batch_prepare', perm.data_ptr(), 300, None, B, labels.data_ptr(), 1, rows.data_ptr(), lab.data_ptr(),
               lo.data_ptr(), hi.data_ptr(), w.data_ptr(), stats.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    r = rows.cpu().numpy()
    assert sorted(r.tolist()) == sorted(perm[300:500].cpu().numpy().tolist())
    l = lab.cpu().numpy()
    assert (np.diff(l) >= 0).all() and (l == labels.cpu().numpy()[r]).all()
    lo_, hi_ = lo.cpu().numpy(), hi.cpu().numpy()
    for i in range(B):
        assert (l[lo_[i]:hi_[i]] == l[i]).all() and (lo_[i] == 0 or l[lo_[i] - 1] != l[i]) and (hi_[i] == B or l[hi_[i]] != l[i])
    E = torch.randn(B, 4)
    _, w_ref, _, _ = batch_all_triplet_loss(torch.from_numpy(l), E)
    assert np.allclose(w.cpu().numpy(), w_ref.numpy())
    assert stats.cpu().numpy()[6] == pytest.approx(float(w_ref.sum()) / 3)


def test_graph_replay_matches_eager_steps(gemm_mode):
    """A captured CUDA graph of the step, replayed with device-side cursors, gives the same trajectory as eager launches."""
    from dae_rnn_news_recommendation_b200.engine import DeviceCSR
    from dae_rnn_news_recommendation_b200._cabi import STAT
    F, H, B, steps = 400, 32, 64, 5
    x = random_csr(B * steps, F, 12, seed=41)
    xc, _ = mask_csr(x, 0.3, seed=42)
    labels = np.random.default_rng(43).integers(0, 4, B * steps).astype(np.float32)
    W0 = xavier(F, H, 44) * 3
    res = []
    for mode in ('eager', 'graph'):
        eng = _engine(F, H, opt='adam', learning_rate=0.01, triplet_strategy='batch_all')
        eng.set_parameters(W0)
        eng.set_data(DeviceCSR(x, eng.device), torch.from_numpy(xc.data.astype(np.float32)).to(eng.device),
                     torch.from_numpy(labels).to(eng.device))
        perm = torch.from_numpy(np.random.default_rng(45).permutation(B * steps).astype(np.int32)).to(eng.device)
        log = torch.zeros(steps, 16, dtype=torch.float64, device=eng.device)
        if mode == 'eager':
            for s in range(steps):
                eng.step(perm, s * B, B, log[s])
        else:
            eng.capture_step_graph(perm, B, log)
            eng.set_step_cursor(0, 0)
            for s in range(steps):
                eng.replay_step()
        torch.cuda.synchronize()
        res.append((log.cpu().numpy().copy(), eng.get_parameters()))
        assert eng.step_count == steps
    assert rel_err(res[1][0][:, STAT['cost']], res[0][0][:, STAT['cost']]) < 1e-5
    assert rel_err(res[1][0][:, STAT['triplet_loss']], res[0][0][:, STAT['triplet_loss']]) < 1e-5
    assert rel_err(res[1][1]['enc_w'], res[0][1]['enc_w']) < 5e-3  # adam (see _run_step_pair)
    assert rel_err(res[1][1]['dec_b'], res[0][1]['dec_b']) < 5e-3


def test_corrupted_copy_with_its_own_structure(gemm_mode):
    """salt-and-pepper noise ADDS entries, so the corrupted matrix is a second CSR (encode reads it, the loss reads the clean one)."""
    from dae_rnn_news_recommendation_b200.engine import DeviceCSR
    from dae_rnn_news_recommendation_b200.autoencoder import utils
    F, H, B = 300, 24, 80
    x = random_csr(B, F, 12, seed=51)
    np.random.seed(3)
    xc = utils.salt_and_pepper_noise(x, 9)
    assert xc.nnz != x.nnz
    labels = np.random.default_rng(52).integers(0, 3, B).astype(np.float32)
    W0 = xavier(F, H, 53) * 3
    kw = dict(enc_act_func='sigmoid', dec_act_func='sigmoid', loss_func='cross_entropy', opt='gradient_descent', learning_rate=0.05,
              triplet_strategy='batch_hard')
    eng = _engine(F, H, **kw)
    eng.set_parameters(W0)
    eng.set_data(DeviceCSR(x, eng.device), None, torch.from_numpy(labels).to(eng.device), csr_corrupt=DeviceCSR(xc, eng.device))
    eng.step(None, 0, B)
    torch.cuda.synchronize()
    o = _oracle(W0, **kw).step(x, xc, labels)
    st = eng.read_stats()
    assert rel_err(st['cost'], o['cost']) < REL_TOL and rel_err(st['triplet_loss'], o['triplet_loss']) < REL_TOL
    g = eng.grad.cpu().numpy()
    assert rel_err(g[:F * H].reshape(F, H), o['grads'][0]) < REL_TOL


def test_triplet_estimator_fit(gemm_mode):
    """DenoisingAutoencoderTriplet.fit on {'org','pos','neg'}: first step == oracle step on the same rows, loss falls, transform works."""
    from dae_rnn_news_recommendation_b200.autoencoder import DenoisingAutoencoderTriplet
    from dae_rnn_news_recommendation_b200._cabi import STAT
    N, F = 120, 200
    data = {k: random_csr(N, F, 10, seed=s) for k, s in (('org', 61), ('pos', 62), ('neg', 63))}
    W0 = xavier(F, 20, 64) * 3
    m = DenoisingAutoencoderTriplet(model_name='t', main_dir='t', compress_factor=10, enc_act_func='sigmoid', dec_act_func='sigmoid',
                                    loss_func='cross_entropy', num_epochs=3, batch_size=40.0, opt='gradient_descent', learning_rate=0.05,
                                    corr_type='none', verbose=False, verbose_step=1, seed=5, alpha=2, W_init=W0, rng_mode='numpy')
    m.fit(data)
    assert len(m.train_cost_batch[0]) == 3 and np.isfinite(m.train_cost_batch[0]).all()
    emb = m.transform(data['org'])
    assert emb.shape == (N, 20) and np.isfinite(emb).all()
    # replay the first step on the oracle: same seed -> same shuffle (no corruption draws with corr_type none)
    np.random.seed(5)
    order = list(range(N)); np.random.shuffle(order)
    idx = order[:40]
    orc = _oracle(W0, enc_act_func='sigmoid', dec_act_func='sigmoid', loss_func='cross_entropy', opt='gradient_descent',
                  learning_rate=0.05, alpha=2.0, triplet_strategy='none')
    o = orc.step_explicit([data[k][idx] for k in ('org', 'pos', 'neg')], [data[k][idx] for k in ('org', 'pos', 'neg')])
    # first epoch's first step is the first row of the (last) epoch log only if num_epochs == 1 -> refit one epoch
    m1 = DenoisingAutoencoderTriplet(model_name='t1', main_dir='t1', compress_factor=10, enc_act_func='sigmoid', dec_act_func='sigmoid',
                                     loss_func='cross_entropy', num_epochs=1, batch_size=40.0, opt='gradient_descent', learning_rate=0.05,
                                     corr_type='none', verbose=False, verbose_step=1, seed=5, alpha=2, W_init=W0, rng_mode='numpy')
    m1.fit(data)
    assert rel_err(m1.train_cost_batch[0][0], o['cost']) < REL_TOL
    assert rel_err(m1.train_cost_batch[2][0], o['triplet_loss']) < REL_TOL


def test_host_feed_graph_replay_matches_eager(gemm_mode, monkeypatch):
    """TrainEngine.run_feed (the session.run(feed_dict) analogue): feeds with a common cap_nnz are replayed from one captured
    graph and give the same trajectory as eager per-feed launches."""
    from dae_rnn_news_recommendation_b200.engine import HostFeed
    F, H, B, steps = 400, 32, 64, 7
    rng = np.random.default_rng(71)
    W0 = xavier(F, H, 72) * 3
    batches = []
    for s in range(steps):
        xb = random_csr(B, F, 12, seed=80 + s)
        keep = rng.random(xb.nnz) >= 0.3
        batches.append((xb, (xb.data * keep).astype(np.float32), rng.integers(0, 4, B).astype(np.float32)))
    cap = max(b[0].nnz for b in batches)
    res = []
    for fixed in (False, True):
        eng = _engine(F, H, opt='momentum', learning_rate=0.05, triplet_strategy='batch_all')
        eng.set_parameters(W0)
        costs = [eng.run_feed(HostFeed(xb, xc, lb, cap_nnz=cap if fixed else None))['cost'] for xb, xc, lb in batches]
        res.append((costs, eng.get_parameters()['enc_w'], eng.step_count))
    assert res[0][2] == res[1][2] == steps
    assert rel_err(res[1][0], res[0][0]) < 1e-5 and rel_err(res[1][1], res[0][1]) < 1e-5
    # streamed form (the next feed's H2D copy overlaps the current step, scalars through a pinned ring): same trajectory
    eng = _engine(F, H, opt='momentum', learning_rate=0.05, triplet_strategy='batch_all')
    eng.set_parameters(W0)
    feeds = [HostFeed(xb, xc, lb, cap_nnz=cap) for xb, xc, lb in batches]
    outs = eng.run_feeds(feeds[:3]) + eng.run_feeds(feeds[3:])     # (the second call streams every feed: the layout is captured)
    assert len(outs) == steps and eng.step_count == steps
    assert rel_err([o['cost'] for o in outs], res[1][0]) < 1e-6 and rel_err(eng.get_parameters()['enc_w'], res[1][1]) < 1e-6
