"""BASELINE.json configs[0] on the REAL data (UCI news, 8000 x 10000 binary CSR, nnz 1 241 293; tests/golden/uci_c1.npz):
step parity against the oracle on real batches, and the reference's end-to-end recipe -- fit, transform with decayed input,
cosine similarity, related-vs-unrelated AUROC (main_autoencoder.py:277-347) -- entirely on the GPU."""
import numpy as np
import pytest
import torch

from helpers import REL_TOL, rel_err, load_uci_c1, mask_csr, xavier

pytestmark = pytest.mark.gpu


def test_fixture_is_the_reference_configuration():
    d = load_uci_c1()
    assert d['train'].shape == (8000, 10000) and d['validate'].shape == (2000, 10000)
    assert abs(d['train'].nnz - 1241293) < 12413                         # nnz 1 241 293 (1.55 %)
    assert len(np.unique(d['train_label_category_publish_name'])) == 4


@pytest.mark.parametrize('strategy', ['none', 'batch_all'])
def test_real_batches_step_parity(strategy):
    """Two B=800 steps over real articles with real category labels (unbalanced classes, Zipf columns, rows of 5..1000+ words)."""
    from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR
    from oracle.dae_oracle import OracleDAE
    d = load_uci_c1()
    B, steps, F, H = 800, 2, 10000, 500
    x = d['train'][:B * steps]
    labels = d['train_label_category_publish_name'][:B * steps].astype(np.float32)
    xc, _ = mask_csr(x, 0.3, seed=4)
    W0 = xavier(F, H, 1)
    kw = dict(enc_act_func='sigmoid', dec_act_func='sigmoid', loss_func='cross_entropy', opt='gradient_descent', learning_rate=0.1,
              alpha=1.0, triplet_strategy=strategy)
    eng = TrainEngine(F, H, device='cuda:0', **kw)
    eng.set_parameters(W0)
    orc = OracleDAE(W0, **kw)
    eng.set_data(DeviceCSR(x, eng.device), torch.from_numpy(xc.data.astype(np.float32)).to(eng.device), torch.from_numpy(labels).to(eng.device))
    for s in range(steps):
        sl = slice(s * B, (s + 1) * B)
        o = orc.step(x[sl], xc[sl], labels[sl])
        eng.step(None, s * B, B)
        torch.cuda.synchronize()
        st = eng.read_stats()
        assert rel_err(st['cost'], o['cost']) < REL_TOL and rel_err(st['ae_loss'], o['autoencoder_loss']) < REL_TOL
        if strategy != 'none':
            assert abs(st['triplet_loss'] - float(o['triplet_loss'])) <= REL_TOL * abs(float(o['triplet_loss']))
            assert st['num'] == pytest.approx(float(o['num']), rel=1e-3)
    p, q = eng.get_parameters(), orc.get_parameters()
    assert rel_err(p['enc_w'], q['enc_w']) < REL_TOL and rel_err(p['dec_b'], q['dec_b']) < REL_TOL
    assert rel_err(eng.encode(DeviceCSR(x, eng.device)).cpu().numpy(), orc.transform(x)) < REL_TOL


def test_reference_recipe_end_to_end_quality():
    """The reference's recipe with its defaults but 10 epochs: the category AUROC of the embeddings must beat the binary-count
    cosine baseline on the training articles (the claim the reference's evaluation exists to show), and training must converge."""
    from dae_rnn_news_recommendation_b200.autoencoder import DenoisingAutoencoder, utils
    from dae_rnn_news_recommendation_b200 import helpers
    d = load_uci_c1()
    lab, lab_v = d['train_label_category_publish_name'], d['validate_label_category_publish_name']
    model = DenoisingAutoencoder(model_name='uci', main_dir='uci', compress_factor=20, enc_act_func='sigmoid', dec_act_func='sigmoid',
                                 loss_func='cross_entropy', corr_type='masking', corr_frac=0.3, opt='gradient_descent', learning_rate=0.1,
                                 num_epochs=10, batch_size=0.1, alpha=1, triplet_strategy='batch_all', seed=0, verbose=False)
    model.fit(d['train'], None, lab)
    hist = np.concatenate(model.history)
    assert hist.shape[0] == 100 and np.isfinite(hist[:, 0]).all() and hist[-10:, 0].mean() < hist[:10, 0].mean()
    res = {}
    for name, x, y in (('train', d['train'], lab), ('validate', d['validate'], lab_v)):
        emb = model.transform(utils.decay_noise(x, 0.3))
        assert emb.shape == (x.shape[0], 500) and np.isfinite(emb).all()
        res[name, 'encoded'] = helpers.visualize_pairwise_similarity(y, helpers.pairwise_similarity(emb, to_host=False))['auroc']
        res[name, 'binary'] = helpers.visualize_pairwise_similarity(y, helpers.pairwise_similarity(x, to_host=False))['auroc']
    print({'%s/%s' % k: round(v, 4) for k, v in res.items()})
    assert res['train', 'binary'] > 0.5 and res['train', 'encoded'] > 0.5 and res['validate', 'encoded'] > 0.5
