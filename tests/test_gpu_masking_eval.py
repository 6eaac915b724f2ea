"""Device-side masking noise (dae_mask_values, Philox mode: what bench.py and rng_mode='device' run) and the validation pass
(TrainEngine.evaluate, reference autoencoder/autoencoder.py:300-312) against the reference's own expectations / the oracle."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from helpers import REL_TOL, rel_err, random_csr, xavier

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _mask(values, frac, seed, epoch):
    from dae_rnn_news_recommendation_b200 import _cabi
    out = torch.full_like(values, float('nan'))
    _cabi.call('dae_mask_values', values.data_ptr(), None, values.numel(), float(frac), int(seed), int(epoch), out.data_ptr(),
               torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out


def test_philox_masking_noise_distribution():
    """Port of the reference's test_masking_noise (autoencoder/tests/test_utils.py:108-125) to the Philox mode of dae_mask_values:
    v = 0 is the identity, v = 1 zeroes everything, v = 0.3 keeps ~70 % of the stored entries and creates no new ones; plus what a
    counter-based generator must give: the same (seed, epoch) reproduces the mask, another epoch or seed draws a different one,
    and the kept fraction is uniform over the array (no position-dependent bias)."""
    num_data, F = 10, 10000
    X = sp.csr_matrix(np.random.default_rng(0).random((num_data, F)).astype(np.float32))
    v = torch.from_numpy(X.data).to(DEV)
    assert torch.equal(_mask(v, 0.0, 1, 0), v)
    assert int((_mask(v, 1.0, 1, 0) != 0).sum()) == 0
    m = _mask(v, 0.3, 1, 0)
    kept = m != 0
    assert abs(float(kept.float().mean()) - 0.7) <= 1e-2
    assert torch.equal(m[kept], v[kept])                      # survivors are unchanged: "no new data is created"
    assert torch.equal(_mask(v, 0.3, 1, 0), m)                # deterministic in (seed, epoch)
    for other in (_mask(v, 0.3, 1, 1), _mask(v, 0.3, 2, 0)):
        agree = float(((other != 0) == kept).float().mean())  # independent masks agree on 0.7^2 + 0.3^2 = 58 % of the entries
        assert abs(agree - 0.58) < 0.02
    parts = kept.float().view(10, -1).mean(1)
    assert float((parts - 0.7).abs().max()) < 0.02
    # an odd length exercises the tail of the 4-wide Philox draw
    w = v[:1003].clone()
    mw = _mask(w, 0.5, 3, 0)
    assert torch.isfinite(mw).all() and abs(float((mw != 0).float().mean()) - 0.5) < 0.06


@pytest.mark.parametrize('strategy', ['none', 'batch_all', 'batch_hard'])
def test_evaluate_matches_oracle_forward(strategy):
    """Validation pass: the whole validation set as ONE batch, x_corr = x, forward only (autoencoder.py:300-309): cost,
    autoencoder and triplet losses equal the oracle's forward; parameters, optimizer state and the training data are untouched."""
    from oracle.dae_oracle import OracleDAE
    from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR
    F, H, Ntr, Nval = 400, 40, 256, 150
    kw = dict(enc_act_func='sigmoid', dec_act_func='sigmoid', loss_func='cross_entropy', opt='ada_grad', learning_rate=0.05,
              alpha=0.7, triplet_strategy=strategy)
    xtr, xval = random_csr(Ntr, F, 12, seed=1), random_csr(Nval, F, 12, seed=2)
    ltr = np.random.default_rng(3).integers(0, 3, Ntr).astype(np.float32)
    lval = np.random.default_rng(4).integers(0, 3, Nval).astype(np.float32)
    W0 = xavier(F, H, 5) * 3
    eng = TrainEngine(F, H, device=DEV, **kw)
    eng.set_parameters(W0)
    tr = DeviceCSR(xtr, eng.device)
    eng.set_data(tr, None, torch.from_numpy(ltr).to(eng.device))
    eng.corrupt_masking(0.3, seed=9)
    eng.step(None, 0, 128)
    torch.cuda.synchronize()
    theta, slot, vc = eng.theta.clone(), eng.slot1.clone(), eng.values_c.clone()
    res = eng.evaluate(DeviceCSR(xval, eng.device), torch.from_numpy(lval).to(eng.device))
    orc = OracleDAE(eng.W.cpu().numpy(), bh0=eng.bh.cpu().numpy(), bv0=eng.bv.cpu().numpy(), **kw)
    o = orc.forward(xval, xval, lval)
    assert rel_err(res['cost'], float(o['cost'])) < REL_TOL
    assert rel_err(res['ae_loss'], float(o['autoencoder_loss'])) < REL_TOL
    if strategy != 'none':
        assert abs(res['triplet_loss'] - float(o['triplet_loss'])) <= REL_TOL * max(abs(float(o['triplet_loss'])), 1e-3)
    assert torch.equal(eng.theta, theta) and torch.equal(eng.slot1, slot)         # forward only
    assert eng.csr is tr and torch.equal(eng.values_c, vc)                        # training data restored
    eng.step(None, 128, 128)                                                       # and training continues
    torch.cuda.synchronize()
    assert np.isfinite(eng.read_stats()['cost'])


def test_fit_prints_validation_stats(capsys):
    """DenoisingAutoencoder.fit(validation_set=...) runs the validation pass every verbose_step epochs and stores its numbers."""
    from dae_rnn_news_recommendation_b200.autoencoder import DenoisingAutoencoder
    from oracle.dae_oracle import OracleDAE
    F, N, Nval = 300, 240, 100
    x, xv = random_csr(N, F, 10, seed=1), random_csr(Nval, F, 10, seed=2)
    lab = np.random.default_rng(3).integers(0, 3, N)
    labv = np.random.default_rng(4).integers(0, 3, Nval)
    m = DenoisingAutoencoder(model_name='val', compress_factor=10, enc_act_func='sigmoid', dec_act_func='sigmoid', loss_func='cross_entropy',
                             num_epochs=2, batch_size=80, opt='gradient_descent', learning_rate=0.1, corr_type='masking', corr_frac=0.3,
                             verbose=1, verbose_step=1, seed=3, triplet_strategy='batch_all')
    m.fit(x, xv, lab, labv)
    out = capsys.readouterr().out
    assert out.count('[Validation Stat (at this step)]') == 2
    p = m.get_model_parameters()
    o = OracleDAE(p['enc_w'], bh0=p['enc_b'], bv0=p['dec_b'], triplet_strategy='batch_all').forward(xv, xv, labv.astype(np.float32))
    assert rel_err(m.validation_cost['cost'], float(o['cost'])) < REL_TOL
