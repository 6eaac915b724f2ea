"""End-to-end drop-in check on the GPU: DenoisingAutoencoder.fit / transform with the reference's seeds must reproduce
the golden trajectories recorded from the reference's own code (tests/golden/fit_*.npz), within 1e-4 relative."""
import glob
import os

import numpy as np
import pytest

from helpers import rel_err, REL_TOL
from test_oracle_golden import load_fit, GOLD

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('gemm', ['tc', 'ffma'])
@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLD, 'fit_*.npz'))), ids=lambda p: os.path.basename(p)[4:-4])
def test_fit_reproduces_reference_trajectory(path, gemm, monkeypatch):
    monkeypatch.setenv('DAE_GEMM', gemm)
    from dae_rnn_news_recommendation_b200.autoencoder import DenoisingAutoencoder
    from dae_rnn_news_recommendation_b200._cabi import STAT
    z, x, kw = load_fit(path)
    kw = {k: (v if not isinstance(v, (np.floating, np.integer)) else v.item()) for k, v in kw.items()}
    model = DenoisingAutoencoder(seed=int(z['seed']), model_name='g', main_dir='g', compress_factor=int(z['compress_factor']),
                                 num_epochs=int(z['num_epochs']), batch_size=float(z['batch_size']), verbose=False, verbose_step=1,
                                 W_init=z['W0'], rng_mode='numpy', **kw)
    model.fit(x, None, z['labels'])
    hist = np.concatenate(model.history)
    assert rel_err(hist[:, STAT['cost']], z['step_cost']) < REL_TOL
    if kw['triplet_strategy'] != 'none':
        assert rel_err(hist[:, STAT['ae_loss']], z['step_ae']) < REL_TOL
        assert rel_err(hist[:, STAT['triplet_loss']], z['step_tri']) < REL_TOL
        assert np.allclose(hist[:, STAT['num']], z['step_num'], rtol=2e-3, atol=2.0)
        assert np.allclose(hist[:, STAT['fraction']], z['step_fraction'], rtol=2e-3, atol=1e-5)
    p = model.get_model_parameters()
    assert rel_err(p['enc_w'], z['enc_w']) < REL_TOL and rel_err(p['enc_b'], z['enc_b']) < REL_TOL
    assert rel_err(p['dec_b'], z['dec_b']) < REL_TOL
    emb = model.transform(x)
    assert emb.dtype == np.float32 and rel_err(emb, z['transform']) < REL_TOL
    from dae_rnn_news_recommendation_b200.autoencoder import utils
    assert rel_err(model.transform(utils.decay_noise(x, kw['corr_frac'])), z['transform_decay']) < REL_TOL
    # checkpoint round trip through load_model
    m2 = DenoisingAutoencoder(model_name='g', main_dir='g', enc_act_func=kw['enc_act_func'])
    m2.load_model((x.shape[1], int(model.n_components)), model.model_path)
    assert np.array_equal(m2.transform(x), emb)
