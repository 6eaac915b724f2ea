"""The reference's own unit tests (autoencoder/tests/test_triplet_loss_utils.py) ported 1:1 onto the eager GPU functions:
unseeded-style random inputs, NumPy brute-force loops as the expected values, np.allclose tolerances."""
import numpy as np
import pytest
from sklearn.preprocessing import normalize

from oracle.dae_oracle import batch_all_bruteforce, batch_hard_bruteforce

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('classes', [1, 3, 5])
def test_batch_all_triplet_loss(classes):  # reference :72-138
    from dae_rnn_news_recommendation_b200.autoencoder.triplet_loss_utils import batch_all_triplet_loss
    rng = np.random.default_rng(100 + classes)
    n, h = 20, 6
    E = rng.random((n, h)).astype(np.float32)
    lab = rng.integers(0, classes, n).astype(np.float32)
    bf = batch_all_bruteforce(lab, E)
    loss, w, frac, num = batch_all_triplet_loss(False, lab, E, False)
    assert np.allclose(bf['loss'], loss, rtol=1e-4) and np.allclose(bf['weight'], w)
    assert np.allclose(bf['fraction'], frac, rtol=1e-4) and np.allclose(bf['num'], num)
    loss, w, _, _ = batch_all_triplet_loss(False, lab, E, True)
    assert np.allclose(bf['loss_pos'], loss, rtol=1e-4) and np.allclose(bf['weight_pos'], w)


@pytest.mark.parametrize('classes', [1, 3, 5])
def test_batch_hard_triplet_loss(classes):  # reference :140-203
    from dae_rnn_news_recommendation_b200.autoencoder.triplet_loss_utils import batch_hard_triplet_loss
    rng = np.random.default_rng(200 + classes)
    n, h = 20, 6
    E = rng.random((n, h)).astype(np.float32)
    lab = rng.integers(0, classes, n).astype(np.float32)
    bf = batch_hard_bruteforce(lab, E)
    loss, w, frac, num = batch_hard_triplet_loss(False, lab, E)
    assert np.allclose(bf['loss'], loss, rtol=1e-4), (bf['loss'], loss)
    assert np.allclose(bf['weight'], w) and np.allclose(bf['fraction'], frac) and np.allclose(bf['num'], num)


def test_weighted_loss():  # reference :205-234
    from dae_rnn_news_recommendation_b200.autoencoder.triplet_loss_utils import weighted_loss
    rng = np.random.default_rng(300)
    n, d = 20, 20
    x = rng.integers(0, 2, (n, d)).astype(np.float32)
    dec = rng.random((n, d)).astype(np.float32)
    w = rng.integers(0, 50, n).astype(np.float32)
    ce = -(x * np.log(dec + 1e-16) + (1. - x) * np.log(1. - dec + 1e-16)).sum(1)
    assert np.allclose(ce.mean(), weighted_loss(False, x, dec, loss_func='cross_entropy'), rtol=1e-4)
    assert np.allclose((ce * w).sum() / w.sum(), weighted_loss(False, x, dec, loss_func='cross_entropy', weight=w), rtol=1e-4)
    ms = np.square(x - dec).sum(1)
    assert np.allclose(ms.mean(), weighted_loss(False, x, dec, loss_func='mean_squared'), rtol=1e-4)
    assert np.allclose((ms * w).sum() / w.sum(), weighted_loss(False, x, dec, loss_func='mean_squared', weight=w), rtol=1e-4)
    cs = -(normalize(x, axis=1) * normalize(dec, axis=1)).sum(1)
    assert np.allclose(cs.mean(), weighted_loss(False, x, dec, loss_func='cosine_proximity'), rtol=1e-4)
    assert np.allclose((cs * w).sum() / w.sum(), weighted_loss(False, x, dec, loss_func='cosine_proximity', weight=w), rtol=1e-4)
