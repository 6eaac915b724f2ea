"""The reference's own tests (autoencoder/tests/test_triplet_loss_utils.py) re-expressed against the oracle: the in-test
NumPy brute-force loops are an independent statement of what batch_all / batch_hard / weighted_loss must return."""
import numpy as np
import pytest
import torch
from sklearn.preprocessing import normalize

from oracle import dae_oracle as O


@pytest.mark.parametrize('classes', [1, 3, 5])
def test_get_all_triplet_mask(classes):  # reference :11-30
    rng = np.random.default_rng(classes)
    n = 5
    lab = rng.integers(0, classes, n).astype(np.float32)
    want = np.zeros((n, n, n), bool)
    for i in range(n):
        for j in range(n):
            for k in range(n):
                if i == j or j == k or i == k:
                    continue
                if lab[i] == lab[j] and lab[i] != lab[k]:
                    want[i, j, k] = True
    assert (O.triplet_mask(torch.from_numpy(lab)).numpy() == want).all()


@pytest.mark.parametrize('classes', [1, 3, 5])
def test_anchor_masks(classes):  # reference :32-70
    rng = np.random.default_rng(10 + classes)
    n = 5
    lab = rng.integers(0, classes, n).astype(np.float32)
    ap = np.zeros((n, n), bool)
    an = np.zeros((n, n), bool)
    for i in range(n):
        for j in range(n):
            if i != j and lab[i] == lab[j]:
                ap[i, j] = True
            if i != j and lab[i] != lab[j]:
                an[i, j] = True
    assert (O.anchor_positive_mask(torch.from_numpy(lab)).numpy() == ap).all()
    assert (O.anchor_negative_mask(torch.from_numpy(lab)).numpy() == an).all()


@pytest.mark.parametrize('classes', [1, 3, 5])
def test_batch_all_triplet_loss(classes):  # reference :72-138
    rng = np.random.default_rng(20 + classes)
    n, h = 20, 6
    E = rng.random((n, h)).astype(np.float32)
    lab = rng.integers(0, classes, n).astype(np.float32)
    bf = O.batch_all_bruteforce(lab, E)
    loss, w, frac, num = O.batch_all_triplet_loss(torch.from_numpy(lab), torch.from_numpy(E), False)
    assert np.allclose(bf['loss'], loss.numpy())
    assert np.allclose(bf['weight'], w.numpy())
    assert np.allclose(bf['fraction'], frac.numpy())
    assert np.allclose(bf['num'], num.numpy())
    loss, w, _, _ = O.batch_all_triplet_loss(torch.from_numpy(lab), torch.from_numpy(E), True)
    assert np.allclose(bf['loss_pos'], loss.numpy())
    assert np.allclose(bf['weight_pos'], w.numpy())


@pytest.mark.parametrize('classes', [1, 3, 5])
def test_batch_hard_triplet_loss(classes):  # reference :140-203
    rng = np.random.default_rng(30 + classes)
    n, h = 20, 6
    E = rng.random((n, h)).astype(np.float32)
    lab = rng.integers(0, classes, n).astype(np.float32)
    bf = O.batch_hard_bruteforce(lab, E)
    loss, w, frac, num = O.batch_hard_triplet_loss(torch.from_numpy(lab), torch.from_numpy(E))
    assert np.allclose(bf['loss'], loss.numpy()), (bf['loss'], loss)
    assert np.allclose(bf['weight'], w.numpy())
    assert np.allclose(bf['fraction'], frac.numpy())
    assert np.allclose(bf['num'], num.numpy())


def test_weighted_loss():  # reference :205-234
    rng = np.random.default_rng(40)
    n, d = 20, 20
    x = rng.integers(0, 2, (n, d)).astype(np.float32)
    dec = rng.random((n, d)).astype(np.float32)
    w = rng.integers(0, 50, n).astype(np.float32)
    xt, dt, wt = torch.from_numpy(x), torch.from_numpy(dec), torch.from_numpy(w)
    ce = -(x * np.log(dec + 1e-16) + (1. - x) * np.log(1. - dec + 1e-16)).sum(1)
    assert np.allclose(ce.mean(), O.weighted_loss(xt, dt, 'cross_entropy').numpy())
    assert np.allclose((ce * w).sum() / w.sum(), O.weighted_loss(xt, dt, 'cross_entropy', wt).numpy())
    ms = np.square(x - dec).sum(1)
    assert np.allclose(ms.mean(), O.weighted_loss(xt, dt, 'mean_squared').numpy())
    assert np.allclose((ms * w).sum() / w.sum(), O.weighted_loss(xt, dt, 'mean_squared', wt).numpy())
    cs = -(normalize(x, axis=1) * normalize(dec, axis=1)).sum(1)
    assert np.allclose(cs.mean(), O.weighted_loss(xt, dt, 'cosine_proximity').numpy())
    assert np.allclose((cs * w).sum() / w.sum(), O.weighted_loss(xt, dt, 'cosine_proximity', wt).numpy())


def test_closed_form_weights_match_b3_reductions():
    """The closed forms used by dae_batch_prepare == the three axis reductions of the B^3 mask (triplet_loss_utils.py:129)."""
    rng = np.random.default_rng(50)
    for classes in (1, 2, 4, 9):
        B = 60
        lab = rng.integers(0, classes, B).astype(np.float32)
        _, w, _, _ = O.batch_all_triplet_loss(torch.from_numpy(lab), torch.randn(B, 3))
        vals, cnt = np.unique(lab, return_counts=True)
        n_of = dict(zip(vals, cnt))
        T = sum(c * (c - 1) for c in cnt)
        w_cf = np.array([2 * (n_of[l] - 1) * (B - n_of[l]) + T - n_of[l] * (n_of[l] - 1) for l in lab], dtype=np.float64)
        assert np.allclose(w.numpy(), w_cf)
        assert float(w.sum()) / 3 == sum(c * (c - 1) * (B - c) for c in cnt)


def test_fp32_vs_fp64_oracle_agree():
    from helpers import random_csr, mask_csr, xavier, rel_err
    x = random_csr(64, 200, 10, seed=1)
    xc, _ = mask_csr(x, 0.3)
    lab = np.random.default_rng(2).integers(0, 3, 64).astype(np.float32)
    W0 = xavier(200, 16, 3)
    o32 = O.OracleDAE(W0, dtype=torch.float32).step(x, xc, lab)
    o64 = O.OracleDAE(W0, dtype=torch.float64).step(x, xc, lab)
    for k in ('cost', 'autoencoder_loss', 'triplet_loss'):
        assert rel_err(o32[k], o64[k]) < 1e-5
    for a, b in zip(o32['grads'], o64['grads']):
        assert rel_err(a, b) < 1e-4
