"""Pins the oracle (CPU restatement) to golden vectors produced by executing the reference's own Python files under the
TF1 shim (oracle/gen_golden.py): forward values AND gradients of the loss ops, and whole `fit` trajectories."""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from helpers import rel_err
from oracle import dae_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_fit(path):
    z = dict(np.load(path, allow_pickle=False))
    x = sp.csr_matrix((z['x_data'], z['x_indices'], z['x_indptr']), shape=tuple(z['x_shape']))
    kw = {k[3:]: (v.item() if v.ndim == 0 else v) for k, v in z.items() if k.startswith('kw_')}
    return z, x, kw


@pytest.mark.parametrize('classes', [1, 3, 5])
def test_triplet_ops_match_reference_outputs(classes):
    z = np.load(os.path.join(GOLD, 'triplet_ops.npz'))
    k = 'c%d_' % classes
    lab = torch.from_numpy(z[k + 'label'])
    for name, fn in (('all', lambda e: O.batch_all_triplet_loss(lab, e, False)), ('allpos', lambda e: O.batch_all_triplet_loss(lab, e, True)),
                     ('hard', lambda e: O.batch_hard_triplet_loss(lab, e))):
        E = torch.from_numpy(z[k + 'encode']).clone().requires_grad_(True)
        loss, w, frac, num = fn(E)
        assert np.allclose(loss.detach().numpy(), z[k + name + '_loss'], rtol=1e-5, atol=1e-7)
        assert np.allclose(w.numpy(), z[k + name + '_weight'])
        assert np.allclose(frac.numpy(), z[k + name + '_fraction'], rtol=1e-5) and np.allclose(num.numpy(), z[k + name + '_num'])
        g = torch.autograd.grad(loss, [E], allow_unused=True)[0]
        g = np.zeros_like(z[k + 'encode']) if g is None else g.numpy()
        assert np.allclose(g, z[k + name + '_dencode'], rtol=1e-4, atol=1e-7)
    x, dec, w = (torch.from_numpy(z[k + n]) for n in ('x', 'decode', 'w'))
    for lf in ('cross_entropy', 'mean_squared', 'cosine_proximity'):
        assert np.allclose(O.weighted_loss(x, dec, lf).numpy(), z[k + lf], rtol=1e-5)
        assert np.allclose(O.weighted_loss(x, dec, lf, w).numpy(), z[k + lf + '_w'], rtol=1e-5)


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLD, 'fit_*.npz'))), ids=lambda p: os.path.basename(p)[4:-4])
def test_oracle_replays_reference_fit(path):
    """Same seeds -> same NumPy RNG stream (masking + shuffles) -> the oracle must reproduce the reference's per-step
    losses, final parameters and transform() output."""
    z, x, kw = load_fit(path)
    seed, bs, epochs = int(z['seed']), float(z['batch_size']), int(z['num_epochs'])
    model = O.OracleDAE(z['W0'], enc_act_func=kw['enc_act_func'], dec_act_func=kw['dec_act_func'], loss_func=kw['loss_func'],
                        opt=kw['opt'], learning_rate=float(kw['learning_rate']), momentum=float(kw.get('momentum', 0.5)),
                        alpha=float(kw.get('alpha', 1)), triplet_strategy=kw['triplet_strategy'])
    np.random.seed(seed)  # reference autoencoder.py:72-73
    labels = z['labels']
    cost, ae, tri, frac, num = [], [], [], [], []
    for _ in range(epochs):
        if kw['corr_type'] == 'masking':
            xc = O.masking_noise(x, float(kw['corr_frac']))
        elif kw['corr_type'] == 'decay':
            xc = O.decay_noise(x, float(kw['corr_frac']))
        else:
            xc = x
        for idx in O.gen_batch_indices(x.shape[0], bs):
            o = model.step(x[idx], xc[idx], labels[idx])
            cost.append(o['cost'])
            if kw['triplet_strategy'] != 'none':
                ae.append(o['autoencoder_loss']); tri.append(o['triplet_loss']); frac.append(o['fraction']); num.append(o['num'])
    assert rel_err(cost, z['step_cost']) < 2e-5
    if kw['triplet_strategy'] != 'none':
        assert rel_err(ae, z['step_ae']) < 2e-5 and rel_err(tri, z['step_tri']) < 2e-5
        assert np.allclose(num, z['step_num']) and np.allclose(frac, z['step_fraction'], rtol=1e-5)
    p = model.get_parameters()
    assert rel_err(p['enc_w'], z['enc_w']) < 2e-5 and rel_err(p['enc_b'], z['enc_b']) < 2e-5 and rel_err(p['dec_b'], z['dec_b']) < 2e-5
    assert rel_err(model.transform(x), z['transform']) < 2e-5
    assert rel_err(model.transform(x * (1.0 - float(kw['corr_frac']))), z['transform_decay']) < 2e-5
