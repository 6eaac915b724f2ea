#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN FILES (test infrastructure).

    python oracle/gen_golden.py            # needs /root/reference; run in the builder container only

`/root/reference/autoencoder/{autoencoder,autoencoder_triplet,triplet_loss_utils,utils}.py` are imported unmodified;
the only substitution is the `tensorflow` module, which resolves to oracle/tf1_shim (TensorFlow 1.12.0 -- reference
requirements.txt:4 -- is not installable offline).  Two environment patches are applied because the reference
targets NumPy 1.15: `np.int` (removed in NumPy 1.24; autoencoder.py:187) is aliased to `int`.

Outputs (committed; the GPU box never needs /root/reference):
  tests/golden/triplet_ops.npz   forward values + d(loss)/d(encode) of batch_all / batch_hard / weighted_loss
  tests/golden/fit_<case>.npz    data, W0, seeds, per-step losses, final parameters and transform() output of
                                 DenoisingAutoencoder.fit on a seeded sparse problem, one file per configuration
"""
import os
import sys
import tempfile

import numpy as np
import scipy.sparse as sp
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('DAE_REFERENCE', '/root/reference')
OUT = os.path.join(ROOT, 'tests', 'golden')

sys.path.insert(0, os.path.join(HERE, 'tf1_shim'))
sys.path.insert(0, REF)
if not hasattr(np, 'int'):
    np.int = int  # NumPy >= 1.24 compatibility for reference autoencoder.py:187

import tensorflow as tf  # noqa: E402  (the shim)
from autoencoder import triplet_loss_utils as ref_tl  # noqa: E402  (reference code)
from autoencoder import autoencoder as ref_ae  # noqa: E402
from autoencoder import autoencoder_triplet as ref_aet  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import random_csr  # noqa: E402


def gen_triplet_ops():
    out = {}
    rng = np.random.RandomState(1234)
    for classes in (1, 3, 5):
        n, h, d = 24, 6, 20
        E = rng.rand(n, h).astype(np.float32) - 0.3
        lab = rng.randint(0, classes, n).astype(np.float32)
        x = rng.randint(0, 2, (n, d)).astype(np.float32)
        dec = (rng.rand(n, d) * 0.98 + 0.01).astype(np.float32)
        w = rng.randint(0, 50, n).astype(np.float32)
        key = 'c%d_' % classes
        out[key + 'encode'], out[key + 'label'], out[key + 'x'], out[key + 'decode'], out[key + 'w'] = E, lab, x, dec, w
        for name, fn in (('all', lambda e: ref_tl.batch_all_triplet_loss(False, lab, e, False)),
                         ('allpos', lambda e: ref_tl.batch_all_triplet_loss(False, lab, e, True)),
                         ('hard', lambda e: ref_tl.batch_hard_triplet_loss(False, lab, e))):
            tf.reset_default_graph()
            v = tf.Variable(E, name='e')
            loss, weight, frac, num = fn(v)
            with tf.Session() as s:
                s.run(tf.global_variables_initializer())
                ctx = tf._Ctx(None)
                lv = tf._eval(loss, ctx)
                g, = torch.autograd.grad(lv, [v.value], allow_unused=True)
                vals = s.run([loss, weight, frac, num])
            out[key + name + '_loss'], out[key + name + '_weight'] = np.float32(vals[0]), np.asarray(vals[1], np.float32)
            out[key + name + '_fraction'], out[key + name + '_num'] = np.float32(vals[2]), np.float32(vals[3])
            out[key + name + '_dencode'] = np.zeros_like(E) if g is None else g.numpy()
        for lf in ('cross_entropy', 'mean_squared', 'cosine_proximity'):
            with tf.Session() as s:
                out[key + lf] = np.float32(s.run(ref_tl.weighted_loss(False, x, dec, loss_func=lf)))
                out[key + lf + '_w'] = np.float32(s.run(ref_tl.weighted_loss(False, x, dec, loss_func=lf, weight=w)))
    np.savez_compressed(os.path.join(OUT, 'triplet_ops.npz'), **out)
    print('triplet_ops.npz', len(out), 'arrays')


FIT_CASES = {
    # name: ctor kwargs (reference defaults of main_autoencoder.py unless stated), data kind, n_classes
    'uci_like_none_sgd': dict(kw=dict(triplet_strategy='none', opt='gradient_descent', loss_func='cross_entropy',
                                      enc_act_func='sigmoid', dec_act_func='sigmoid', corr_type='masking', corr_frac=0.3,
                                      learning_rate=0.1), kind='binary', classes=4),
    'batch_all_ce_adagrad': dict(kw=dict(triplet_strategy='batch_all', opt='ada_grad', loss_func='cross_entropy',
                                         enc_act_func='sigmoid', dec_act_func='sigmoid', corr_type='masking', corr_frac=0.3,
                                         learning_rate=0.1, alpha=10), kind='binary', classes=4),
    'batch_hard_mse_momentum': dict(kw=dict(triplet_strategy='batch_hard', opt='momentum', loss_func='mean_squared',
                                            enc_act_func='tanh', dec_act_func='none', corr_type='masking', corr_frac=0.3,
                                            learning_rate=0.01, momentum=0.5, alpha=1), kind='tfidf', classes=5),
    'batch_all_cosine_decay': dict(kw=dict(triplet_strategy='batch_all', opt='gradient_descent', loss_func='cosine_proximity',
                                           enc_act_func='sigmoid', dec_act_func='sigmoid', corr_type='decay', corr_frac=0.3,
                                           learning_rate=0.1, alpha=1), kind='tfidf', classes=3),
}


def gen_fit(name, spec, seed=7):
    N, F, cf, bs, epochs = 240, 400, 20, 0.25, 3   # H = 20, 4 batches of 60 per epoch, 12 steps
    x = random_csr(N, F, 14, kind=spec['kind'], seed=100 + len(name))
    labels = np.random.RandomState(5).randint(0, spec['classes'], N).astype(np.float32)
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    try:
        tf.reset_default_graph()
        model = ref_ae.DenoisingAutoencoder(seed=seed, model_name=name, main_dir=name, compress_factor=cf, num_epochs=epochs,
                                            batch_size=bs, verbose=False, verbose_step=1, **spec['kw'])
        # W0: what utils.xavier_init draws from the (shim's) graph-level RNG seeded by tf.set_random_seed(seed)
        H = F // cf
        b = np.sqrt(6.0 / (F + H))
        W0 = np.random.RandomState(seed).uniform(-b, b, size=(F, H)).astype(np.float32)
        steps = []
        orig = model._run_train_step

        def spy(*a, **k):
            r = orig(*a, **k)
            steps.append([np.array(c, dtype=np.float64) for c in model.train_cost_batch] +
                         [np.array(model.fraction_triplet_batch, dtype=np.float64), np.array(model.num_triplet_batch, dtype=np.float64)])
            return r
        model._run_train_step = spy
        model.fit(x, None, labels)
        params = model.get_model_parameters()
        emb = model.transform(x)
        emb_decay = model.transform(x * (1.0 - spec['kw']['corr_frac']))
    finally:
        os.chdir(cwd)
    out = dict(x_data=x.data, x_indices=x.indices, x_indptr=x.indptr, x_shape=np.array(x.shape), labels=labels, W0=W0,
               seed=np.int64(seed), compress_factor=np.int64(cf), batch_size=np.float64(bs), num_epochs=np.int64(epochs),
               enc_w=params['enc_w'], enc_b=params['enc_b'], dec_b=params['dec_b'], transform=emb, transform_decay=emb_decay)
    for k in ('cost', 'ae', 'tri', 'fraction', 'num'):
        idx = ('cost', 'ae', 'tri', 'fraction', 'num').index(k)
        out['step_' + k] = np.concatenate([s[idx] for s in steps]) if steps and len(steps[0][idx]) else np.zeros(0)
    for k, v in spec['kw'].items():
        out['kw_' + k] = np.array(v)
    np.savez_compressed(os.path.join(OUT, 'fit_%s.npz' % name), **out)
    print('fit_%s.npz' % name, 'steps', len(out['step_cost']), 'cost', out['step_cost'][0], '->', out['step_cost'][-1])


def check_w0_is_the_shim_draw():
    tf.set_random_seed(7)
    from autoencoder import utils as ref_utils
    with tf.Session() as s:
        w = s.run(ref_utils.xavier_init(400, 20, 1))
    b = np.sqrt(6.0 / 420)
    assert np.array_equal(w, np.random.RandomState(7).uniform(-b, b, size=(400, 20)).astype(np.float32))


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    check_w0_is_the_shim_draw()
    gen_triplet_ops()
    for name, spec in FIT_CASES.items():
        gen_fit(name, spec)
