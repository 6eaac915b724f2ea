"""transform-sized K1: the hot-rows-in-shared-memory kernel (dae_encode_csr_fwd_hot, bulk-TMA staged W rows) against the row-gather
kernel and the oracle; row-range (data-parallel shard) encodes."""
import numpy as np
import pytest
import torch

from helpers import REL_TOL, rel_err, elem_err, xavier

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('F,H,act,groups', [(10000, 500, 'sigmoid', 8), (6000, 1000, 'tanh', 4), (3000, 52, 'none', 8)])
def test_hot_kernel_matches_row_kernel_and_oracle(F, H, act, groups):
    from oracle.dae_oracle import OracleDAE
    from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR
    from dae_rnn_news_recommendation_b200.synth import make_sparse
    N = 20000
    x = make_sparse(N, F, 60, 'tfidf', seed=3)
    x[7] = 0                                    # an empty row
    x.eliminate_zeros()
    W0 = xavier(F, H, 4) * 3
    bh = (np.random.default_rng(5).standard_normal(H) * 0.1).astype(np.float32)
    eng = TrainEngine(F, H, enc_act_func=act, triplet_strategy='none', device=DEV)
    eng.set_parameters(W0, bh)
    csr = DeviceCSR(x, eng.device)
    row = eng.encode(csr, in_scale=0.7).cpu().numpy()                     # row-gather kernel (the default)
    eng.HOT_MIN_ROWS, eng.HOT_GROUPS = 1, groups
    hot = eng.encode(csr, in_scale=0.7).cpu().numpy()                     # hot-rows kernel
    assert np.abs(hot - row).max() <= 1e-6 * np.abs(row).max()            # same products, same order inside a row
    sub = slice(0, 3000)
    orc = OracleDAE(W0, bh0=bh, enc_act_func=act, triplet_strategy='none')
    want = orc.transform(x[sub] * 0.7)
    assert rel_err(hot[sub], want) < REL_TOL and elem_err(hot[sub], want, floor=0.1) < REL_TOL
    assert np.all(hot[7] == 0.0)


def test_row_range_encode_is_a_slice():
    from dae_rnn_news_recommendation_b200.engine import TrainEngine, DeviceCSR
    from dae_rnn_news_recommendation_b200.autoencoder import DenoisingAutoencoder
    from dae_rnn_news_recommendation_b200.synth import make_sparse
    F, H, N = 2000, 100, 5000
    x = make_sparse(N, F, 40, 'binary', seed=1)
    eng = TrainEngine(F, H, triplet_strategy='none', device=DEV)
    eng.set_parameters(xavier(F, H, 2))
    csr = DeviceCSR(x, eng.device)
    full = eng.encode(csr).cpu().numpy()
    world = 3
    parts = [eng.encode(csr, rows=DenoisingAutoencoder.shard_rows(N, world, r)).cpu().numpy() for r in range(world)]
    # the shards tile the set, no collective needed (a short launch splits each row over 4 thread groups: same products, another
    # summation order than the long launch)
    assert np.abs(np.concatenate(parts) - full).max() <= 2e-6 * np.abs(full).max()
    assert eng.encode(csr, rows=(10, 10)).shape == (0, H)
