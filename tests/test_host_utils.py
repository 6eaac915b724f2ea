"""Host utilities keep the reference's behaviour (reference autoencoder/tests/test_utils.py re-expressed)."""
import numpy as np
import pandas as pd
from scipy import sparse

from dae_rnn_news_recommendation_b200.autoencoder import utils
from oracle import dae_oracle as O


def test_gen_batches():  # reference test_utils.py:11-61
    n = 30
    data = np.arange(n, dtype=np.float32).reshape(-1, 1)
    corrupted = np.random.randint(0, 2, (n, 10)).astype(np.float32)
    lab = np.random.randint(0, 10, n).astype(np.float32)
    for label in [None, lab, lab.reshape(-1, 1), pd.Series(lab), pd.DataFrame(lab)]:
        for func in [lambda v: v, sparse.csr_matrix, pd.DataFrame]:
            a_in, b_in = func(data), func(corrupted)
            if isinstance(a_in, pd.DataFrame):
                a_in.index = np.random.choice(n * 2, n, replace=False)
                b_in.index = a_in.index
            if isinstance(label, (pd.DataFrame, pd.Series)):
                label.index = np.random.choice(n * 2, n, replace=False)
            for bs in [4, 0.3]:
                seen = np.zeros(n)
                for res in utils.gen_batches(a_in, b_in, batch_size=bs, data_label=label):
                    a, b = res[0], res[1]
                    if sparse.issparse(a):
                        a, b = a.toarray(), b.toarray()
                    a, b = np.asarray(a), np.asarray(b)
                    idx = list(a[:, 0].astype(int))
                    assert (corrupted[idx, :] == b).all()
                    if label is not None:
                        got = np.asarray(res[2]).reshape(-1)
                        assert (lab[idx] == got).all()
                    seen[idx] += 1
                assert (seen == 1).all()


def test_gen_batches_triplet():  # reference test_utils.py:63-106
    n = 30
    data = {k: np.arange(n, dtype=np.float32).reshape(-1, 1) for k in ('org', 'pos', 'neg')}
    corr = {k: np.random.randint(0, 2, (n, 10)).astype(np.float32) for k in ('org', 'pos', 'neg')}
    for func in [lambda v: v, sparse.csr_matrix]:
        d = {k: func(v) for k, v in data.items()}
        c = {k: func(v) for k, v in corr.items()}
        seen = np.zeros(n)
        for a, b in utils.gen_batches_triplet(d, c, batch_size=4.0):
            if sparse.issparse(a[0]):
                a = [m.toarray() for m in a]
                b = [m.toarray() for m in b]
            assert (a[0] == a[1]).all() and (a[0] == a[2]).all()
            idx = list(a[0][:, 0].astype(int))
            for i, k in enumerate(corr):
                assert (corr[k][idx, :] == b[i]).all()
            seen[idx] += 1
        assert (seen == 1).all()


def test_masking_noise():  # reference test_utils.py:108-125
    n = 10
    X = sparse.csr_matrix(np.random.rand(n, 10000).astype(np.float32))
    for in_X in [X, X.toarray()]:
        for prob in [0., 0.3, 1.]:
            Xm = sparse.csr_matrix(utils.masking_noise(in_X, prob))
            if prob == 0.:
                assert (X != Xm).nnz == 0
            elif prob == 1.:
                assert Xm.nnz == 0
            else:
                assert abs(Xm.nnz / X.nnz - (1. - prob)) <= 1e-2
                for i in range(n):
                    assert set(Xm.indices[Xm.indptr[i]:Xm.indptr[i + 1]]) <= set(X.indices[X.indptr[i]:X.indptr[i + 1]])


def test_masking_keep_mask_is_the_same_rng_draw():
    """masking_keep_mask consumes exactly the draw masking_noise makes, so a seeded run corrupts the same entries."""
    X = sparse.random(50, 400, density=0.05, format='csr', dtype=np.float32, random_state=1)
    X.sort_indices()
    np.random.seed(7)
    a = utils.masking_noise(X, 0.3)
    after_a = np.random.rand()
    np.random.seed(7)
    keep = utils.masking_keep_mask(X, 0.3)
    after_b = np.random.rand()
    assert after_a == after_b
    b = X.copy()
    b.data = b.data * keep
    b.eliminate_zeros()
    assert (a != b).nnz == 0
    np.random.seed(7)
    c = O.masking_noise(X, 0.3)
    assert (a != c).nnz == 0


def test_decay_noise_and_sparse_feed():
    X = sparse.coo_matrix(np.random.randint(0, 5, (10, 3)).astype(np.float32))
    assert np.allclose(utils.decay_noise(X.tocsr(), 0.25).toarray(), X.toarray() * 0.75)
    ind, val, shape = utils.get_sparse_ind_val_shape(X)  # reference test_utils.py:133-140
    dense = np.zeros(shape, dtype=np.float32)
    dense[ind[:, 0], ind[:, 1]] = val
    assert (dense == X.toarray()).all()
    assert (np.diff(ind[:, 0]) >= 0).all()


def test_batch_size_resolution_and_xavier():
    assert utils._resolve_batch_size(8000, 0.1) == 800  # main_autoencoder.py:72 default
    assert utils._resolve_batch_size(30, 4.0) == 4
    W = utils.xavier_init(100, 20, const=1, rng=np.random.default_rng(0))
    b = np.sqrt(6.0 / 120)
    assert W.dtype == np.float32 and W.shape == (100, 20) and np.abs(W).max() <= b
    assert O.xavier_bounds(100, 20) == (-b, b)


def test_transform_shards_tile_the_rows():
    """DenoisingAutoencoder.shard_rows: contiguous, disjoint, covering, balanced to within one row."""
    from dae_rnn_news_recommendation_b200.autoencoder import DenoisingAutoencoder
    for n in (0, 1, 7, 8000, 100003):
        for world in (1, 2, 3, 8):
            r = [DenoisingAutoencoder.shard_rows(n, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
