"""Synthetic bag-of-words generators for the BASELINE.json configs (there is no network for datasets).

Rows look like CountVectorizer / TfidfTransformer output (reference datasets/articles.py:131-174): ~Poisson(mean_nnz)
distinct words per article, word ids Zipf-distributed (frequent words shared across articles), values 1 (binary) or
row-L2-normalised positive weights (tf-idf like).  Deterministic in `seed`.
"""
import numpy as np
import scipy.sparse as sp


def make_sparse(n_rows, n_features, mean_nnz=100, kind='binary', seed=0, zipf_s=1.1, chunk=32768):
    rng = np.random.default_rng(seed)
    if zipf_s > 0:
        p = 1.0 / np.arange(1, n_features + 1, dtype=np.float64) ** zipf_s
    else:
        p = np.ones(n_features)
    cdf = np.cumsum(p / p.sum())
    cdf[-1] = 1.0
    colperm = rng.permutation(n_features).astype(np.int32)
    indptr = [np.zeros(1, dtype=np.int64)]
    idx_parts, val_parts = [], []
    base = 0
    for r0 in range(0, n_rows, chunk):
        m = min(chunk, n_rows - r0)
        k = np.clip(rng.poisson(mean_nnz, m), 1, n_features)
        kmax = int(k.max())
        D = 3 * kmax + 8  # candidates per row; the first k DISTINCT ones are kept (sampling without replacement)
        cand = colperm[np.searchsorted(cdf, rng.random((m, D)), side='left')].astype(np.int64)
        order = np.argsort(cand, axis=1, kind='stable')
        srt = np.take_along_axis(cand, order, axis=1)
        first_sorted = np.ones_like(srt, dtype=bool)
        first_sorted[:, 1:] = srt[:, 1:] != srt[:, :-1]
        is_first = np.zeros_like(first_sorted)
        np.put_along_axis(is_first, order, first_sorted, axis=1)
        rank = np.cumsum(is_first, axis=1)
        sel = is_first & (rank <= k[:, None])
        cols = np.where(sel, cand, n_features)  # sentinel: not selected
        cols.sort(axis=1)
        keep = cols < n_features
        cnt = keep.sum(1)
        idx = cols[keep].astype(np.int32)
        if kind == 'binary':
            val = np.ones(idx.shape[0], dtype=np.float32)
        else:
            val = (1.0 - rng.random(idx.shape[0])).astype(np.float32)  # (0,1]
            rows = np.repeat(np.arange(m), cnt)
            nrm = np.sqrt(np.bincount(rows, weights=val.astype(np.float64) ** 2, minlength=m))
            val = (val / nrm[rows]).astype(np.float32)
        ip = base + np.cumsum(cnt)
        base = int(ip[-1])
        indptr.append(ip.astype(np.int64))
        idx_parts.append(idx)
        val_parts.append(val)
    m = sp.csr_matrix((np.concatenate(val_parts), np.concatenate(idx_parts), np.concatenate(indptr)),
                      shape=(n_rows, n_features))
    m.has_sorted_indices = True
    return m


def make_labels(n_rows, n_classes=4, seed=0):
    return np.random.default_rng(seed + 7919).integers(0, n_classes, n_rows).astype(np.float32)


def perturb_rows(m, frac=0.3, seed=1, zipf_s=1.1):
    """'pos' rows for the explicit-triplet config: the anchor with ~frac of its entries resampled."""
    rng = np.random.default_rng(seed)
    coo = m.tocoo()
    keep = rng.random(coo.nnz) >= frac
    extra = make_sparse(m.shape[0], m.shape[1], mean_nnz=max(1, int(frac * m.nnz / m.shape[0])), kind='binary',
                        seed=seed + 1, zipf_s=zipf_s)
    base = sp.coo_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=m.shape).tocsr()
    out = (base + extra).tocsr()
    out.data[:] = 1.0
    out.sort_indices()
    return out.astype(np.float32)
