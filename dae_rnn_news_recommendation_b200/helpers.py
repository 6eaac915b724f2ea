"""GPU versions of the evaluation helpers that consume transform()'s output (reference helpers.py:11-50 and the
nearest-article lookup of main_autoencoder.py:307-318,352-359) -- SURVEY section 8f, rank 1.

    pairwise_similarity(in_df, norm='', metric='cosine', set_diagonal_zero=True) -> ndarray [N, N]     (reference signature)
    nearest_neighbors(embeddings, metric='cosine', chunk=8192) -> (index[N], score[N])                (no N x N matrix on the host)
    visualize_pairwise_similarity(labels, pairwise_similarity_metrics, ...) -> dict                   (rank 2: AUROC + box statistics)

Dense inputs (embeddings) go through the tcgen05 bf16x3 GEMM on row-normalised operands; sparse inputs (count / tf-idf
matrices) through the CSR encode kernel against the dense transpose.  No CPU path.
"""
import numpy as np
import scipy.sparse as sp
import torch

from . import _cabi
from ._cabi import call
from .engine import DeviceCSR
from .io_formats import save_file, read_file  # noqa: F401  (reference helpers.py:138-264)

_NORM = {'': 0, 'l1': 1, 'l2': 2, 'max': 3}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _normalised_operands(x_dev, norm_kind):
    n, h = x_dev.shape
    ld = (h + 7) // 8 * 8
    hi = torch.empty(n, ld, dtype=torch.bfloat16, device=x_dev.device)
    lo = torch.empty(n, ld, dtype=torch.bfloat16, device=x_dev.device)
    call('dae_rownorm_split_bf16', x_dev.data_ptr(), n, h, x_dev.stride(0), norm_kind, hi.data_ptr(), lo.data_ptr(), ld, None, 0, _stream())
    return hi, lo, ld


def _gemm_nt(a, b, n_a, n_b, k, out):
    call('dae_gemm_bf16x3', n_a, n_b, k, 1.0, a[0].data_ptr(), a[1].data_ptr(), a[0].stride(0), 0, b[0].data_ptr(), b[1].data_ptr(),
         b[0].stride(0), 0, out.data_ptr(), out.stride(0), 0, -1, None, 1, 0, _stream())


def _to_device_dense(in_df, device):
    if isinstance(in_df, list):
        in_df = np.asarray(in_df)
    if hasattr(in_df, 'values') and not isinstance(in_df, np.ndarray):
        in_df = in_df.values
    return torch.from_numpy(np.ascontiguousarray(in_df, dtype=np.float32)).to(device)


def pairwise_similarity(in_df, norm='', metric='cosine', set_diagonal_zero=True, device='cuda:0', to_host=True):
    """Reference helpers.pairwise_similarity: optional `norm` ('l1','l2','max'), then cosine similarity or the linear kernel
    of every pair of rows, diagonal zeroed.  Returns a float32 ndarray [N, N] (`to_host=False`: the device tensor, which
    visualize_pairwise_similarity accepts as is)."""
    assert metric in ['cosine', 'linear kernel']
    assert norm in _NORM
    if sp.issparse(in_df):
        out = _pairwise_sparse(in_df, norm, metric, set_diagonal_zero, device)
        return out.cpu().numpy() if to_host else out
    x = _to_device_dense(in_df, device)
    n, h = x.shape
    if norm != '':   # sklearn.preprocessing.normalize first (helpers.py:42-43) ...
        xn = torch.empty_like(x)
        call('dae_rownorm_split_bf16', x.data_ptr(), n, h, x.stride(0), _NORM[norm], None, None, 0, xn.data_ptr(), xn.stride(0), _stream())
        x = xn
    hi, lo, _ = _normalised_operands(x, 2 if metric == 'cosine' else 0)   # ... then the metric's own L2 normalisation (cosine)
    out = torch.empty(n, n, dtype=torch.float32, device=device)
    _gemm_nt((hi, lo), (hi, lo), n, n, h, out)
    if set_diagonal_zero:
        out.diagonal().zero_()
    return out.cpu().numpy() if to_host else out


def _pairwise_sparse(m, norm, metric, set_diagonal_zero, device):
    """X_hat . X_hat^T for a sparse X through the CSR encode kernel: the dense operand is X_hat^T [F x N]."""
    from sklearn.preprocessing import normalize   # host-side row scaling of the CSR values only (data prep, not the contraction)
    m = sp.csr_matrix(m, dtype=np.float32)
    if norm != '':
        m = normalize(m, norm=norm)
    if metric == 'cosine':
        m = normalize(m, norm='l2')
    n, f = m.shape
    csr = DeviceCSR(m, device)
    out = torch.empty(n, n, dtype=torch.float32, device=device)
    rows = torch.repeat_interleave(torch.arange(n, device=device), (csr.indptr[1:] - csr.indptr[:-1]))
    indptr_host = m.indptr
    blk = 2048   # the encode kernel keeps one output row of <= 4096 floats in registers: X_hat^T goes through it in column blocks
    for c0 in range(0, n, blk):
        c1 = min(n, c0 + blk)
        wp = (c1 - c0 + 3) // 4 * 4
        p0, p1 = int(indptr_host[c0]), int(indptr_host[c1])
        dense_t = torch.zeros(f, wp, dtype=torch.float32, device=device)
        dense_t[csr.indices[p0:p1].long(), rows[p0:p1] - c0] = csr.values[p0:p1]
        out_blk = torch.empty(n, wp, dtype=torch.float32, device=device)
        zero_b = torch.zeros(wp, dtype=torch.float32, device=device)
        call('dae_encode_csr_fwd', csr.indptr.data_ptr(), csr.indices.data_ptr(), csr.values.data_ptr(), None, n, f, wp, 1.0,
             dense_t.data_ptr(), zero_b.data_ptr(), _cabi.ACT['none'], out_blk.data_ptr(), wp, None, None, None, 0, _stream())
        out[:, c0:c1] = out_blk[:, :c1 - c0]
    if set_diagonal_zero:
        out.diagonal().zero_()
    return out


def nearest_neighbors(embeddings, metric='cosine', chunk=8192, device='cuda:0'):
    """For every row the most similar OTHER row and its score (np.nanargmax over the zero-diagonal similarity matrix,
    main_autoencoder.py:352-353) without materialising N x N: row chunks of the similarity are produced by the GEMM and
    reduced by dae_row_argmax on the device."""
    x = _to_device_dense(embeddings, device)
    n, h = x.shape
    hi, lo, _ = _normalised_operands(x, 2 if metric == 'cosine' else 0)
    idx = torch.empty(n, dtype=torch.int32, device=device)
    val = torch.empty(n, dtype=torch.float32, device=device)
    buf = torch.empty(min(chunk, n), n, dtype=torch.float32, device=device)
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        _gemm_nt((hi[r0:r1], lo[r0:r1]), (hi, lo), r1 - r0, n, h, buf)
        call('dae_row_argmax', buf.data_ptr(), r1 - r0, n, buf.stride(0), r0, 0, idx[r0:r1].data_ptr(), val[r0:r1].data_ptr(), _stream())
    torch.cuda.synchronize()
    return idx.cpu().numpy(), val.cpu().numpy()


def _group_sizes(labels):
    """R = #related pairs, U = #unrelated pairs of the strict lower triangle among rows with label >= 0."""
    lab = labels[labels >= 0]
    m = int(lab.shape[0])
    counts = np.unique(lab, return_counts=True)[1].astype(np.int64)
    r = int((counts * (counts - 1) // 2).sum())
    return r, m * (m - 1) // 2 - r


def related_unrelated_scores(labels, pairwise_similarity_metrics, device='cuda:0'):
    """The two groups the reference compares (helpers.py:88-97) as ASCENDING device tensors: scores of same-label pairs and
    of different-label pairs of the strict lower triangle, rows labelled -1 dropped."""
    labels = np.asarray(labels.values if hasattr(labels, 'values') else labels).reshape(-1)
    assert labels.shape[0] == pairwise_similarity_metrics.shape[0]
    assert pairwise_similarity_metrics.shape[0] == pairwise_similarity_metrics.shape[1]
    lab_i = np.where(labels >= 0, np.unique(labels, return_inverse=True)[1].reshape(-1), -1).astype(np.int32)  # any numeric dtype
    if isinstance(pairwise_similarity_metrics, torch.Tensor):
        sim = pairwise_similarity_metrics.to(device=device, dtype=torch.float32)
    else:
        sim = torch.from_numpy(np.ascontiguousarray(pairwise_similarity_metrics, dtype=np.float32)).to(device)
    if sim.stride(1) != 1:
        sim = sim.contiguous()
    n = sim.shape[0]
    n_rel, n_unrel = _group_sizes(lab_i)
    rel = torch.empty(max(n_rel, 1), dtype=torch.float32, device=sim.device)
    unrel = torch.empty(max(n_unrel, 1), dtype=torch.float32, device=sim.device)
    cursors = torch.zeros(2, dtype=torch.int64, device=sim.device)
    lab_dev = torch.from_numpy(lab_i).to(sim.device)
    call('dae_pair_partition', sim.data_ptr(), sim.stride(0), n, lab_dev.data_ptr(), rel.data_ptr(), unrel.data_ptr(), cursors.data_ptr(),
         _stream())
    got = cursors.cpu().numpy()
    if int(got[0]) != n_rel or int(got[1]) != n_unrel:
        raise RuntimeError('dae_pair_partition wrote %s pairs, expected (%d, %d)' % (got.tolist(), n_rel, n_unrel))
    return torch.sort(rel[:n_rel])[0], torch.sort(unrel[:n_unrel])[0]   # device radix sort (library call), keys only


def auroc_from_groups(related_sorted, unrelated_sorted):
    """AUROC with 'Related' as the positive class (helpers.py:99-100) = (#(r > u) + #(r == u) / 2) / (R U), counted exactly
    on the device: the smaller group queries the larger (sorted) one."""
    n_rel, n_unrel = int(related_sorted.shape[0]), int(unrelated_sorted.shape[0])
    if n_rel == 0 or n_unrel == 0:
        return float('nan'), 0
    acc = torch.zeros(1, dtype=torch.int64, device=related_sorted.device)
    if n_rel <= n_unrel:
        call('dae_auroc_count', related_sorted.data_ptr(), n_rel, unrelated_sorted.data_ptr(), n_unrel, 1, acc.data_ptr(), _stream())
    else:
        call('dae_auroc_count', unrelated_sorted.data_ptr(), n_unrel, related_sorted.data_ptr(), n_rel, 0, acc.data_ptr(), _stream())
    twice = int(acc.item())
    return twice / (2.0 * n_rel * n_unrel), twice


def _box_stats(d):
    """Quartiles (linear interpolation, as np.percentile / plt.boxplot) and Tukey whiskers of one ASCENDING device tensor."""
    n = int(d.shape[0])
    if n == 0:
        return {'n': 0}

    def pct(q):
        pos = q * (n - 1)
        i0 = int(np.floor(pos))
        i1 = min(i0 + 1, n - 1)
        a, b = float(d[i0].item()), float(d[i1].item())
        return a + (b - a) * (pos - i0)
    q1, med, q3 = pct(0.25), pct(0.5), pct(0.75)
    iqr = q3 - q1
    dd = d.double()   # thresholds are compared in float64, like numpy does on the host
    lim = torch.tensor([q1 - 1.5 * iqr, q3 + 1.5 * iqr], dtype=torch.float64, device=d.device)
    lo_i = int(torch.searchsorted(dd, lim[:1], right=False).item())    # first datum >= q1 - 1.5 IQR
    hi_i = int(torch.searchsorted(dd, lim[1:], right=True).item())     # one past the last datum <= q3 + 1.5 IQR
    return {'q1': q1, 'median': med, 'q3': q3, 'whisker_lo': float(d[lo_i].item()) if lo_i < n else q1,
            'whisker_hi': float(d[hi_i - 1].item()) if hi_i > 0 else q3, 'mean': float(dd.mean().item()), 'n': n}


def visualize_pairwise_similarity(labels, pairwise_similarity_metrics, plot='boxplot', title=None, figsize=(16, 9), save_path=None,
                                  device='cuda:0', **plot_kwargs):
    """Reference helpers.visualize_pairwise_similarity (helpers.py:79-135), numeric part on the GPU: the related / unrelated
    split of the lower triangle, the AUROC shown in its ROC legend and the statistics its boxplot draws (computed on ALL pairs;
    the reference subsamples each group to 1e7 before plotting).  Drawing itself is out of scope (no matplotlib in this image):
    the numbers are returned, and written as JSON next to `save_path` when one is given."""
    assert plot in ['scatter', 'boxplot']
    rel, unrel = related_unrelated_scores(labels, pairwise_similarity_metrics, device=device)
    auroc, twice = auroc_from_groups(rel, unrel)
    out = {'title': title, 'auroc': auroc, 'twice_u': twice, 'related': _box_stats(rel), 'unrelated': _box_stats(unrel)}
    if save_path is not None:
        import json
        import os
        with open(os.path.splitext(save_path)[0] + '.json', 'w') as f:
            json.dump(out, f, indent=1)
    return out
