"""GPU versions of the evaluation helpers that consume transform()'s output (reference helpers.py:11-50 and the
nearest-article lookup of main_autoencoder.py:307-318,352-359) -- SURVEY section 8f, rank 1.

    pairwise_similarity(in_df, norm='', metric='cosine', set_diagonal_zero=True) -> ndarray [N, N]     (reference signature)
    nearest_neighbors(embeddings, metric='cosine', chunk=8192) -> (index[N], score[N])                (no N x N matrix on the host)

Dense inputs (embeddings) go through the tcgen05 bf16x3 GEMM on row-normalised operands; sparse inputs (count / tf-idf
matrices) through the CSR encode kernel against the dense transpose.  No CPU path.
"""
import numpy as np
import scipy.sparse as sp
import torch

from . import _cabi
from ._cabi import call
from .engine import DeviceCSR

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


def pairwise_similarity(in_df, norm='', metric='cosine', set_diagonal_zero=True, device='cuda:0'):
    """Reference helpers.pairwise_similarity: optional `norm` ('l1','l2','max'), then cosine similarity or the linear kernel
    of every pair of rows, diagonal zeroed.  Returns a float32 ndarray [N, N]."""
    assert metric in ['cosine', 'linear kernel']
    assert norm in _NORM
    if sp.issparse(in_df):
        return _pairwise_sparse(in_df, norm, metric, set_diagonal_zero, device)
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
    return out.cpu().numpy()


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
    dense_t = torch.zeros(f, n, dtype=torch.float32, device=device)
    rows = torch.repeat_interleave(torch.arange(n, device=device), (csr.indptr[1:] - csr.indptr[:-1]))
    dense_t[csr.indices.long(), rows] = csr.values
    out = torch.empty(n, n, dtype=torch.float32, device=device)
    zero_b = torch.zeros(n, dtype=torch.float32, device=device)
    call('dae_encode_csr_fwd', csr.indptr.data_ptr(), csr.indices.data_ptr(), csr.values.data_ptr(), None, n, f, n, 1.0,
         dense_t.data_ptr(), zero_b.data_ptr(), _cabi.ACT['none'], out.data_ptr(), n, None, None, None, 0, _stream())
    if set_diagonal_zero:
        out.diagonal().zero_()
    return out.cpu().numpy()


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
