"""Eager, GPU-backed versions of the reference's loss builders (reference autoencoder/triplet_loss_utils.py).

Same names, argument order and 4-tuple returns, but they take NumPy arrays and return NumPy values immediately (the
reference returns TF graph nodes that its tests evaluate with `tf.Session().run`), so the reference's tests port 1:1:

    loss, data_weight, fraction, num = batch_all_triplet_loss(False, labels, encode[, pos_triplets_only])
    loss, data_weight, fraction, num = batch_hard_triplet_loss(False, labels, encode)
    loss = weighted_loss(False, input_data, decode, loss_func='cross_entropy', weight=None)

Everything is computed by the kernels of libdae_sm100.so (the same ones `fit` uses); there is no CPU path.
"""
import numpy as np
import scipy.sparse as sp
import torch

from .. import _cabi
from .._cabi import call, STAT

_DEV = 'cuda:0'


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _prepare(labels, strategy):
    """label-sort the batch with dae_batch_prepare; returns (order, sorted labels, seg_lo, seg_hi, weight, stats)."""
    B = len(labels)
    lab = torch.from_numpy(np.asarray(labels, dtype=np.float32).reshape(-1)).to(_DEV)
    rows = torch.empty(B, dtype=torch.int32, device=_DEV)
    lab_s = torch.empty(B, device=_DEV)
    lo = torch.empty(B, dtype=torch.int32, device=_DEV)
    hi = torch.empty(B, dtype=torch.int32, device=_DEV)
    w = torch.empty(B, device=_DEV)
    stats = torch.zeros(_cabi.STAT_SLOTS, dtype=torch.float64, device=_DEV)
    call('dae_batch_prepare', None, 0, None, B, lab.data_ptr(), strategy, rows.data_ptr(), lab_s.data_ptr(), lo.data_ptr(),
         hi.data_ptr(), w.data_ptr(), stats.data_ptr(), _stream())
    return rows.long(), lab_s, lo, hi, w, stats


def _gram(E):
    B, H = E.shape
    S = torch.empty(B, B, device=_DEV)
    call('dae_sgemm', B, B, H, 1.0, E.data_ptr(), H, 1, E.data_ptr(), H, 1, 0.0, S.data_ptr(), B, _stream())
    return S


def batch_all_triplet_loss(sparse_input, input_label, encode, pos_triplets_only=False):
    """reference triplet_loss_utils.py:79-131 -> (loss, data_weight[B], fraction_positive, num_positive)."""
    B = len(input_label)
    order, _, lo, hi, w, stats = _prepare(input_label, 1)
    E = torch.from_numpy(np.asarray(encode, dtype=np.float32)).to(_DEV)[order].contiguous()
    S = _gram(E)
    G = torch.empty(B, B, device=_DEV)
    call('dae_triplet_batch_all', S.data_ptr(), B, B, lo.data_ptr(), hi.data_ptr(), G.data_ptr(), B, stats.data_ptr(),
         1 if pos_triplets_only else 0, None, None, 0, _stream())
    torch.cuda.synchronize()
    st = stats.cpu().numpy()
    n_valid, n_pos, tsum = st[STAT['n_valid']], st[STAT['num']], st[STAT['triplet_sum']]
    inv = np.empty(B, dtype=np.int64)
    inv[order.cpu().numpy()] = np.arange(B)
    if pos_triplets_only:
        C = G.abs()                                     # counts of positive triplets: C[i,j] = #k, C[i,k] = #j
        same = (torch.arange(B, device=_DEV)[None, :] >= lo[:, None]) & (torch.arange(B, device=_DEV)[None, :] < hi[:, None])
        weight = (C * same).sum(1) + C.sum(0)           # as anchor + as positive / negative
        loss = tsum / (n_pos + 1e-16)
        weight = weight.cpu().numpy()[inv]
    else:
        loss = tsum / (n_valid + 1e-16)
        weight = w.cpu().numpy()[inv]
    return np.float32(loss), weight.astype(np.float32), np.float32(n_pos / (n_valid + 1e-16)), np.float32(n_pos)


def batch_hard_triplet_loss(sparse_input, input_label, encode):
    """reference triplet_loss_utils.py:202-259 -> (loss, data_weight[B], fraction_active, num_active)."""
    B = len(input_label)
    lab = torch.from_numpy(np.asarray(input_label, dtype=np.float32).reshape(-1)).to(_DEV)
    E = torch.from_numpy(np.asarray(encode, dtype=np.float32)).to(_DEV).contiguous()
    S = _gram(E)
    G = torch.empty(B, B, device=_DEV)
    w = torch.empty(B, device=_DEV)
    stats = torch.zeros(_cabi.STAT_SLOTS, dtype=torch.float64, device=_DEV)
    call('dae_triplet_batch_hard', S.data_ptr(), B, B, lab.data_ptr(), G.data_ptr(), B, w.data_ptr(), stats.data_ptr(), _stream())
    torch.cuda.synchronize()
    st = stats.cpu().numpy()
    na = st[STAT['n_active']]
    return (np.float32(st[STAT['triplet_sum']] / (na + 1e-16)), w.cpu().numpy(), np.float32(na / B), np.float32(na))


def weighted_loss(sparse_input, input_data, decode, loss_func='cross_entropy', weight=None):
    """reference triplet_loss_utils.py:262-277: sum_i l_i w_i / (sum w + 1e-16), l_i the per-row CE / MSE / cosine loss."""
    from ..engine import DeviceCSR
    x = input_data if sp.issparse(input_data) else sp.csr_matrix(np.asarray(input_data, dtype=np.float32))
    csr = DeviceCSR(x, _DEV)
    B, F = x.shape
    D = torch.from_numpy(np.ascontiguousarray(np.asarray(decode, dtype=np.float32))).to(_DEV).clone()
    wt = None if weight is None else torch.from_numpy(np.asarray(weight, dtype=np.float32).reshape(-1)).to(_DEV)
    stats = torch.zeros(_cabi.STAT_SLOTS, dtype=torch.float64, device=_DEV)
    stats[STAT['sum_w']] = float(B) if wt is None else float(wt.double().sum())
    row_loss = torch.empty(B, device=_DEV)
    zero_bias = torch.zeros(F, device=_DEV)
    # `decode` is already activated: identity activation + zero bias make the kernel evaluate the loss on D itself
    call('dae_decode_loss_bwd', csr.indptr.data_ptr(), csr.indices.data_ptr(), csr.values.data_ptr(), None, B, F,
         zero_bias.data_ptr(), _cabi.ACT['none'], _cabi.LOSS[loss_func], None if wt is None else wt.data_ptr(), stats.data_ptr(),
         D.data_ptr(), F, row_loss.data_ptr(), _stream())
    call('dae_step_finalize', row_loss.data_ptr(), None, 0, None if wt is None else wt.data_ptr(), B, 0, 1.0, stats.data_ptr(),
         None, None, _stream())
    torch.cuda.synchronize()
    return np.float32(stats.cpu().numpy()[STAT['ae_loss']])
