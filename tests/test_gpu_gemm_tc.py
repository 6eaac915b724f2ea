"""tcgen05 bf16x3 GEMM (dae_gemm_bf16x3 / dae_decode_fused_bf16x3) against fp64 matmul and against the CUDA-core path."""
import numpy as np
import pytest
import torch

from helpers import rel_err, random_csr, xavier

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(autouse=True, params=['auto', 'pair', 'lean'])
def _tile_engine(request):
    """Every test runs three times: with the default configuration (size-based choice between one-CTA tiles and CTA pairs, deepest
    operand rings), with the pairs forced (small, ragged and single-m-tile shapes through the cta_group::2 kernel), and with the
    lean 2-stage rings (128 x 128 decode tiles, 2-stage pairs)."""
    from dae_rnn_news_recommendation_b200 import _cabi
    _cabi.call('dae_gemm_config', 1 if request.param == 'pair' else -1, 1 if request.param == 'lean' else 0)
    yield
    _cabi.call('dae_gemm_config', -1, 0)


def _split(x, ld, ones_col=-1):
    from dae_rnn_news_recommendation_b200 import _cabi
    rows, cols = x.shape
    hi = torch.empty(rows, ld, dtype=torch.bfloat16, device=DEV)
    lo = torch.empty(rows, ld, dtype=torch.bfloat16, device=DEV)
    _cabi.call('dae_split_bf16', x.data_ptr(), rows, cols, x.stride(0), hi.data_ptr(), lo.data_ptr(), ld, ones_col, 1.0,
               torch.cuda.current_stream().cuda_stream)
    return hi, lo


def _gemm(M, N, K, A, a_mn, B, b_mn, C, **kw):
    from dae_rnn_news_recommendation_b200 import _cabi
    (ah, al), (bh, bl) = A, B
    _cabi.call('dae_gemm_bf16x3', M, N, K, kw.get('alpha', 1.0), ah.data_ptr(), al.data_ptr(), ah.stride(0), a_mn, bh.data_ptr(),
               bl.data_ptr(), bh.stride(0), b_mn, C.data_ptr(), C.stride(0), kw.get('n_store', 0), kw.get('special_col', -1),
               kw['special_out'].data_ptr() if kw.get('special_out') is not None else None, kw.get('k_splits', 1),
               kw.get('accumulate', 0), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()


def test_split_is_exact_to_2pow17():
    x = torch.randn(300, 200, device=DEV) * 3
    hi, lo = _split(x, 208, ones_col=203)
    rec = hi[:, :200].float() + lo[:, :200].float()
    assert float(((rec - x).abs() / x.abs().clamp_min(1e-30)).max()) < 2.0 ** -16
    assert float(hi[:, 200:203].abs().max()) == 0 and float(hi[:, 203].min()) == 1.0 and float(lo[:, 203].abs().max()) == 0


@pytest.mark.parametrize('M,N,K', [(128, 256, 64), (800, 800, 500), (200, 1000, 130), (1000, 501, 800)])
@pytest.mark.parametrize('a_mn,b_mn', [(0, 0), (1, 1), (0, 1), (1, 0)])
def test_gemm_all_majorness(M, N, K, a_mn, b_mn):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = torch.randn(M, K, device=DEV, generator=g)
    B = torch.randn(N, K, device=DEV, generator=g)
    want = (A.double() @ B.double().t()).cpu().numpy()
    pad = lambda n: (n + 7) // 8 * 8
    Aop = _split(A.t().contiguous(), pad(M)) if a_mn else _split(A, pad(K))
    Bop = _split(B.t().contiguous(), pad(N)) if b_mn else _split(B, pad(K))
    C = torch.full((M, N), float('nan'), device=DEV)
    _gemm(M, N, K, Aop, a_mn, Bop, b_mn, C)
    assert rel_err(C.cpu().numpy(), want) < 2e-5


def test_gemm_split_k_accumulate_and_special_column():
    M, N, K = 300, 501, 4000
    g = torch.Generator(device=DEV).manual_seed(1)
    A = torch.randn(M, K, device=DEV, generator=g)
    B = torch.randn(N, K, device=DEV, generator=g)
    want = (A.double() @ B.double().t()).cpu().numpy()
    Aop, Bop = _split(A, K), _split(B, K)
    C = torch.full((M, 500), float('nan'), device=DEV)
    sp = torch.full((M,), float('nan'), device=DEV)
    _gemm(M, N, K, Aop, 0, Bop, 0, C, n_store=500, special_col=500, special_out=sp, k_splits=7, alpha=0.5)
    assert rel_err(C.cpu().numpy(), 0.5 * want[:, :500]) < 2e-5
    assert rel_err(sp.cpu().numpy(), 0.5 * want[:, 500]) < 2e-5
    C2 = torch.ones(M, 500, device=DEV)
    _gemm(M, 500, K, Aop, 0, Bop, 0, C2, accumulate=1)
    assert rel_err(C2.cpu().numpy(), 1.0 + want[:, :500]) < 2e-5


@pytest.mark.parametrize('M,N,K,a_mn,b_mn', [(10000, 501, 800, 1, 1), (800, 500, 10000, 0, 1), (300, 700, 2000, 0, 0), (130, 90, 64, 0, 0)])
def test_gemm_stream_k(M, N, K, a_mn, b_mn):
    """k_splits = -1: stream-K (the tile x k-block units shared evenly by the SMs) -- the dW / dE shapes of the C2 step, with the
    [dW | dbv] special column and with accumulation into a non-zero C."""
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = torch.randn(M, K, device=DEV, generator=g)
    B = torch.randn(N, K, device=DEV, generator=g)
    want = (A.double() @ B.double().t()).cpu().numpy()
    pad = lambda n: (n + 7) // 8 * 8
    Aop = _split(A.t().contiguous(), pad(M)) if a_mn else _split(A, pad(K))
    Bop = _split(B.t().contiguous(), pad(N)) if b_mn else _split(B, pad(K))
    C = torch.full((M, N - 1), float('nan'), device=DEV)
    sp = torch.full((M,), float('nan'), device=DEV)
    _gemm(M, N, K, Aop, a_mn, Bop, b_mn, C, n_store=N - 1, special_col=N - 1, special_out=sp, k_splits=-1)
    assert rel_err(C.cpu().numpy(), want[:, :N - 1]) < 2e-5
    assert rel_err(sp.cpu().numpy(), want[:, N - 1]) < 2e-5
    C2 = torch.ones(M, N, device=DEV)
    _gemm(M, N, K, Aop, a_mn, Bop, b_mn, C2, k_splits=-1, accumulate=1, alpha=2.0)
    assert rel_err(C2.cpu().numpy(), 1.0 + 2.0 * want) < 2e-5


@pytest.mark.parametrize('M,N', [(800, 500), (130, 52), (64, 100), (333, 24)])
def test_gemm_sym_is_g_plus_gt_times_b(M, N):
    """dae_gemm_sym_bf16x3: C = alpha (G + G^T) B in one launch (k loop over G's columns, then over its rows), with and without
    accumulation, ragged sizes (M not a multiple of the 64-wide k block)."""
    from dae_rnn_news_recommendation_b200 import _cabi
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N)
    G = torch.randn(M, M, device=DEV, generator=g)
    Bm = torch.randn(M, N, device=DEV, generator=g)
    pad = lambda n: (n + 7) // 8 * 8
    Ghl, Bhl = _split(G, pad(M)), _split(Bm, pad(N))
    want = ((G.double() + G.double().t()) @ Bm.double()).cpu().numpy()
    st = torch.cuda.current_stream().cuda_stream
    C = torch.full((M, N), float('nan'), device=DEV)
    _cabi.call('dae_gemm_sym_bf16x3', M, N, 0.5, Ghl[0].data_ptr(), Ghl[1].data_ptr(), pad(M), Bhl[0].data_ptr(), Bhl[1].data_ptr(), pad(N),
               C.data_ptr(), N, 0, st)
    torch.cuda.synchronize()
    assert rel_err(C.cpu().numpy(), 0.5 * want) < 2e-5
    C2 = torch.ones(M, N, device=DEV)
    _cabi.call('dae_gemm_sym_bf16x3', M, N, 1.0, Ghl[0].data_ptr(), Ghl[1].data_ptr(), pad(M), Bhl[0].data_ptr(), Bhl[1].data_ptr(), pad(N),
               C2.data_ptr(), N, 1, st)
    torch.cuda.synchronize()
    assert rel_err(C2.cpu().numpy(), 1.0 + want) < 2e-5


@pytest.mark.parametrize('loss,dec,zscale', [('cross_entropy', 'sigmoid', 1.0), ('cross_entropy', 'sigmoid', 12.0), ('mean_squared', 'none', 1.0),
                                             ('mean_squared', 'tanh', 1.0), ('mean_squared', 'sigmoid', 1.0)])
def test_fused_decode_matches_unfused(loss, dec, zscale):
    """dae_decode_fused_bf16x3 == dae_sgemm + dae_decode_loss_bwd (the parity-checked CUDA-core path).  zscale = 12 drives many
    pre-activations past +-10 (the sigmoid/CE fast path must hand those chunks to the exact evaluation); some rows carry weight 0."""
    from dae_rnn_news_recommendation_b200 import _cabi
    from dae_rnn_news_recommendation_b200.engine import DeviceCSR
    B, F, H = 200, 1000, 52
    st = torch.cuda.current_stream().cuda_stream
    x = random_csr(B, F, 30, kind='tfidf' if loss != 'cross_entropy' else 'binary', seed=3)
    csr = DeviceCSR(x, DEV)
    E = torch.randn(B, H, device=DEV) * 0.3
    W = torch.from_numpy(xavier(F, H, 4) * 3 * zscale).to(DEV)
    bv = torch.randn(F, device=DEV) * 0.1
    w = torch.rand(B, device=DEV) * 5
    w[::7] = 0.0
    stats = torch.zeros(16, dtype=torch.float64, device=DEV)
    stats[5] = float(w.sum())
    Z = torch.empty(B, F, device=DEV)
    _cabi.call('dae_sgemm', B, F, H, 1.0, E.data_ptr(), H, 1, W.data_ptr(), H, 1, 0.0, Z.data_ptr(), F, st)
    rl = torch.empty(B, device=DEV)
    _cabi.call('dae_decode_loss_bwd', csr.indptr.data_ptr(), csr.indices.data_ptr(), csr.values.data_ptr(), None, B, F, bv.data_ptr(),
               _cabi.ACT[dec], _cabi.LOSS[loss], w.data_ptr(), stats.data_ptr(), Z.data_ptr(), F, rl.data_ptr(), st)
    Hp, Fp = 64, (F + 31) // 32 * 32
    Ehl, Whl = _split(E, Hp), _split(W, Hp)
    dzh = torch.full((B, Fp), float('nan'), dtype=torch.bfloat16, device=DEV)
    dzl = torch.full((B, Fp), float('nan'), dtype=torch.bfloat16, device=DEV)
    parts = torch.empty(2 * ((F + 255) // 256), B, device=DEV)
    tptr = torch.empty(B, 4 * ((F + 255) // 256) + 1, dtype=torch.int32, device=DEV)
    _cabi.call('dae_decode_fused_bf16x3', B, F, H, Ehl[0].data_ptr(), Ehl[1].data_ptr(), Hp, Whl[0].data_ptr(), Whl[1].data_ptr(), Hp,
               csr.indptr.data_ptr(), csr.indices.data_ptr(), csr.values.data_ptr(), None, bv.data_ptr(), _cabi.ACT[dec],
               _cabi.LOSS[loss], w.data_ptr(), stats.data_ptr(), dzh.data_ptr(), dzl.data_ptr(), Fp, parts.data_ptr(), tptr.data_ptr(), 0, st)
    rl2 = parts.view(-1)[:B]
    torch.cuda.synchronize()
    assert rel_err(rl2.cpu().numpy(), rl.cpu().numpy()) < 2e-5
    dz = (dzh.float() + dzl.float())[:, :F]
    assert rel_err(dz.cpu().numpy(), Z.cpu().numpy()) < 3e-5
    assert float(dzh[:, F:].float().abs().max()) == 0.0


def test_reduce_parts_and_finalize_with_partials():
    """dae_reduce_parts and the `parts` input of dae_step_finalize give the same weighted mean as a plain row-loss vector."""
    from dae_rnn_news_recommendation_b200 import _cabi
    B, P = 300, 7
    st = torch.cuda.current_stream().cuda_stream
    parts = torch.rand(P, B, device=DEV)
    w = torch.rand(B, device=DEV) * 3
    rl = torch.empty(B, device=DEV)
    _cabi.call('dae_reduce_parts', parts.data_ptr(), P, B, rl.data_ptr(), st)
    assert torch.allclose(rl, parts.sum(0), rtol=1e-6)
    out = []
    for use_parts in (False, True):
        stats = torch.zeros(16, dtype=torch.float64, device=DEV)
        stats[_cabi.STAT['sum_w']] = float(w.double().sum())
        _cabi.call('dae_step_finalize', None if use_parts else rl.data_ptr(), parts.data_ptr() if use_parts else None, P if use_parts else 0,
                   w.data_ptr(), B, 0, 1.0, stats.data_ptr(), None, None, st)
        torch.cuda.synchronize()
        out.append(float(stats[_cabi.STAT['ae_loss']]))
    want = float((parts.sum(0).double() * w.double()).sum() / w.double().sum())
    assert abs(out[0] - want) < 1e-6 * abs(want) and abs(out[1] - want) < 1e-6 * abs(want)
