"""CPU ORACLE (test infrastructure, NOT product code).

A CPU restatement of the reference's DAE-with-triplet-loss training step, op for op, on
torch-CPU tensors (the reference's arithmetic lives in TensorFlow 1.12.0, which is pinned in
/root/reference/requirements.txt:4 and is not installable offline, so every `tf.*` call is
restated with the torch op that has the same published semantics; gradients come from
autograd exactly like `Optimizer.minimize`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.  The product path (dae_rnn_news_recommendation_b200) never does.

Parity pin: tests/golden/*.npz are produced by executing the reference's *own* python files
under the TF1 shim in oracle/tf1_shim (see oracle/gen_golden.py); tests/test_oracle_golden.py
checks this restatement against them, and tests/test_reference_test_ports.py re-runs the
reference's in-test NumPy brute-force loops against it.

All citations are relative to /root/reference/.
"""
import numpy as np
import scipy.sparse as sp
import torch

EPS = 1e-16


# --------------------------------------------------------------------------------------
# activations  (autoencoder/autoencoder.py:380-387, 402-409: anything but sigmoid/tanh = identity)
# --------------------------------------------------------------------------------------
def _act(name):
    if name == 'sigmoid':
        return torch.sigmoid
    if name == 'tanh':
        return torch.tanh
    return lambda x: x


def to_torch_sparse(m, dtype):
    """scipy sparse -> torch sparse COO with sorted indices (autoencoder/utils.py:162-180)."""
    m = sp.csr_matrix(m)
    m.sort_indices()
    coo = m.tocoo()
    idx = torch.from_numpy(np.vstack([coo.row, coo.col]).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(coo.data.astype(np.float64)).to(dtype),
                                   size=coo.shape).coalesce()


def _as_dense(x, dtype):
    if sp.issparse(x):
        return torch.from_numpy(np.asarray(x.todense(), dtype=np.float64)).to(dtype)
    if isinstance(x, torch.Tensor):
        return x.to_dense().to(dtype) if x.is_sparse else x.to(dtype)
    return torch.from_numpy(np.asarray(x, dtype=np.float64)).to(dtype)


# --------------------------------------------------------------------------------------
# encode / decode   (autoencoder/autoencoder.py:389, :411)
# --------------------------------------------------------------------------------------
def encode(xc, W, bh, enc_act_func):
    """E = f(Xc.W + bh) - f(bh)   (autoencoder.py:389).  xc: torch sparse or dense."""
    f = _act(enc_act_func)
    a = (torch.sparse.mm(xc, W) if xc.is_sparse else xc @ W) + bh
    return f(a) - f(bh)


def decode(E, W, bv, dec_act_func):
    """D = g(E.W^T + bv)   (autoencoder.py:411)."""
    return _act(dec_act_func)(E @ W.t() + bv)


# --------------------------------------------------------------------------------------
# weighted_loss   (autoencoder/triplet_loss_utils.py:262-277)
# --------------------------------------------------------------------------------------
def row_loss(x_dense, D, loss_func):
    if loss_func == 'cross_entropy':  # :269   (1.-decode+1e-16 evaluated left to right)
        return -torch.sum(x_dense * torch.log(D + EPS) + (1. - x_dense) * torch.log(1. - D + EPS), 1)
    if loss_func == 'mean_squared':  # :271
        return torch.sum((x_dense - D) ** 2, 1)
    if loss_func == 'cosine_proximity':  # :273  tf.nn.l2_normalize: x * rsqrt(max(sum x^2, 1e-12))
        def l2n(v):
            return v * torch.rsqrt(torch.clamp(torch.sum(v * v, 1, keepdim=True), min=1e-12))
        return -torch.sum(l2n(x_dense) * l2n(D), 1)
    raise AssertionError(loss_func)


def weighted_loss(x_dense, D, loss_func='cross_entropy', weight=None):
    if weight is None:
        weight = torch.ones(x_dense.shape[0], dtype=D.dtype)  # :266
    ell = row_loss(x_dense, D, loss_func)
    return torch.sum(ell * weight) / (torch.sum(weight) + EPS)  # :275


# --------------------------------------------------------------------------------------
# masks   (triplet_loss_utils.py:6-76)
# --------------------------------------------------------------------------------------
def anchor_positive_mask(labels):
    n = labels.shape[0]
    return (~torch.eye(n, dtype=torch.bool)) & (labels[None, :] == labels[:, None])


def anchor_negative_mask(labels):
    return ~(labels[None, :] == labels[:, None])


def triplet_mask(labels):
    n = labels.shape[0]
    ne = ~torch.eye(n, dtype=torch.bool)
    distinct = ne[:, :, None] & ne[:, None, :] & ne[None, :, :]
    leq = labels[None, :] == labels[:, None]
    return distinct & (leq[:, :, None] & ~leq[:, None, :])


# --------------------------------------------------------------------------------------
# batch_all   (triplet_loss_utils.py:79-131) -- materialises the B^3 tensors like the reference
# --------------------------------------------------------------------------------------
def batch_all_triplet_loss(labels, E, pos_triplets_only=False):
    S = E @ E.t()  # :93
    dist = -S[:, :, None] + S[:, None, :]  # :106   dist[i,j,k] = S_ik - S_ij
    valid = triplet_mask(labels).to(E.dtype)  # :110
    n_valid = valid.sum()
    pos = ((valid * dist) > 1e-16).to(E.dtype)  # :114
    n_pos = pos.sum()
    mask, n = (pos, n_pos) if pos_triplets_only else (valid, n_valid)
    loss = (torch.nn.functional.softplus(dist) * mask).sum() / (n + EPS)  # :126-127  -log_sigmoid(-d) = softplus(d)
    w = mask.sum((1, 2)) + mask.sum((0, 1)) + mask.sum((0, 2))  # :129
    return loss, w, n_pos / (n_valid + EPS), n_pos


# --------------------------------------------------------------------------------------
# batch_hard   (triplet_loss_utils.py:202-259)
# --------------------------------------------------------------------------------------
def batch_hard_triplet_loss(labels, E):
    S = E @ E.t()  # :219
    ap = anchor_positive_mask(labels).to(E.dtype)
    m = torch.amax(S, 1, keepdim=True)  # :227   (amax: ties share the gradient, like tf.reduce_max)
    hp = torch.amin(S + m * (1.0 - ap), 1, keepdim=True)  # :228-231
    an = anchor_negative_mask(labels).to(E.dtype)
    hn = torch.amax(an * S, 1, keepdim=True)  # :240-243
    td = torch.clamp(hn - hp, min=0.0)  # :247
    c = (td > 0.0).to(E.dtype)  # :249
    w = c.squeeze(1) + (c * (S == hp).to(E.dtype)).sum(0) + (c * (S == hn).to(E.dtype)).sum(0)  # :251-253
    loss = (torch.nn.functional.softplus(td) * c).sum() / (c.sum() + EPS)  # :256-257
    return loss, w, c.sum() / float(labels.shape[0]), c.sum()


def explicit_triplet_loss(E, Ep, En):
    """autoencoder/autoencoder_triplet.py:308-311: mean softplus(e.e_neg - e.e_pos)."""
    return torch.nn.functional.softplus(-((E * Ep) - (E * En)).sum(1)).mean()


# --------------------------------------------------------------------------------------
# host utilities restated  (autoencoder/utils.py)
# --------------------------------------------------------------------------------------
def xavier_bounds(fan_in, fan_out, const=1):
    """utils.py:24-25."""
    hi = const * np.sqrt(6.0 / (fan_in + fan_out))
    return -hi, hi


def masking_noise(X, v, rng=np.random):
    """utils.py:94-115 (global NumPy RNG stream: one rand(nnz) draw in COO order)."""
    assert 0. <= v <= 1.
    if isinstance(X, np.ndarray):
        mask = rng.choice(a=[0, 1], size=X.shape, p=[v, 1 - v])
        return mask * X
    coo = X.tocoo(True)
    keep = rng.rand(coo.nnz) >= v
    return sp.coo_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=coo.shape).tocsr()


def decay_noise(X, v):
    """utils.py:147-159."""
    return X.copy() * (1. - v)


def resolve_batch_size(n, batch_size):
    """utils.py:47-48."""
    assert batch_size > 0.
    if batch_size < 1.:
        batch_size = max(round(n * batch_size), 1)
    return int(batch_size)


def gen_batch_indices(n, batch_size, rng=np.random, random=True):
    """utils.py:50-53: one np.random.shuffle of list(range(n)), then consecutive slices."""
    bs = resolve_batch_size(n, batch_size)
    index = list(range(n))
    if random:
        rng.shuffle(index)
    return [np.asarray(index[i:i + bs], dtype=np.int64) for i in range(0, n, bs)]


# --------------------------------------------------------------------------------------
# the model + one training step   (autoencoder.py:206-246, 417-477)
# --------------------------------------------------------------------------------------
class OracleDAE:
    """Parameters + optimizer slots; `step` = one `session.run([train_step, losses...])`."""

    def __init__(self, W0, bh0=None, bv0=None, enc_act_func='sigmoid', dec_act_func='sigmoid',
                 loss_func='cross_entropy', opt='gradient_descent', learning_rate=0.1, momentum=0.5,
                 alpha=1.0, triplet_strategy='batch_all', dtype=torch.float32):
        self.dtype = dtype
        F, H = W0.shape
        t = lambda a: torch.from_numpy(np.array(a, dtype=np.float64)).to(dtype).requires_grad_(True)
        self.W = t(W0)
        self.bh = t(np.zeros(H) if bh0 is None else bh0)
        self.bv = t(np.zeros(F) if bv0 is None else bv0)
        self.enc_act_func, self.dec_act_func, self.loss_func = enc_act_func, dec_act_func, loss_func
        self.opt, self.lr, self.momentum, self.alpha = opt, learning_rate, momentum, alpha
        self.triplet_strategy = triplet_strategy
        self.t = 0
        ps = self.params()
        # TF-1.12 slot initialisation: Adagrad accumulator 0.1, Momentum/Adam zeros.
        self.slot1 = [torch.full_like(p, 0.1) if opt == 'ada_grad' else torch.zeros_like(p) for p in ps]
        self.slot2 = [torch.zeros_like(p) for p in ps]

    def params(self):
        return [self.W, self.bh, self.bv]

    def _sparse_or_dense(self, x):
        if sp.issparse(x):
            return to_torch_sparse(x, self.dtype)
        return _as_dense(x, self.dtype)

    # ---- forward of the whole graph (autoencoder.py:371-442)
    def forward(self, x, xc, labels=None):
        xd = _as_dense(x, self.dtype)
        E = encode(self._sparse_or_dense(xc), self.W, self.bh, self.enc_act_func)
        D = decode(E, self.W, self.bv, self.dec_act_func)
        out = {'encode': E, 'decode': D}
        if self.triplet_strategy != 'none':
            lab = torch.from_numpy(np.asarray(labels, dtype=np.float32).reshape(-1))  # fed as 'float' (:352)
            if self.triplet_strategy == 'batch_all':
                tl, w, frac, num = batch_all_triplet_loss(lab, E)
            else:
                tl, w, frac, num = batch_hard_triplet_loss(lab, E)
            w = w.detach()
            ael = weighted_loss(xd, D, self.loss_func, w)
            out.update(triplet_loss=tl, autoencoder_loss=ael, cost=ael + self.alpha * tl,  # :438
                       fraction=frac, num=num, weight=w)
        else:
            ael = weighted_loss(xd, D, self.loss_func)
            out.update(autoencoder_loss=ael, cost=ael)  # :441
        return out

    def forward_explicit(self, xs, xcs):
        """DenoisingAutoencoderTriplet graph (autoencoder_triplet.py:256-258, 286-288, 303-314)."""
        Es, ael = [], 0.
        for x, xc in zip(xs, xcs):
            E = encode(self._sparse_or_dense(xc), self.W, self.bh, self.enc_act_func)
            D = decode(E, self.W, self.bv, self.dec_act_func)
            ael = ael + weighted_loss(_as_dense(x, self.dtype), D, self.loss_func)
            Es.append(E)
        tl = explicit_triplet_loss(*Es)
        return {'encode': Es[0], 'encode_pos': Es[1], 'encode_neg': Es[2], 'autoencoder_loss': ael,
                'triplet_loss': tl, 'cost': ael + self.alpha * tl}

    # ---- optimizers (autoencoder.py:451-472; TF-1.12 kernels' documented update rules)
    def apply_gradients(self, grads):
        self.t += 1
        with torch.no_grad():
            for p, g, s1, s2 in zip(self.params(), grads, self.slot1, self.slot2):
                if self.opt == 'gradient_descent':
                    p -= self.lr * g
                elif self.opt == 'ada_grad':  # accum += g^2 ; var -= lr * g * rsqrt(accum)
                    s1 += g * g
                    p -= self.lr * g / torch.sqrt(s1)
                elif self.opt == 'momentum':  # accum = mu*accum + g ; var -= lr*accum
                    s1.mul_(self.momentum).add_(g)
                    p -= self.lr * s1
                elif self.opt == 'adam':  # beta1 .9 beta2 .999 eps 1e-8
                    b1, b2, eps = 0.9, 0.999, 1e-8
                    lr_t = self.lr * np.sqrt(1 - b2 ** self.t) / (1 - b1 ** self.t)
                    s1.mul_(b1).add_((1 - b1) * g)
                    s2.mul_(b2).add_((1 - b2) * g * g)
                    p -= lr_t * s1 / (torch.sqrt(s2) + eps)
                else:
                    raise AssertionError(self.opt)

    def grads(self, out):
        return torch.autograd.grad(out['cost'], self.params())

    def step(self, x, xc, labels=None):
        out = self.forward(x, xc, labels)
        g = self.grads(out)
        self.apply_gradients(g)
        res = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
        res['grads'] = [t.detach().numpy() for t in g]
        return res

    def step_explicit(self, xs, xcs):
        out = self.forward_explicit(xs, xcs)
        g = self.grads(out)
        self.apply_gradients(g)
        res = {k: v.detach().numpy() for k, v in out.items()}
        res['grads'] = [t.detach().numpy() for t in g]
        return res

    def transform(self, data):
        """autoencoder.py:479-505 (encode of the whole set in one call)."""
        with torch.no_grad():
            return encode(self._sparse_or_dense(data), self.W, self.bh, self.enc_act_func).numpy()

    def get_parameters(self):
        return {'enc_w': self.W.detach().numpy().copy(), 'enc_b': self.bh.detach().numpy().copy(),
                'dec_b': self.bv.detach().numpy().copy()}


# --------------------------------------------------------------------------------------
# brute-force loops from the reference's own tests (independent of the tensor restatement)
# --------------------------------------------------------------------------------------
def batch_all_bruteforce(labels, E):
    """autoencoder/tests/test_triplet_loss_utils.py:95-122."""
    E = np.asarray(E, dtype=np.float64)
    S = E @ E.T
    n = len(labels)
    w = np.zeros(n)
    wpos = np.zeros(n)
    loss = lpos = 0.
    nv = npos = 0
    for i in range(n):
        for j in range(n):
            for k in range(n):
                if i == j or j == k or i == k:
                    continue
                if labels[i] == labels[j] and labels[i] != labels[k]:
                    d = S[i, k] - S[i, j]
                    l = np.log1p(np.exp(d))
                    w[[i, j, k]] += 1
                    loss += l
                    nv += 1
                    if d > 1e-16:
                        wpos[[i, j, k]] += 1
                        lpos += l
                        npos += 1
    return {'loss': loss / (nv + 1e-16), 'weight': w, 'fraction': npos / (nv + 1e-16), 'num': npos,
            'loss_pos': lpos / (npos + 1e-16), 'weight_pos': wpos}


def batch_hard_bruteforce(labels, E):
    """autoencoder/tests/test_triplet_loss_utils.py:163-194."""
    E = np.asarray(E, dtype=np.float64)
    S = E @ E.T
    n = len(labels)
    hp = np.full(n, np.nan); hpi = np.full(n, np.nan)
    hn = np.full(n, np.nan); hni = np.full(n, np.nan)
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            if labels[i] == labels[j]:
                if np.isnan(hp[i]) or S[i, j] < hp[i]:
                    hp[i], hpi[i] = S[i, j], j
            else:
                if np.isnan(hn[i]) or S[i, j] > hn[i]:
                    hn[i], hni[i] = S[i, j], j
    w = np.zeros(n)
    loss = 0.
    num = 0
    with np.errstate(invalid='ignore'):
        td = hn - hp
    for idx, val in enumerate(td):
        if val > 0:
            w[idx] += 1
            w[int(hpi[idx])] += 1
            w[int(hni[idx])] += 1
            loss += np.log1p(np.exp(val))
            num += 1
    return {'loss': loss / (num + 1e-16), 'weight': w, 'fraction': num / n, 'num': num}
