"""Minimal TensorFlow-1.x graph-mode shim on torch-CPU (TEST INFRASTRUCTURE -- part of oracle/).

Purpose: execute the reference's own, unmodified Python files (autoencoder/autoencoder.py, triplet_loss_utils.py,
utils.py) in this container, where tensorflow==1.12.0 (reference requirements.txt:4) cannot be installed, to
generate the golden vectors under tests/golden/ (see oracle/gen_golden.py).  Every `tf.*` symbol those files use is
mapped to the torch op with the same published semantics; graphs are lazy (placeholders + feed_dict), gradients
come from torch.autograd (like Optimizer.minimize), optimizers follow the TF-1.12 update rules.

Only what the reference's hot path touches is implemented.  Nothing in the product imports this.
"""
import contextlib

import numpy as np
import torch

float32 = torch.float32
bool = torch.bool  # noqa: A001  (tf.bool)
int32 = torch.int32
_DT = {'float': torch.float32, 'float32': torch.float32, float32: torch.float32}

_state = {'graph_seed': None, 'rng': None, 'variables': [], 'default_session': None}


def reset_default_graph():
    _state['variables'] = []
    _state['rng'] = None if _state['graph_seed'] is None else np.random.RandomState(_state['graph_seed'])


def set_random_seed(seed):
    _state['graph_seed'] = seed
    _state['rng'] = np.random.RandomState(seed)


def _rng():
    if _state['rng'] is None:
        _state['rng'] = np.random.RandomState()
    return _state['rng']


class _Shape(list):
    pass


class Tensor(object):
    """A lazy graph node."""
    __array_ufunc__ = None     # numpy operands defer to our reflected operators (ndarray * Tensor -> Tensor.__rmul__)
    __array_priority__ = 1000

    def __init__(self, fn, inputs=(), shape=None, name=None):
        self.fn, self.inputs, self._shape, self.name = fn, tuple(inputs), shape, name

    @property
    def shape(self):
        return _Shape(self._shape if self._shape is not None else [])

    def eval(self, feed_dict=None, session=None):
        sess = session or _state['default_session']
        return sess.run(self, feed_dict)

    # operators
    def __add__(self, o): return _bin(torch.add, self, o)
    def __radd__(self, o): return _bin(torch.add, o, self)
    def __sub__(self, o): return _bin(torch.sub, self, o)
    def __rsub__(self, o): return _bin(torch.sub, o, self)
    def __mul__(self, o): return _bin(torch.mul, self, o)
    def __rmul__(self, o): return _bin(torch.mul, o, self)
    def __truediv__(self, o): return _bin(torch.div, self, o)
    def __rtruediv__(self, o): return _bin(torch.div, o, self)
    def __neg__(self): return Tensor(lambda a: -a, [self], self._shape)
    def __getitem__(self, idx): return Tensor(lambda a: a[idx], [self])
    __hash__ = object.__hash__


class _SparseValue(object):
    def __init__(self, t):
        self.t = t  # torch sparse COO


def _const(x):
    if isinstance(x, Tensor):
        return x
    if isinstance(x, torch.Tensor):
        v = x
    else:
        a = np.asarray(x)
        if a.dtype == np.float64:
            a = a.astype(np.float32)  # python floats / float64 literals become float32 like TF's default
        v = torch.from_numpy(a) if a.ndim else torch.tensor(a.item(), dtype=torch.from_numpy(a.reshape(1)).dtype)
    return Tensor(lambda: v, [], list(v.shape))


def _bin(op, a, b):
    a, b = _const(a), _const(b)

    def f(x, y):
        if isinstance(x, torch.Tensor) and isinstance(y, torch.Tensor) and x.dtype != y.dtype:
            if x.is_floating_point() and not y.is_floating_point():
                y = y.to(x.dtype)
            elif y.is_floating_point() and not x.is_floating_point():
                x = x.to(y.dtype)
        return op(x, y)
    return Tensor(f, [a, b])


def _un(op, a, shape=None):
    a = _const(a)
    return Tensor(op, [a], shape if shape is not None else a._shape)


# ---- placeholders / variables -------------------------------------------------------------------------------------------
class _Placeholder(Tensor):
    def __init__(self, dtype, name, sparse_=False):
        Tensor.__init__(self, None, [], None, name)
        self.dtype, self.sparse = _DT.get(dtype, torch.float32), sparse_


def placeholder(dtype, shape=None, name=None):
    return _Placeholder(dtype, name)


def sparse_placeholder(dtype, shape=None, name=None):
    return _Placeholder(dtype, name, sparse_=True)


class Variable(Tensor):
    def __init__(self, initial_value, name=None, **_):
        Tensor.__init__(self, None, [], None, name)
        self.initial = _const(initial_value)
        self.value = None
        _state['variables'].append(self)

    def initialize(self, ctx):
        self.value = _eval(self.initial, ctx).detach().clone().to(torch.float32).requires_grad_(True)


def global_variables_initializer():
    vs = list(_state['variables'])
    return _Op(lambda ctx: [v.initialize(ctx) for v in vs] and None)


class _Op(Tensor):
    def __init__(self, run):
        Tensor.__init__(self, None, [])
        self.run_op = run


def zeros(shape, dtype=float32): return Tensor(lambda: torch.zeros(*[int(s) for s in shape], dtype=torch.float32), [])
def ones(shape, dtype=float32):
    s = _const(shape)
    return Tensor(lambda v: torch.ones(*([int(v)] if v.ndim == 0 else [int(i) for i in v]), dtype=torch.float32), [s])


def random_uniform(shape, minval=0, maxval=None, dtype=float32, seed=None):
    def f():
        return torch.from_numpy(_rng().uniform(minval, maxval, size=tuple(int(s) for s in shape)).astype(np.float32))
    return Tensor(f, [], list(shape))


def eye(n):
    n = _const(n)
    return Tensor(lambda v: torch.eye(int(v)), [n])


def shape(x):  # noqa: A001
    return Tensor(lambda a: torch.tensor(list(a.shape)), [_const(x)])


# ---- math ---------------------------------------------------------------------------------------------------------------------
def matmul(a, b): return Tensor(lambda x, y: x @ y, [_const(a), _const(b)], [None, None])
def transpose(a): return _un(lambda x: x.t(), a, [None, None])
def log(a): return _un(torch.log, a)
def log_sigmoid(a): return _un(torch.nn.functional.logsigmoid, a)
def squared_difference(a, b): return _bin(lambda x, y: (x - y) ** 2, a, b)
def multiply(a, b): return _bin(torch.mul, a, b)
def maximum(a, b): return _bin(lambda x, y: torch.maximum(x, torch.as_tensor(y, dtype=x.dtype)), a, b)
def logical_not(a): return _un(torch.logical_not, a)
def logical_and(a, b): return _bin(torch.logical_and, a, b)
def equal(a, b): return _bin(torch.eq, a, b)
def greater(a, b): return _bin(torch.gt, a, b)
def to_float(a): return _un(lambda x: x.to(torch.float32), a)
def cast(a, dtype): return _un(lambda x: x.to(dtype), a)
def squeeze(a): return _un(torch.squeeze, a, [])


def expand_dims(a, axis):
    a = _const(a)
    shp = list(a._shape) if a._shape else [None, None]
    shp.insert(axis if axis >= 0 else len(shp) + 1 + axis, 1)
    return Tensor(lambda x: x.unsqueeze(axis), [a], shp)


def _red(fn):
    def op(a, axis=None, keepdims=False, **kw):
        ax = kw.get('reduction_indices', axis)

        def f(x):
            if ax is None:
                return fn(x, tuple(range(x.dim())), keepdims)
            return fn(x, tuple(ax) if isinstance(ax, (list, tuple)) else (ax,), keepdims)
        return Tensor(f, [_const(a)])
    return op


reduce_sum = _red(lambda x, ax, k: torch.sum(x, dim=ax, keepdim=k))
reduce_mean = _red(lambda x, ax, k: torch.mean(x, dim=ax, keepdim=k))
reduce_max = _red(lambda x, ax, k: torch.amax(x, dim=ax, keepdim=k))  # ties share the gradient equally, like TF
reduce_min = _red(lambda x, ax, k: torch.amin(x, dim=ax, keepdim=k))


class _NN(object):
    sigmoid = staticmethod(lambda a: _un(torch.sigmoid, a))
    tanh = staticmethod(lambda a: _un(torch.tanh, a))

    @staticmethod
    def l2_normalize(a, axis=None, epsilon=1e-12, dim=None):
        ax = axis if axis is not None else dim
        return _un(lambda x: x * torch.rsqrt(torch.clamp(torch.sum(x * x, dim=ax, keepdim=True), min=epsilon)), a)


nn = _NN()
sigmoid, tanh = nn.sigmoid, nn.tanh


class _Sparse(object):
    placeholder = staticmethod(sparse_placeholder)

    @staticmethod
    def matmul(sp, dense):  # tf.sparse.matmul == sparse_tensor_dense_matmul
        return Tensor(lambda s, d: torch.sparse.mm(s, d) if s.is_sparse else s @ d, [sp, _const(dense)], [None, None])

    @staticmethod
    def to_dense(sp):
        return Tensor(lambda s: s.to_dense() if s.is_sparse else s, [sp])

    @staticmethod
    def reduce_sum(sp, axis=None):
        return Tensor(lambda s: torch.sparse.sum(s, dim=axis).to_dense() if axis is not None else torch.sparse.sum(s), [sp])


sparse = _Sparse()
sparse_tensor_dense_matmul = _Sparse.matmul


# ---- evaluation ---------------------------------------------------------------------------------------------------------------
class _Ctx(object):
    def __init__(self, feed):
        self.feed, self.memo = feed or {}, {}


def _feed_value(ph, v):
    if ph.sparse:
        idx, val, shp = v
        t = torch.sparse_coo_tensor(torch.from_numpy(np.asarray(idx).T.astype(np.int64)),
                                    torch.from_numpy(np.asarray(val).astype(np.float32)), size=tuple(int(s) for s in shp))
        return t.coalesce()
    if hasattr(v, 'values') and not isinstance(v, np.ndarray):  # pandas Series / DataFrame
        v = v.values
    return torch.from_numpy(np.asarray(v).astype(np.float32))


def _eval(node, ctx):
    if id(node) in ctx.memo:
        return ctx.memo[id(node)]
    if isinstance(node, _Placeholder):
        if node not in ctx.feed:
            raise KeyError('placeholder %r was not fed' % node.name)
        out = _feed_value(node, ctx.feed[node])
    elif isinstance(node, Variable):
        out = node.value
    elif isinstance(node, _Op):
        out = node.run_op(ctx)
    else:
        out = node.fn(*[_eval(i, ctx) for i in node.inputs])
    ctx.memo[id(node)] = out
    return out


class Session(object):
    def __init__(self, *a, **k):
        self.graph = None

    def __enter__(self):
        self._prev = _state['default_session']
        _state['default_session'] = self
        return self

    def __exit__(self, *exc):
        _state['default_session'] = self._prev
        return False

    def run(self, fetches, feed_dict=None):
        ctx = _Ctx(feed_dict)
        single = not isinstance(fetches, (list, tuple))
        outs = []
        for f in ([fetches] if single else fetches):
            if f is None:
                outs.append(None)
                continue
            v = _eval(f, ctx)
            if isinstance(v, torch.Tensor):
                v = v.detach()
                v = (v.to_dense() if v.is_sparse else v).numpy()
                if v.ndim == 0:
                    v = v[()]
            outs.append(v)
        return outs[0] if single else outs


# ---- training ------------------------------------------------------------------------------------------------------------------
class _Optimizer(object):
    def __init__(self, learning_rate, **kw):
        self.lr = learning_rate
        self.kw = kw
        self.slots = {}
        self.t = 0

    def minimize(self, loss):
        vs = list(_state['variables'])
        opt = self

        def run(ctx):
            cost = _eval(loss, ctx)
            grads = torch.autograd.grad(cost, [v.value for v in vs], allow_unused=True)
            opt.t += 1
            with torch.no_grad():
                for v, g in zip(vs, grads):
                    if g is None:
                        continue
                    opt.apply(v, g)
            return None
        return _Op(run)


class GradientDescentOptimizer(_Optimizer):
    def apply(self, v, g):
        v.value -= self.lr * g


class AdagradOptimizer(_Optimizer):
    def __init__(self, learning_rate, initial_accumulator_value=0.1, **kw):
        _Optimizer.__init__(self, learning_rate)
        self.init_acc = initial_accumulator_value

    def apply(self, v, g):
        acc = self.slots.setdefault(id(v), torch.full_like(v.value, self.init_acc))
        acc += g * g
        v.value -= self.lr * g / torch.sqrt(acc)


class MomentumOptimizer(_Optimizer):
    def __init__(self, learning_rate, momentum, **kw):
        _Optimizer.__init__(self, learning_rate)
        self.mu = momentum

    def apply(self, v, g):
        acc = self.slots.setdefault(id(v), torch.zeros_like(v.value))
        acc.mul_(self.mu).add_(g)
        v.value -= self.lr * acc


class AdamOptimizer(_Optimizer):
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **kw):
        _Optimizer.__init__(self, learning_rate)
        self.b1, self.b2, self.eps = beta1, beta2, epsilon

    def apply(self, v, g):
        m, s = self.slots.setdefault(id(v), (torch.zeros_like(v.value), torch.zeros_like(v.value)))
        lr_t = self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        m.mul_(self.b1).add_((1 - self.b1) * g)
        s.mul_(self.b2).add_((1 - self.b2) * g * g)
        v.value -= lr_t * m / (torch.sqrt(s) + self.eps)


class Saver(object):
    def __init__(self, *a, **k):
        self.vars = list(_state['variables'])

    def save(self, sess, path):
        np.savez(path + '.npz', **{v.name: v.value.detach().numpy() for v in self.vars})

    def restore(self, sess, path):
        with np.load(path + '.npz') as z:
            for v in self.vars:
                v.value = torch.from_numpy(z[v.name]).clone().requires_grad_(True)


class _Train(object):
    GradientDescentOptimizer = GradientDescentOptimizer
    AdagradOptimizer = AdagradOptimizer
    MomentumOptimizer = MomentumOptimizer
    AdamOptimizer = AdamOptimizer
    Saver = Saver


train = _Train()


# ---- no-op plumbing ----------------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def name_scope(name):
    yield


class _FileWriter(object):
    def __init__(self, *a, **k): pass
    def add_summary(self, *a, **k): pass
    def close(self): pass


class _Summary(object):
    FileWriter = _FileWriter
    histogram = staticmethod(lambda *a, **k: None)
    scalar = staticmethod(lambda *a, **k: None)
    merge_all = staticmethod(lambda *a, **k: None)


summary = _Summary()
