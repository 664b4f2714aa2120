"""A minimal stand-in for the TensorFlow-1 graph API, backed by PyTorch (CPU), sufficient to execute the UNMODIFIED
reference network code (agents/utils.py layers, agents/policies.py policies, agents/models.py agents) -- used only
by tests/golden/make_golden.py in the authoring container to produce golden training traces.

TensorFlow 1.12 itself cannot be installed here.  What this file restates is therefore not the reference's
networks (those run from the reference's own source) but the documented semantics of the ~35 TF primitives that
source calls: array ops (concat, split, squeeze, expand_dims, transpose, reshape, slice, boolean_mask, one_hot),
math (matmul, sigmoid, tanh, relu, softmax, log, square, clip_by_value, reduce_sum/mean, argmax), variables and
scopes, `tf.gradients` (torch.autograd), `tf.clip_by_global_norm` and `tf.train.RMSPropOptimizer` (formulas from
the TF 1.x sources: `scale = clip * min(1/norm, 1/clip)`; `ms += (g*g - ms) * (1 - decay)`, slot initialised to
ones, `mom = momentum*mom + lr*g/sqrt(ms + epsilon)`, `var -= mom`).

Graph mode is reproduced lazily: every op returns a Node holding a closure and a prototype value computed on zero
placeholders (that is where static shapes come from); `Session.run(fetches, feed_dict)` evaluates the closures with
the fed values, memoised per run.
"""
import contextlib

import numpy as np
import torch

float32 = torch.float32
int32 = torch.int32


class Dim(int):
    @property
    def value(self):
        return int(self)


def _t(a):
    if isinstance(a, torch.Tensor):
        return a
    if isinstance(a, np.ndarray):
        return torch.as_tensor(a)
    return a


def _proto(a):
    if isinstance(a, Node):
        return a.proto
    if isinstance(a, (list, tuple)):
        return [_proto(x) for x in a]
    return _t(a)


def _eval(a, env):
    if isinstance(a, Node):
        return a.eval(env)
    if isinstance(a, (list, tuple)):
        return [_eval(x, env) for x in a]
    return _t(a)


_DUMMY = torch.zeros(())


class Node:
    __array_priority__ = 1000.0          # numpy scalars defer to the reflected operators below

    def __init__(self, fn, args=(), proto=None):
        self.fn, self.args = fn, tuple(args)
        self.proto = proto if proto is not None else fn(*[_proto(a) for a in self.args])

    @property
    def shape(self):
        return tuple(Dim(d) for d in self.proto.shape)

    def get_shape(self):
        return self.shape

    @property
    def dtype(self):
        return self.proto.dtype

    def eval(self, env):
        k = id(self)
        if k not in env:
            env[k] = self.fn(*[_eval(a, env) for a in self.args])
        return env[k]

    def __add__(self, o): return Node(lambda a, b: a + b, (self, o))
    def __radd__(self, o): return Node(lambda a, b: b + a, (self, o))
    def __sub__(self, o): return Node(lambda a, b: a - b, (self, o))
    def __rsub__(self, o): return Node(lambda a, b: b - a, (self, o))
    def __mul__(self, o): return Node(lambda a, b: a * b, (self, o))
    def __rmul__(self, o): return Node(lambda a, b: b * a, (self, o))
    def __truediv__(self, o): return Node(lambda a, b: a / b, (self, o))
    def __neg__(self): return Node(lambda a: -a, (self,))
    def __getitem__(self, idx): return Node(lambda a: a[idx], (self,))
    def __hash__(self): return id(self)
    def __eq__(self, o): return self is o


class Placeholder(Node):
    def __init__(self, dtype, shape):
        self.fn, self.args = None, ()
        self.proto = torch.zeros([int(s) for s in shape], dtype=dtype)

    def eval(self, env):
        if id(self) not in env:
            raise KeyError('placeholder of shape %r was not fed' % (tuple(self.proto.shape),))
        return env[id(self)]


class Variable(Node):
    def __init__(self, name, value):
        self.fn, self.args, self.name = None, (), name
        self.tensor = torch.tensor(np.asarray(value, dtype=np.float32), requires_grad=True)
        self.proto = self.tensor

    def eval(self, env):
        return self.tensor

    def assign(self, value):
        return _Assign(self, value)


# ---- graph state ------------------------------------------------------------------------------------------------
_vars = {}
_scope = []
_slots = {}


def reset_default_graph():
    _vars.clear(); _slots.clear(); del _scope[:]


def set_random_seed(seed):
    pass


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    _scope.append(name)
    try:
        yield
    finally:
        _scope.pop()


def get_variable(name, shape=None, initializer=None, **kw):
    full = '/'.join(_scope + [name])
    if full not in _vars:
        shape = [int(s) for s in shape]
        _vars[full] = Variable(full, initializer(shape, float32, partition_info=None))
    return _vars[full]


def trainable_variables(scope=None):
    return [v for n, v in _vars.items() if scope is None or n.startswith(scope)]


def global_variables_initializer():
    return Node(None, (), proto=_DUMMY)


def constant_initializer(value=0.0):
    return lambda shape, dtype=None, partition_info=None: np.full(tuple(shape), value, dtype=np.float32)


def placeholder(dtype, shape=None, name=None):
    return Placeholder(dtype, shape)


# ---- array / math ops -------------------------------------------------------------------------------------------
def expand_dims(input, axis=None, dim=None):
    ax = axis if axis is not None else dim
    return Node(lambda a: torch.unsqueeze(a, ax), (input,))


def squeeze(input, axis=None):
    return Node((lambda a: torch.squeeze(a)) if axis is None else (lambda a: torch.squeeze(a, axis)), (input,))


def concat(values=None, axis=None, **kw):
    vals = list(values)
    return Node(lambda vs: torch.cat(vs, dim=axis), (vals,))


def split(value=None, num_or_size_splits=None, axis=0, **kw):
    n = int(num_or_size_splits)
    return [Node(lambda a, k=k: torch.chunk(a, n, dim=axis)[k], (value,)) for k in range(n)]


def transpose(a, perm=None):
    if perm is None:
        return Node(lambda x: x.permute(*reversed(range(x.dim()))), (a,))
    return Node(lambda x: x.permute(*perm), (a,))


def reshape(tensor, shape):
    shp = [int(s) for s in shape]
    return Node(lambda a: a.reshape(shp), (tensor,))


def slice(input_, begin, size):  # noqa: A001 (TF name)
    def f(a):
        idx = tuple(builtins_slice(int(b), int(b) + int(s)) for b, s in zip(begin, size))
        return a[idx]
    return Node(f, (input_,))


import builtins as _b  # noqa: E402
builtins_slice = _b.slice


def boolean_mask(tensor, mask):
    idx = torch.as_tensor(np.nonzero(np.asarray(mask))[0])
    return Node(lambda a: a.index_select(0, idx), (tensor,))


def one_hot(indices, depth, axis=-1):
    assert axis == -1
    return Node(lambda a: torch.nn.functional.one_hot(a.long(), int(depth)).float(), (indices,))


def matmul(a, b):
    return Node(lambda x, y: x @ y, (a, b))


def _unary(f):
    return lambda x, *a, **k: Node(f, (x,))


tanh = _unary(torch.tanh)
log = _unary(torch.log)
square = _unary(lambda a: a * a)


def clip_by_value(t, lo, hi):
    return Node(lambda a: torch.clamp(a, lo, hi), (t,))


def reduce_sum(x, axis=None, keepdims=False):
    return Node((lambda a: a.sum()) if axis is None else (lambda a: a.sum(dim=axis, keepdim=keepdims)), (x,))


def reduce_mean(x, axis=None, keepdims=False):
    return Node((lambda a: a.mean()) if axis is None else (lambda a: a.mean(dim=axis, keepdim=keepdims)), (x,))


def argmax(x, axis=None, **kw):
    return Node(lambda a: torch.argmax(a, dim=axis), (x,))


class _Assign(Node):
    """`var.assign(value)`; inside tf.group all values are computed before any variable is written (TF leaves the
    order of grouped assigns unspecified; simultaneous assignment is the evident intent of the consensus update)."""

    def __init__(self, var, value):
        self.fn, self.args, self.var, self.value, self.proto = True, (), var, value, _DUMMY

    def eval(self, env):
        with torch.no_grad():
            self.var.tensor.copy_(_eval(self.value, env))
        return None


def group(*nodes):
    def run_group(env):
        vals = [(n, _eval(n.value, env).detach().clone()) for n in nodes if isinstance(n, _Assign)]
        with torch.no_grad():
            for n, v in vals:
                n.var.tensor.copy_(v)
        for n in nodes:
            if not isinstance(n, _Assign):
                n.eval(env)
        return None
    g = Node(None, (), proto=_DUMMY)
    g.fn, g.eval = True, run_group
    return g


class nn:  # noqa: N801 (TF name)
    sigmoid = staticmethod(_unary(torch.sigmoid))
    tanh = staticmethod(_unary(torch.tanh))
    relu = staticmethod(_unary(torch.relu))
    softmax = staticmethod(_unary(lambda a: torch.softmax(a, dim=-1)))


class summary:  # noqa: N801
    @staticmethod
    def scalar(name, tensor):
        return Node(None, (), proto=_DUMMY)

    @staticmethod
    def merge(nodes):
        return Node(lambda: None, (), proto=_DUMMY)


# ---- autodiff / optimizer ---------------------------------------------------------------------------------------
def gradients(ys, xs):
    xs = list(xs)

    def all_grads(y, *ws):
        g = torch.autograd.grad(y, ws, allow_unused=True, retain_graph=True)
        return [torch.zeros_like(w) if gi is None else gi for gi, w in zip(g, ws)]
    joint = Node(all_grads, [ys] + xs)
    return [Node(lambda g, k=k: g[k], (joint,)) for k in range(len(xs))]


def clip_by_global_norm(t_list, clip_norm):
    t_list = list(t_list)
    norm = Node(lambda ts: torch.sqrt(sum((t * t).sum() for t in ts)), (t_list,))
    c = float(clip_norm)
    scale = Node(lambda n: c * torch.minimum(1.0 / n, torch.tensor(1.0) / c), (norm,))
    return [Node(lambda t, s: t * s, (t, scale)) for t in t_list], norm


class _RMSProp:
    def __init__(self, learning_rate, decay=0.9, momentum=0.0, epsilon=1e-10):
        self.lr, self.decay, self.momentum, self.eps = learning_rate, float(decay), float(momentum), float(epsilon)

    def apply_gradients(self, grads_and_vars):
        gv = list(grads_and_vars)
        grads, vs = [g for g, _ in gv], [v for _, v in gv]

        def step(lr, gs):
            with torch.no_grad():
                gs = [g.detach().clone() for g in gs]                 # all gradients are taken before any update
                for g, v in zip(gs, vs):
                    ms = _slots.setdefault((id(self), v.name, 'rms'), torch.ones_like(v.tensor))
                    mom = _slots.setdefault((id(self), v.name, 'mom'), torch.zeros_like(v.tensor))
                    ms.add_((g * g - ms) * (1.0 - self.decay))
                    mom.mul_(self.momentum).add_(g * lr / torch.sqrt(ms + self.eps))
                    v.tensor.sub_(mom)
            return None
        return Node(step, (self.lr, grads), proto=_DUMMY)


class train:  # noqa: N801
    RMSPropOptimizer = _RMSProp

    class Saver:
        def __init__(self, *a, **k):
            pass


class ConfigProto:
    def __init__(self, *a, **k):
        pass


class Session:
    def __init__(self, *a, **k):
        pass

    def run(self, fetches, feed_dict=None):
        env = {}
        for ph, val in (feed_dict or {}).items():
            env[id(ph)] = torch.as_tensor(np.asarray(val)).to(ph.proto.dtype).reshape(ph.proto.shape)

        def out(f):
            if isinstance(f, (list, tuple)):
                return [out(x) for x in f]
            r = f.eval(env) if f.fn is not None or isinstance(f, (Placeholder, Variable)) else None
            return r.detach().numpy().copy() if isinstance(r, torch.Tensor) else r
        return out(fetches)


def variable_values():
    return {n: v.tensor.detach().numpy().copy() for n, v in _vars.items()}
