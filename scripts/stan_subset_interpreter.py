"""A mechanical evaluator for the subset of the Stan language the reference's model files use
(bayes_drt/stan_model_files/*_modelcode.txt): data / transformed data / parameters / transformed parameters / model
blocks; int, real, vector, matrix declarations with <lower=0>; arithmetic with Stan's typing (int / int is integer
division, matrix * vector is a product, .* and ./ are element-wise), 1-based inclusive indexing, single-statement for
loops, the functions square, sqrt, sum, pi, rep_vector, append_row and the sampling statements std_normal, normal,
inv_gamma, exponential.

It reads the reference's OWN source text at generation time (nothing of it is copied) and evaluates
log p(theta(u) | data) [+ log|J| of the lower=0 transforms] with torch, so gradients come from autograd.  Densities keep
all their normalising constants (Stan's `~` drops those that do not depend on parameters); consumers therefore compare
differences of log-densities between points, and gradients.  A transformed parameter that violates its declared
<lower=0> makes Stan reject the point: the evaluator returns -inf.

Used by scripts/make_golden_stan_logdensity.py; not part of the product or of the tests' imports."""
import math
import re

import torch


class I(int):
    """Stan int: int / int is integer division."""
    def _w(self, v):
        return I(v) if isinstance(v, int) and not isinstance(v, bool) else v

    def __add__(self, o):
        return self._w(int(self) + o) if isinstance(o, int) else int(self) + o
    __radd__ = __add__

    def __sub__(self, o):
        return self._w(int(self) - o) if isinstance(o, int) else int(self) - o

    def __rsub__(self, o):
        return self._w(o - int(self)) if isinstance(o, int) else o - int(self)

    def __mul__(self, o):
        return self._w(int(self) * o) if isinstance(o, int) else o * int(self)
    __rmul__ = __mul__

    def __truediv__(self, o):
        return I(int(self) // int(o)) if isinstance(o, int) else int(self) / o

    def __rtruediv__(self, o):
        return I(int(o) // int(self)) if isinstance(o, int) else o / int(self)


def _t(v):
    if isinstance(v, V):
        return v.t
    return torch.as_tensor(float(v), dtype=torch.float64)


class V:
    """Stan real / vector / matrix value (0-, 1-, 2-dimensional float64 tensor)."""
    def __init__(self, t):
        self.t = t

    def __add__(self, o):
        return V(self.t + _t(o))
    __radd__ = __add__

    def __sub__(self, o):
        return V(self.t - _t(o))

    def __rsub__(self, o):
        return V(_t(o) - self.t)

    def __neg__(self):
        return V(-self.t)

    def __mul__(self, o):
        a, b = self.t, _t(o)
        if a.dim() == 0 or b.dim() == 0:
            return V(a * b)
        if a.dim() == 2 and b.dim() in (1, 2):
            return V(a @ b)
        raise TypeError('vector * vector is not used by the reference models')

    def __rmul__(self, o):
        return V(_t(o) * self.t)  # scalar on the left

    def __truediv__(self, o):
        b = _t(o)
        if b.dim() != 0:
            raise TypeError('division by a non-scalar must be ./')
        return V(self.t / b)

    def __rtruediv__(self, o):
        if self.t.dim() != 0:
            raise TypeError('division by a non-scalar must be ./')
        return V(_t(o) / self.t)

    def __mod__(self, o):       # .*
        return V(self.t * _t(o))

    def __floordiv__(self, o):  # ./
        return V(self.t / _t(o))

    def __rfloordiv__(self, o):
        return V(_t(o) / self.t)

    def __getitem__(self, k):
        if isinstance(k, slice):  # 1-based, inclusive
            return V(self.t[int(k.start) - 1:int(k.stop)])
        return V(self.t[int(k) - 1])

    def __setitem__(self, k, v):
        # element assignment builds a new tensor so that autograd sees it
        t = self.t.clone()
        t[int(k) - 1] = _t(v)
        self.t = t


def _fun():
    f = {}
    f['pi'] = lambda: math.pi
    f['square'] = lambda a: V(_t(a) ** 2)
    f['sqrt'] = lambda a: V(torch.sqrt(_t(a)))
    f['sum'] = lambda a: V(_t(a).sum())
    f['rep_vector'] = lambda a, n: V(torch.full((int(n),), float(a), dtype=torch.float64))
    f['append_row'] = lambda a, b: V(torch.cat((_t(a).reshape(-1), _t(b).reshape(-1))))
    return f


def _lpdf(dist, y, args):
    y = _t(y)
    if dist == 'std_normal':
        return (-0.5 * y ** 2 - 0.5 * math.log(2 * math.pi)).sum()
    if dist == 'normal':
        mu, sg = _t(args[0]), _t(args[1])
        return (-0.5 * ((y - mu) / sg) ** 2 - torch.log(sg) - 0.5 * math.log(2 * math.pi) + 0 * y).sum()
    if dist == 'inv_gamma':
        a, b = _t(args[0]), _t(args[1])
        return (a * torch.log(b) - torch.lgamma(a) - (a + 1) * torch.log(y) - b / y).sum()
    if dist == 'exponential':
        lam = _t(args[0])
        return (torch.log(lam) - lam * y).sum()
    raise NotImplementedError(dist)


_DECL = re.compile(r'^(int|real|vector|matrix)\s*(<[^>]*>)?\s*(\[[^\]]*\])?\s*(\w+)\s*(?:=\s*(.*))?$', re.S)


def _blocks(src):
    src = re.sub(r'//[^\n]*', '', src)
    out, i = {}, 0
    for m in re.finditer(r'(transformed data|transformed parameters|generated quantities|data|parameters|model)\s*\{', src):
        if m.start() < i:
            continue
        depth, j = 1, m.end()
        while depth:
            depth += {'{': 1, '}': -1}.get(src[j], 0)
            j += 1
        out[m.group(1)] = src[m.end():j - 1]
        i = j
    return out


def _statements(body):
    """split at ';' -- a `for (...)` header stays attached to its single statement"""
    return [s.strip() for s in body.split(';') if s.strip()]


def _py(expr):
    return expr.replace('.*', ' % ').replace('./', ' // ')


class Program:
    def __init__(self, source):
        self.blocks = _blocks(source)

    # -- helpers ------------------------------------------------------------------------------------------------------
    def _eval(self, expr, env):
        return eval(_py(expr), {'__builtins__': {}}, env)

    def _dims(self, dims, env):
        return [int(self._eval(d, env)) for d in dims.strip('[]').split(',')] if dims else []

    def _run(self, stmt, env, check):
        m = re.match(r'^for\s*\(\s*(\w+)\s+in\s+(.+?):(.+?)\)\s*(.*)$', stmt, re.S)
        if m:
            var, lo, hi, body = m.groups()
            for k in range(int(self._eval(lo, env)), int(self._eval(hi, env)) + 1):
                env[var] = I(k)
                self._run(body, env, check)
            return
        m = _DECL.match(stmt)
        if m:
            typ, cons, dims, name, init = m.groups()
            if init is not None:
                v = self._eval(init, env)
                env[name] = v if isinstance(v, (V, I)) else (I(v) if typ == 'int' else V(_t(v)))
            else:
                env[name] = V(torch.zeros(self._dims(dims, env), dtype=torch.float64))
            if cons and 'lower=0' in cons.replace(' ', ''):
                check.append(name)
            return
        m = re.match(r'^(\w+)\s*\[(.+)\]\s*=\s*(.*)$', stmt, re.S)
        if m:
            name, idx, rhs = m.groups()
            env[name][self._eval(idx, env)] = self._eval(rhs, env)
            return
        m = re.match(r'^(\w+)\s*=\s*(.*)$', stmt, re.S)
        if m:
            env[m.group(1)] = self._eval(m.group(2), env)
            return
        raise SyntaxError(f'statement outside the supported subset: {stmt!r}')

    # -- public -------------------------------------------------------------------------------------------------------
    def n_unconstrained(self, data):
        env = self._data_env(data)
        return sum(n for _, n, _ in self._params(env)[0])

    def _data_env(self, data):
        env = dict(_fun())
        for stmt in _statements(self.blocks['data']):
            typ, cons, dims, name, _ = _DECL.match(stmt).groups()
            v = data[name]
            env[name] = I(int(v)) if typ == 'int' else V(torch.as_tensor(v, dtype=torch.float64))
            if typ in ('vector', 'matrix'):
                assert list(env[name].t.shape) == self._dims(dims, env), (name, tuple(env[name].t.shape), dims)
        for stmt in _statements(self.blocks.get('transformed data', '')):
            self._run(stmt, env, [])
        return env

    def _params(self, env):
        """[(name, size, lower=0?)] in declaration order (= order of the unconstrained vector), {name: type}"""
        out, types = [], {}
        for stmt in _statements(self.blocks['parameters']):
            typ, cons, dims, name, _ = _DECL.match(stmt).groups()
            n = 1 if typ == 'real' else self._dims(dims, env)[0]
            out.append((name, n, bool(cons and 'lower=0' in cons.replace(' ', ''))))
            types[name] = typ
        return out, types

    def log_prob(self, data, u, jacobian, env_out=None):
        """log density at the unconstrained point u (torch float64 vector, Stan's declaration order).  ``env_out``: a
        dict that receives every parameter and transformed parameter by name (what Stan reports back)."""
        env = self._data_env(data)
        params, types = self._params(env)
        lp = torch.zeros((), dtype=torch.float64)
        o = 0
        for name, n, lower0 in params:
            seg = u[o:o + n]
            o += n
            if lower0:
                if jacobian:
                    lp = lp + seg.sum()
                seg = torch.exp(seg)
            env[name] = V(seg[0] if types[name] == 'real' else seg)
        assert o == u.numel()
        check = []
        for stmt in _statements(self.blocks.get('transformed parameters', '')):
            self._run(stmt, env, check)
        if env_out is not None:
            env_out.update({k: v.t.detach().numpy().copy() for k, v in env.items() if isinstance(v, V)})
        for name in check:  # Stan validates the declared bounds at the end of the block and rejects the point
            if bool((env[name].t < 0).any()):
                return torch.tensor(-math.inf, dtype=torch.float64)
        for stmt in _statements(self.blocks['model']):
            m = re.match(r'^(.+?)~\s*(\w+)\s*\((.*)\)$', stmt, re.S)
            if not m:
                raise SyntaxError(f'model statement outside the supported subset: {stmt!r}')
            y = self._eval(m.group(1).strip(), env)
            args = self._eval('(' + m.group(3) + ',)', env) if m.group(3).strip() else ()
            lp = lp + _lpdf(m.group(2), y, args)
        return lp
