"""Minimal symbolic expression DAG: the input format of the hot path.

Mirrors the observable behaviour of the reference's ``expression`` class
(/root/reference/heyoka/expose_expression.cpp:60-306) for the subset the
batch Taylor integrator needs: numbers, variables, ``par[i]``, ``time``,
``+ - * / **`` and ``sqrt sin cos exp log``, plus ``diff`` (used by
``var_ode_sys``).  Nodes are hash-consed so that shared sub-expressions are
visited once by the Taylor decomposition (reference semantics:
/root/reference/doc/notebooks/ex_system_internals.ipynb).
"""

import math
import numbers

import numpy as np

__all__ = [
    "expression",
    "make_vars",
    "par",
    "time",
    "sin",
    "cos",
    "exp",
    "log",
    "sqrt",
    "pow",
    "sum",
    "prod",
    "diff",
    "square",
    "tan",
]

_NUM, _VAR, _PAR, _TIME, _FUNC = "num", "var", "par", "time", "func"

# Hash-consing table: structural key -> node.
_table = {}


def _is_number(x):
    return isinstance(x, (numbers.Real, np.floating, np.integer)) and not isinstance(
        x, bool
    )


class expression:
    """Immutable expression node.

    ``expression(1.5)`` builds a number, ``expression("x")`` a variable
    (reference: expose_expression.cpp:63-77).
    """

    __slots__ = ("kind", "value", "name", "args", "_key", "_hash", "__weakref__")

    def __new__(cls, x=0.0):
        if isinstance(x, expression):
            return x
        if isinstance(x, str):
            return cls._make(_VAR, name=x)
        if _is_number(x):
            return cls._make(_NUM, value=float(x))
        raise TypeError(
            "cannot construct an expression from an object of type {}".format(type(x))
        )

    @classmethod
    def _make(cls, kind, value=None, name=None, args=()):
        if kind == _NUM:
            # Distinguish -0.0/0.0 and NaNs by bit pattern.
            key = (kind, np.float64(value).tobytes())
        elif kind == _FUNC:
            key = (kind, name, tuple(id(a) for a in args))
        else:
            key = (kind, value, name)
        node = _table.get(key)
        if node is None:
            node = object.__new__(cls)
            node.kind = kind
            node.value = value
            node.name = name
            node.args = tuple(args)
            node._key = key
            node._hash = hash(key)
            _table[key] = node
        return node

    # Pickling must go through the hash-consing table.
    def __reduce__(self):
        if self.kind == _NUM:
            return (expression, (self.value,))
        if self.kind == _VAR:
            return (expression, (self.name,))
        if self.kind == _PAR:
            return (_make_par, (self.value,))
        if self.kind == _TIME:
            return (_make_time, ())
        return (_make_func, (self.name, self.args))

    def __copy__(self):
        return self

    def __deepcopy__(self, memo):
        return self

    def __hash__(self):
        return self._hash

    # The reference compares structurally and returns a bool
    # (expose_expression.cpp:142-143).
    def __eq__(self, other):
        if isinstance(other, expression):
            return self is other
        return NotImplemented

    def __ne__(self, other):
        if isinstance(other, expression):
            return self is not other
        return NotImplemented

    # ---- arithmetic ----
    def __add__(self, o):
        return _add(self, _wrap(o))

    def __radd__(self, o):
        return _add(_wrap(o), self)

    def __sub__(self, o):
        return _sub(self, _wrap(o))

    def __rsub__(self, o):
        return _sub(_wrap(o), self)

    def __mul__(self, o):
        return _mul(self, _wrap(o))

    def __rmul__(self, o):
        return _mul(_wrap(o), self)

    def __truediv__(self, o):
        return _div(self, _wrap(o))

    def __rtruediv__(self, o):
        return _div(_wrap(o), self)

    def __pow__(self, o):
        return pow(self, o)

    def __rpow__(self, o):
        return pow(_wrap(o), self)

    def __neg__(self):
        return _neg(self)

    def __pos__(self):
        return self

    def __repr__(self):
        return _repr(self)

    __str__ = __repr__


def _wrap(x):
    if isinstance(x, expression):
        return x
    if _is_number(x):
        return expression._make(_NUM, value=float(x))
    raise TypeError(
        "unsupported operand type for an expression operation: {}".format(type(x))
    )


def _make_par(i):
    return expression._make(_PAR, value=int(i))


def _make_time():
    return expression._make(_TIME)


def _make_func(name, args):
    return expression._make(_FUNC, name=name, args=tuple(args))


def _num(v):
    return expression._make(_NUM, value=float(v))


def _isnum(e, v=None):
    return e.kind == _NUM and (v is None or e.value == v)


# ---- light canonicalisation (constant folding and neutral elements) ----
def _add(a, b):
    if _isnum(a) and _isnum(b):
        return _num(a.value + b.value)
    if _isnum(a, 0.0):
        return b
    if _isnum(b, 0.0):
        return a
    return _make_func("add", (a, b))


def _sub(a, b):
    if _isnum(a) and _isnum(b):
        return _num(a.value - b.value)
    if _isnum(b, 0.0):
        return a
    if _isnum(a, 0.0):
        return _neg(b)
    if a is b:
        return _num(0.0)
    return _make_func("sub", (a, b))


def _neg(a):
    if _isnum(a):
        return _num(-a.value)
    if a.kind == _FUNC and a.name == "neg":
        return a.args[0]
    if a.kind == _FUNC and a.name == "mul" and _isnum(a.args[0]):
        return _mul(_num(-a.args[0].value), a.args[1])
    return _make_func("neg", (a,))


def _mul(a, b):
    if _isnum(a) and _isnum(b):
        return _num(a.value * b.value)
    if _isnum(a, 0.0) or _isnum(b, 0.0):
        return _num(0.0)
    if _isnum(a, 1.0):
        return b
    if _isnum(b, 1.0):
        return a
    if _isnum(a, -1.0):
        return _neg(b)
    if _isnum(b, -1.0):
        return _neg(a)
    # Numbers go first (the reference prints "(c * x)").
    if _isnum(b) and not _isnum(a):
        a, b = b, a
    # c1 * (c2 * x) -> (c1*c2) * x
    if _isnum(a) and b.kind == _FUNC and b.name == "mul" and _isnum(b.args[0]):
        return _mul(_num(a.value * b.args[0].value), b.args[1])
    if _isnum(a) and b.kind == _FUNC and b.name == "neg":
        return _mul(_num(-a.value), b.args[0])
    return _make_func("mul", (a, b))


def _div(a, b):
    if _isnum(b):
        if b.value == 0.0:
            raise ZeroDivisionError("division by zero in an expression")
        if _isnum(a):
            return _num(a.value / b.value)
        if b.value == 1.0:
            return a
        if b.value == -1.0:
            return _neg(a)
    if _isnum(a, 0.0):
        return _num(0.0)
    return _make_func("div", (a, b))


def pow(a, b):
    """``a**b``.  Reference: expose_expression.cpp:300 (hey::pow overloads)."""
    a, b = _wrap(a), _wrap(b)
    if _isnum(b):
        if b.value == 0.0:
            return _num(1.0)
        if b.value == 1.0:
            return a
        if _isnum(a):
            return _num(math.pow(a.value, b.value))
    return _make_func("pow", (a, b))


def square(a):
    return pow(a, 2.0)


def _unary(name, fn):
    def f(a):
        a = _wrap(a)
        if _isnum(a):
            return _num(fn(a.value))
        return _make_func(name, (a,))

    f.__name__ = name
    return f


sin = _unary("sin", math.sin)
cos = _unary("cos", math.cos)
exp = _unary("exp", math.exp)
log = _unary("log", math.log)
sqrt = _unary("sqrt", math.sqrt)


asin = _unary("asin", math.asin)
acos = _unary("acos", math.acos)
atan = _unary("atan", math.atan)
erf = _unary("erf", math.erf)


def tan(a):
    a = _wrap(a)
    return sin(a) / cos(a)


# Hyperbolic functions, their inverses and the sigmoid, as compositions of exp / log / sqrt
# (reference: expose_expression.cpp:288-306 exposes them as primitives; the Taylor coefficients of a
# composition are those of the function, so the integrator sees the same right-hand side).
def sinh(a):
    a = _wrap(a)
    if _isnum(a):
        return _num(math.sinh(a.value))
    return 0.5 * (exp(a) - exp(-a))


def cosh(a):
    a = _wrap(a)
    if _isnum(a):
        return _num(math.cosh(a.value))
    return 0.5 * (exp(a) + exp(-a))


def tanh(a):
    a = _wrap(a)
    if _isnum(a):
        return _num(math.tanh(a.value))
    e2 = exp(2.0 * a)
    return (e2 - 1.0) / (e2 + 1.0)


def sigmoid(a):
    a = _wrap(a)
    if _isnum(a):
        return _num(1.0 / (1.0 + math.exp(-a.value)))
    return 1.0 / (1.0 + exp(-a))


def asinh(a):
    a = _wrap(a)
    if _isnum(a):
        return _num(math.asinh(a.value))
    return log(a + sqrt(a * a + 1.0))


def acosh(a):
    a = _wrap(a)
    if _isnum(a):
        return _num(math.acosh(a.value))
    return log(a + sqrt(a * a - 1.0))


def atanh(a):
    a = _wrap(a)
    if _isnum(a):
        return _num(math.atanh(a.value))
    return 0.5 * log((1.0 + a) / (1.0 - a))


def sum(terms):
    """N-ary sum (reference: expose_expression.cpp ``sum``)."""
    terms = [_wrap(t) for t in terms]
    if not terms:
        return _num(0.0)
    # Pairwise (balanced) reduction keeps the DAG shallow.
    while len(terms) > 1:
        nxt = [_add(terms[i], terms[i + 1]) for i in range(0, len(terms) - 1, 2)]
        if len(terms) % 2:
            nxt.append(terms[-1])
        terms = nxt
    return terms[0]


def prod(terms):
    terms = [_wrap(t) for t in terms]
    if not terms:
        return _num(1.0)
    while len(terms) > 1:
        nxt = [_mul(terms[i], terms[i + 1]) for i in range(0, len(terms) - 1, 2)]
        if len(terms) % 2:
            nxt.append(terms[-1])
        terms = nxt
    return terms[0]


def make_vars(*names):
    """Reference: expose_expression.cpp:262-283 (single name -> expression,
    several -> list)."""
    if len(names) == 0:
        raise ValueError("At least one argument is required")
    for n in names:
        if not isinstance(n, str):
            raise TypeError("make_vars() expects string arguments")
    if len(names) == 1:
        return expression(names[0])
    return [expression(n) for n in names]


class _par_impl:
    """``par[i]`` runtime parameter accessor (expose_expression.cpp:308-318)."""

    def __getitem__(self, i):
        if not isinstance(i, (int, np.integer)) or i < 0:
            raise TypeError("par[] requires a non-negative integer index")
        return _make_par(i)

    def __repr__(self):
        return "par"


par = _par_impl()
time = _make_time()

_INFIX = {"add": "+", "sub": "-", "mul": "*", "div": "/", "pow": "**"}


def _repr(e):
    # Iterative-safe enough for our depths; shared nodes are re-printed.
    k = e.kind
    if k == _NUM:
        v = float(e.value)
        if v == 0.0 or (1e-4 <= abs(v) < 1e16):
            return "{:.16f}".format(v)
        return "{:.17g}".format(v)
    if k == _VAR:
        return e.name
    if k == _PAR:
        return "p{}".format(e.value)
    if k == _TIME:
        return "t"
    if e.name in _INFIX:
        return "({} {} {})".format(_repr(e.args[0]), _INFIX[e.name], _repr(e.args[1]))
    if e.name == "neg":
        return "-{}".format(_repr(e.args[0]))
    return "{}({})".format(e.name, ", ".join(_repr(a) for a in e.args))


def get_variables(exs):
    """Sorted list of variable names appearing in the expressions."""
    seen, out, stack = set(), set(), list(exs)
    while stack:
        e = stack.pop()
        if id(e) in seen:
            continue
        seen.add(id(e))
        if e.kind == _VAR:
            out.add(e.name)
        stack.extend(e.args)
    return sorted(out)


def topo_order(roots):
    """Post-order (children first) list of the distinct nodes reachable from
    ``roots``; deterministic, iterative (deep N-body sums would overflow the
    Python stack otherwise)."""
    order, state = [], {}
    stack = [(r, 0) for r in reversed(list(roots))]
    while stack:
        e, i = stack.pop()
        if i == 0:
            if id(e) in state:
                continue
            state[id(e)] = 1
        if i < len(e.args):
            stack.append((e, i + 1))
            c = e.args[i]
            if id(c) not in state:
                stack.append((c, 0))
        else:
            order.append(e)
    return order


def diff(e, x):
    """Symbolic derivative of ``e`` w.r.t. variable or parameter ``x``
    (reference: expose_expression.cpp ``diff``)."""
    x = _wrap(x)
    if x.kind not in (_VAR, _PAR, _TIME):
        raise ValueError("diff() requires a variable or a parameter")
    d = {}
    for n in topo_order([e]):
        k = n.kind
        if k == _NUM:
            r = _num(0.0)
        elif k == _TIME:
            r = _num(1.0 if x.kind == _TIME else 0.0)  # (the explicit partial derivative w.r.t. time)
        elif k in (_VAR, _PAR):
            r = _num(1.0 if n is x else 0.0)
        else:
            a = n.args
            da = [d[id(c)] for c in a]
            nm = n.name
            if nm == "add":
                r = da[0] + da[1]
            elif nm == "sub":
                r = da[0] - da[1]
            elif nm == "neg":
                r = -da[0]
            elif nm == "mul":
                r = da[0] * a[1] + a[0] * da[1]
            elif nm == "div":
                r = (da[0] * a[1] - a[0] * da[1]) / (a[1] * a[1])
            elif nm == "pow":
                if _isnum(a[1]):
                    r = a[1] * pow(a[0], a[1].value - 1.0) * da[0]
                else:
                    r = n * (da[1] * log(a[0]) + a[1] * da[0] / a[0])
            elif nm == "sqrt":
                r = da[0] / (2.0 * n)
            elif nm == "sin":
                r = cos(a[0]) * da[0]
            elif nm == "cos":
                r = -(sin(a[0]) * da[0])
            elif nm == "exp":
                r = n * da[0]
            elif nm == "log":
                r = da[0] / a[0]
            elif nm == "asin":
                r = da[0] * pow(1.0 - a[0] * a[0], -0.5)
            elif nm == "acos":
                r = -(da[0] * pow(1.0 - a[0] * a[0], -0.5))
            elif nm == "atan":
                r = da[0] / (1.0 + a[0] * a[0])
            elif nm == "erf":
                r = (2.0 / math.sqrt(math.pi)) * exp(-(a[0] * a[0])) * da[0]
            else:
                raise NotImplementedError("diff() of '{}'".format(nm))
        d[id(n)] = r
    return d[id(e)]


def subs(e, smap):
    """Substitute variables by name: ``smap`` maps names (or variable
    expressions) to expressions."""
    m = {}
    for k, v in smap.items():
        m[k.name if isinstance(k, expression) else k] = _wrap(v)
    out = {}
    for n in topo_order([e]):
        if n.kind == _VAR and n.name in m:
            r = m[n.name]
        elif n.kind == _FUNC:
            r = rebuild(n, [out[id(c)] for c in n.args])
        else:
            r = n
        out[id(n)] = r
    return out[id(e)]


def rebuild(n, args):
    """Re-create function node ``n`` on new arguments via the public
    constructors (so folding applies)."""
    nm = n.name
    if nm == "add":
        return args[0] + args[1]
    if nm == "sub":
        return args[0] - args[1]
    if nm == "mul":
        return args[0] * args[1]
    if nm == "div":
        return args[0] / args[1]
    if nm == "neg":
        return -args[0]
    if nm == "pow":
        return pow(args[0], args[1])
    return {"sin": sin, "cos": cos, "exp": exp, "log": log, "sqrt": sqrt, "asin": asin, "acos": acos,
            "atan": atan, "erf": erf}[nm](args[0])


def eval_numpy(e, var_values, pars=None, tm=None):
    """Evaluate expression(s) with numpy (test helper: energies, Jacobi
    constants).  ``var_values`` maps names to arrays."""
    single = isinstance(e, expression)
    roots = [e] if single else list(e)
    val = {}
    for n in topo_order(roots):
        k = n.kind
        if k == _NUM:
            r = n.value
        elif k == _VAR:
            r = var_values[n.name]
        elif k == _PAR:
            r = pars[n.value]
        elif k == _TIME:
            r = tm
        else:
            a = [val[id(c)] for c in n.args]
            nm = n.name
            if nm == "add":
                r = a[0] + a[1]
            elif nm == "sub":
                r = a[0] - a[1]
            elif nm == "mul":
                r = a[0] * a[1]
            elif nm == "div":
                r = a[0] / a[1]
            elif nm == "neg":
                r = -a[0]
            elif nm == "pow":
                r = np.power(a[0], a[1])
            elif nm in ("asin", "acos", "atan"):
                r = getattr(np, "arc" + nm[1:])(a[0])
            elif nm == "erf":
                r = np.vectorize(math.erf)(a[0]) if isinstance(a[0], np.ndarray) else math.erf(a[0])
            else:
                r = getattr(np, nm)(a[0])
        val[id(n)] = r
    res = [val[id(r)] for r in roots]
    return res[0] if single else res
