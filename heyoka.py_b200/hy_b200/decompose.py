"""Taylor decomposition of an ODE system and its lowering to the opcode tape.

This is the host-side replacement for what the reference does inside the
``taylor_adaptive_batch`` constructor (/root/reference/heyoka/
expose_batch_integrators.cpp:166-208 -> [UPSTREAM] Taylor decomposition + LLVM
codegen): the right-hand sides (and event functions) are split into
u-variables, one per elementary operation, shared sub-expressions visited once
(/root/reference/doc/notebooks/ex_system_internals.ipynb).  Instead of emitting
LLVM IR we emit the ``hy_op``/``hy_term`` tape declared in include/hy_cuda.h.

Design points:
 * linear sub-trees (add/sub/neg/scaling by numbers or by one ``par[i]``) are
   collapsed into a single LINCOMB op - they cost one flop per order and need
   no history;
 * only u-variables whose *history* is read by a recurrence (operands of
   mul/square/div/pow/exp/log/sincos, self-referencing outputs, state
   variables, event functions) get a full jet; the rest own a single row;
 * ops are sorted by dependency level so that the G threads cooperating on one
   trajectory can run a level in parallel and synchronise between levels.
"""

import math

import numpy as np

from . import _expression as E

# Opcodes: keep in sync with include/hy_cuda.h.
OP_LINCOMB, OP_MUL, OP_SQUARE, OP_DIV, OP_POW, OP_SQRT, OP_EXP, OP_LOG = range(8)
OP_SINCOS, OP_TIME, OP_SVD, OP_SUMSQ, OP_MULSH, OP_ADDSUB, OP_INTG = 8, 9, 10, 11, 12, 13, 14
INTG_CODES = {"asin": 0, "acos": 1, "atan": 2, "erf": 3}
OPF_EVENT, OPF_NEGA, OPF_NEGB, OPF_SVD = 0x1, 0x2, 0x4, 0x8
REF_JET = 0x80000000
REF_ONE = 0x7FFFFFFF

OP_NAMES = {
    OP_LINCOMB: "lincomb",
    OP_MUL: "mul",
    OP_SQUARE: "square",
    OP_DIV: "div",
    OP_POW: "pow",
    OP_SQRT: "sqrt",
    OP_EXP: "exp",
    OP_LOG: "log",
    OP_SINCOS: "sincos",
    OP_TIME: "time",
    OP_SVD: "svd",
    OP_SUMSQ: "sumsq",
    OP_MULSH: "mulsh",
    OP_ADDSUB: "addsub",
    OP_INTG: "intg",
}

op_dtype = np.dtype(
    [
        ("opcode", "<u2"),
        ("flags", "<u2"),
        ("dst", "<u4"),
        ("dst2", "<u4"),
        ("a", "<u4"),
        ("b", "<u4"),
        ("n", "<u4"),
        ("imm", "<f8"),
    ],
    align=True,
)
term_dtype = np.dtype(
    [("src", "<u4"), ("par", "<i4"), ("coef", "<f8"), ("dst", "<u4"), ("pad", "<u4")],
    align=True,
)
assert op_dtype.itemsize == 32 and term_dtype.itemsize == 24

ONE = -1  # pseudo u-variable id of the constant jet [1, 0, 0, ...]


def taylor_order(tol):
    """Reference order selection: p = max(2, ceil(-ln(tol)/2 + 1))
    (verified on the reference notebooks' printed orders)."""
    return max(2, int(math.ceil(-math.log(tol) / 2.0 + 1.0)))


class _UVar:
    __slots__ = (
        "id", "op", "args", "imm", "terms", "jet", "level", "row", "pair", "event", "mterms",
        "inv_row", "svd_of", "signs",
    )

    def __init__(self, uid, op, args=(), imm=0.0, terms=None):
        self.id = uid
        self.op = op  # None for state variables, else opcode
        self.args = tuple(args)  # u-var ids
        self.imm = imm
        self.terms = terms  # LINCOMB: list of (uid|ONE, par, coef)
        self.jet = False
        self.level = 0
        self.row = None
        self.pair = None  # SINCOS: id of the partner output (cos)
        self.event = False
        self.inv_row = 0
        self.svd_of = None  # state index when fused with the state recurrence
        self.signs = 0


class _Lin:
    """Linear form sum_i coef_i * par_i * u_i, keyed by (uid, par)."""

    __slots__ = ("t",)

    def __init__(self, t=None):
        self.t = t if t is not None else {}

    @staticmethod
    def atom(uid):
        return _Lin({(uid, -1): 1.0})

    @staticmethod
    def const(v, par=-1):
        return _Lin({(ONE, par): float(v)})

    def is_const(self):
        return all(k[0] == ONE for k in self.t)

    def is_num(self):
        return all(k == (ONE, -1) for k in self.t)

    def num(self):
        return self.t.get((ONE, -1), 0.0)

    def single_atom(self):
        if len(self.t) == 1:
            (k, c), = self.t.items()
            if k[1] == -1 and c == 1.0 and k[0] != ONE:
                return k[0]
        return None

    def add(self, o, s=1.0):
        t = dict(self.t)
        for k, c in o.t.items():
            t[k] = t.get(k, 0.0) + s * c
        return _Lin({k: c for k, c in t.items() if c != 0.0})

    def scale(self, s):
        if s == 0.0:
            return _Lin()
        return _Lin({k: c * s for k, c in self.t.items()})

    def has_par(self):
        return any(k[1] >= 0 for k in self.t)

    def scale_par(self, p):
        return _Lin({(k[0], p): c for k, c in self.t.items()})

    def key(self):
        return tuple((k, np.float64(c).tobytes()) for k, c in self.t.items())


class Decomposition:
    """Result of :func:`decompose`."""

    def __init__(self):
        self.n_state = 0
        self.n_par = 0
        self.order = 0
        self.uvars = []
        self.ops = None
        self.terms = None
        self.level_start = None
        self.ev_ref = None
        self.n_rows = 0
        self.n_events = 0
        self.var_names = []

    # ---- cost model used by bench.py / DESIGN.md (FMA = 2 flops) ----
    def flops_per_step(self):
        p = self.order
        fl = 0
        lo = 0  # shared-memory operand loads of the convolutions (no blocking)
        for o in self.ops:
            oc = int(o["opcode"])
            n = int(o["n"])
            for k in range(p):
                if oc == OP_LINCOMB:
                    fl += 2 * n
                    lo += n
                elif oc == OP_ADDSUB:
                    fl += 1
                    lo += 2
                elif oc == OP_MUL:
                    fl += 2 * (k + 1)
                    lo += 2 * (k + 1)
                elif oc == OP_SQUARE:
                    fl += 2 * (k // 2 + 1) + 1
                    lo += 2 * (k // 2 + 1)
                elif oc == OP_SUMSQ:
                    fl += n * (2 * (k // 2 + 1) + 1)
                    lo += n * 2 * (k // 2 + 1)
                elif oc == OP_MULSH:
                    fl += n * 2 * (k + 1)
                    lo += (n + 1) * (k + 1)
                elif oc in (OP_DIV, OP_EXP, OP_LOG, OP_INTG):
                    fl += 2 * k + 2
                    lo += 2 * k + 1
                elif oc in (OP_POW, OP_SQRT):
                    fl += 4 * k + 3
                    lo += 2 * k + 1
                elif oc == OP_SINCOS:
                    fl += 2 * (3 * k) + 2
                    lo += 3 * k
                elif oc == OP_SVD:
                    fl += 1
                    lo += 1
        fl += 2 * self.n_state * p  # Horner
        fl += 4 * self.n_state + 40  # norms + step size
        return fl, lo


def _is_int(v):
    return float(v).is_integer()


def decompose(sys, order, events=(), fuse=True):
    """Decompose ``sys`` = [(var, rhs), ...] (+ event expressions) into a
    tape for Taylor order ``order``.

    ``events`` is the list of event expressions, terminal events first.
    """
    sys = list(sys)
    if len(sys) == 0:
        raise ValueError("Cannot integrate a system of zero equations")
    names = []
    for lhs, _ in sys:
        if not isinstance(lhs, E.expression) or lhs.kind != "var":
            raise ValueError(
                "The left-hand side of an ODE must be a variable, but it is '{}' instead".format(
                    lhs
                )
            )
        if lhs.name in names:
            raise ValueError(
                "Error in the Taylor decomposition: the variable '{}' appears twice in the "
                "left-hand sides".format(lhs.name)
            )
        names.append(lhs.name)
    rhs = [E._wrap(r) for _, r in sys]
    events = [E._wrap(e) for e in events]
    n = len(sys)
    for v in E.get_variables(rhs + events):
        if v not in names:
            raise ValueError(
                "The variable '{}' appears in the right-hand side of the system but not among "
                "the state variables".format(v)
            )

    uv = [_UVar(i, None) for i in range(n)]
    for u in uv:
        u.jet = True
    sv_index = {nm: i for i, nm in enumerate(names)}
    n_par = 0

    def new_u(op, args=(), imm=0.0, terms=None):
        u = _UVar(len(uv), op, args, imm, terms)
        uv.append(u)
        return u.id

    cse = {}  # structural key -> uid (or tuple of uids)
    time_u = [None]

    def materialise(lin):
        """Return a u-var id holding the value of the linear form."""
        a = lin.single_atom()
        if a is not None:
            return a
        k = ("lin", lin.key())
        if k not in cse:
            terms = [(key[0], key[1], c) for key, c in lin.t.items()]
            if not terms:
                terms = [(ONE, -1, 0.0)]
            cse[k] = new_u(OP_LINCOMB, [t[0] for t in terms if t[0] != ONE], terms=terms)
        return cse[k]

    def nonlin(op, args, imm=0.0):
        k = (op, tuple(args), np.float64(imm).tobytes())
        if k not in cse:
            cse[k] = new_u(op, args, imm)
        return cse[k]

    def sincos(arg):
        k = (OP_SINCOS, arg)
        if k not in cse:
            s = new_u(OP_SINCOS, (arg,))
            c = new_u(OP_SINCOS, (arg,))
            uv[s].pair = c
            uv[c].pair = s
            uv[c].op = "cos_of"  # marker: produced by the partner op
            cse[k] = (s, c)
        return cse[k]

    def do_pow(base_lin, alpha):
        if alpha == 0.0:
            return _Lin.const(1.0)
        if alpha == 1.0:
            return base_lin
        if base_lin.is_num():
            return _Lin.const(math.pow(base_lin.num(), alpha))
        b = materialise(base_lin)
        if alpha == 2.0:
            return _Lin.atom(nonlin(OP_SQUARE, (b,)))
        if alpha == 0.5:
            return _Lin.atom(nonlin(OP_SQRT, (b,)))
        if _is_int(alpha) and 3.0 <= alpha <= 16.0:
            # Small positive integer powers by repeated multiplication: the
            # pow recurrence divides by a[0] and would fail at a[0] = 0.
            e, acc, sq = int(alpha), None, b
            while e:
                if e & 1:
                    acc = sq if acc is None else nonlin(OP_MUL, (min(acc, sq), max(acc, sq)))
                e >>= 1
                if e:
                    sq = nonlin(OP_SQUARE, (sq,))
            return _Lin.atom(acc)
        return _Lin.atom(nonlin(OP_POW, (b,), alpha))

    lin_of = {}
    roots = rhs + events
    for node in E.topo_order(roots):
        k = node.kind
        if k == "num":
            L = _Lin.const(node.value)
        elif k == "var":
            L = _Lin.atom(sv_index[node.name])
        elif k == "par":
            n_par = max(n_par, node.value + 1)
            L = _Lin.const(1.0, node.value)
        elif k == "time":
            if time_u[0] is None:
                time_u[0] = new_u(OP_TIME)
            L = _Lin.atom(time_u[0])
        else:
            a = [lin_of[id(c)] for c in node.args]
            nm = node.name
            if nm == "add":
                L = a[0].add(a[1])
            elif nm == "sub":
                L = a[0].add(a[1], -1.0)
            elif nm == "neg":
                L = a[0].scale(-1.0)
            elif nm == "mul":
                x, y = a
                if x.is_num():
                    L = y.scale(x.num())
                elif y.is_num():
                    L = x.scale(y.num())
                elif x.is_const() and len(x.t) == 1 and not y.has_par():
                    ((_, p), c), = x.t.items()
                    L = y.scale(c).scale_par(p)
                elif y.is_const() and len(y.t) == 1 and not x.has_par():
                    ((_, p), c), = y.t.items()
                    L = x.scale(c).scale_par(p)
                else:
                    ux, uy = materialise(x), materialise(y)
                    if ux == uy:
                        L = _Lin.atom(nonlin(OP_SQUARE, (ux,)))
                    else:
                        # Commutative: canonical operand order for CSE.
                        lo_, hi_ = min(ux, uy), max(ux, uy)
                        L = _Lin.atom(nonlin(OP_MUL, (lo_, hi_)))
            elif nm == "div":
                x, y = a
                if y.is_num():
                    L = x.scale(1.0 / y.num())
                elif x.is_const():
                    # c / y  ->  c * y^-1
                    r = do_pow(y, -1.0)
                    if x.is_num():
                        L = r.scale(x.num())
                    elif len(x.t) == 1:
                        ((_, p), c), = x.t.items()
                        L = r.scale(c).scale_par(p)
                    else:
                        L = _Lin.atom(nonlin(OP_MUL, (materialise(x), materialise(r))))
                else:
                    L = _Lin.atom(nonlin(OP_DIV, (materialise(x), materialise(y))))
            elif nm == "pow":
                if a[1].is_num():
                    L = do_pow(a[0], a[1].num())
                else:
                    # a^b = exp(b log a)
                    lg = nonlin(OP_LOG, (materialise(a[0]),))
                    pr = nonlin(OP_MUL, (materialise(a[1]), lg))
                    L = _Lin.atom(nonlin(OP_EXP, (pr,)))
            elif nm == "sqrt":
                L = do_pow(a[0], 0.5)
            elif nm in ("exp", "log"):
                if a[0].is_num():
                    L = _Lin.const(getattr(math, nm)(a[0].num()))
                else:
                    L = _Lin.atom(
                        nonlin(OP_EXP if nm == "exp" else OP_LOG, (materialise(a[0]),))
                    )
            elif nm in ("sin", "cos"):
                if a[0].is_num():
                    L = _Lin.const(getattr(math, nm)(a[0].num()))
                else:
                    s, c = sincos(materialise(a[0]))
                    L = _Lin.atom(s if nm == "sin" else c)
            elif nm in INTG_CODES:
                # F(a) with dF/da = g(a) built from existing ops: F[k] = (1/k) sum_j j a[j] g[k-j]
                if a[0].is_num():
                    L = _Lin.const(getattr(math, nm)(a[0].num()))
                else:
                    ua = materialise(a[0])
                    sq = do_pow(_Lin.atom(ua), 2.0)
                    if nm in ("asin", "acos"):
                        g = do_pow(_Lin.const(1.0).add(sq, -1.0), -0.5)
                        if nm == "acos":
                            g = g.scale(-1.0)
                    elif nm == "atan":
                        g = do_pow(_Lin.const(1.0).add(sq), -1.0)
                    else:
                        g = _Lin.atom(nonlin(OP_EXP, (materialise(sq.scale(-1.0)),))).scale(2.0 / math.sqrt(math.pi))
                    L = _Lin.atom(nonlin(OP_INTG, (ua, materialise(g)), float(INTG_CODES[nm])))
            else:
                raise NotImplementedError(
                    "the function '{}' is not supported by the Taylor decomposition".format(nm)
                )
        lin_of[id(node)] = L

    # State derivatives and event functions must live in u-variables.
    sv_src = []
    for r in rhs:
        L = lin_of[id(r)]
        a = L.single_atom()
        if a is None:
            # Force a dedicated LINCOMB even for constants/params.
            a = materialise(L)
        sv_src.append(a)
    ev_u = []
    for e in events:
        L = lin_of[id(e)]
        a = L.single_atom()
        if a is None:
            a = materialise(L)
        elif a < n:
            # An event on a bare state variable reads the state jet directly.
            pass
        ev_u.append(a)

    # ---- optional fusion of register-reuse super-ops ----
    if fuse:
        _fuse(uv, n, sv_src, ev_u)

    # ---- which u-vars need full jets ----
    for u in uv:
        if u.op in (OP_DIV, OP_POW, OP_SQRT, OP_EXP, OP_LOG, OP_SINCOS, "cos_of", OP_TIME):
            u.jet = True
        if u.op in (OP_MUL, OP_SQUARE, OP_POW, OP_SQRT, OP_EXP, OP_LOG, OP_SINCOS, "cos_of", OP_INTG):
            for a in u.args:
                uv[a].jet = True
        if u.op == OP_DIV:
            uv[u.args[1]].jet = True
        if u.op == OP_SUMSQ:
            for (s, _, _) in u.terms:
                uv[s].jet = True
        if u.op == OP_MULSH:
            for a in u.args:
                uv[a].jet = True
    for a in ev_u:
        uv[a].jet = True

    # ---- events: mark everything the event functions depend on ----
    stack = list(ev_u)
    while stack:
        a = stack.pop()
        if a == ONE or uv[a].event:
            continue
        uv[a].event = True
        stack.extend(x for x in uv[a].args if x != ONE)
        if uv[a].terms:
            stack.extend(t[0] for t in uv[a].terms if t[0] != ONE)
        if uv[a].pair is not None:
            stack.append(uv[a].pair)
        if uv[a].op == "mulsh_out":
            stack.append(uv[a].imm)  # owner op id

    # ---- ADDSUB: two-term +-1 linear combinations need no term records ----
    if fuse:
        for u in uv[n:]:
            if (u.op == OP_LINCOMB and len(u.terms) == 2
                    and all(t[0] != ONE and t[1] == -1 and t[2] in (1.0, -1.0) for t in u.terms)):
                u.op = OP_ADDSUB
                u.args = (u.terms[0][0], u.terms[1][0])
                u.signs = (OPF_NEGA if u.terms[0][2] < 0 else 0) | (OPF_NEGB if u.terms[1][2] < 0 else 0)
                u.terms = None
        # ---- fuse "x_i' = <linear op>" into the linear op itself ----
        nuse = {}
        for u in uv:
            srcs = set(a for a in u.args if a != ONE)
            if u.terms:
                srcs |= set(t[0] for t in u.terms if t[0] != ONE)
            for a in srcs:
                nuse[a] = nuse.get(a, 0) + 1
        for a in list(sv_src) + list(ev_u):
            nuse[a] = nuse.get(a, 0) + 1
        for i, a in enumerate(sv_src):
            u = uv[a]
            if (a >= n and u.op in (OP_LINCOMB, OP_ADDSUB) and not u.jet and not u.event
                    and nuse.get(a, 0) == 1 and u.svd_of is None):
                u.svd_of = i

    # ---- levels ----
    def deps(u):
        d = [x for x in u.args if x != ONE]
        if u.terms:
            d += [t[0] for t in u.terms if t[0] != ONE]
        if u.op == "mulsh_out":
            d.append(u.imm)
        return d

    for u in uv[n:]:
        if u.op == "cos_of":
            continue
        u.level = 1 + max([uv[d].level for d in deps(u)] + [0])
        if u.pair is not None:
            uv[u.pair].level = u.level
        if u.op == OP_MULSH:
            for (_, _, _, d) in u.mterms:
                uv[d].level = u.level
    # mulsh outputs were created before their owner op: fix levels of
    # consumers by iterating to a fixed point (DAG depth is small).
    changed = True
    while changed:
        changed = False
        for u in uv[n:]:
            if u.op in ("cos_of", "mulsh_out"):
                continue
            lv = 1 + max([uv[d].level for d in deps(u)] + [0])
            if lv != u.level:
                u.level = lv
                changed = True
                if u.pair is not None:
                    uv[u.pair].level = lv
                if u.op == OP_MULSH:
                    for (_, _, _, d) in u.mterms:
                        uv[d].level = lv

    # ---- row allocation: state jets first, then jets, then cur rows ----
    P1 = order + 1
    row = 0
    for u in uv[:n]:
        u.row = row
        row += P1
    live = _live_set(uv, n, sv_src, ev_u)
    for u in uv[n:]:
        if u.id in live and u.jet and u.op != OP_MULSH:
            u.row = row
            row += P1
    for u in uv[n:]:
        if u.id in live and not u.jet and u.op != OP_MULSH and u.svd_of is None:
            u.row = row
            row += 1
    # Scratch rows holding 1/a[0] (computed once per step at order 0).
    for u in uv[n:]:
        if u.id in live and u.op in (OP_DIV, OP_POW, OP_SQRT, OP_LOG):
            u.inv_row = row
            row += 1
    n_rows = row

    def ref(uid):
        if uid == ONE:
            return REF_ONE
        u = uv[uid]
        return (u.row | REF_JET) if u.jet else u.row

    # ---- emit ops sorted by (level, opcode) ----
    emit = [
        u
        for u in uv[n:]
        if u.id in live and u.op not in ("cos_of", "mulsh_out")
    ]
    emit.sort(key=lambda u: (u.level, u.op, u.id))
    n_lev = (max([u.level for u in emit]) if emit else 0) + 1  # + SVD level
    fused_sv = set(u.svd_of for u in emit if u.svd_of is not None)
    sv_plain = [i for i in range(n) if i not in fused_sv]
    ops = np.zeros(len(emit) + len(sv_plain), dtype=op_dtype)
    terms = []
    level_start = [0]
    cur_level = 1
    for i, u in enumerate(emit):
        while u.level > cur_level:
            level_start.append(i)
            cur_level += 1
        o = ops[i]
        o["opcode"] = u.op
        o["flags"] = OPF_EVENT if u.event else 0
        if u.svd_of is not None:
            o["flags"] = int(o["flags"]) | OPF_SVD
            o["dst"] = ref(u.svd_of)
        elif u.op != OP_MULSH:
            o["dst"] = ref(u.id)
        o["imm"] = u.imm if u.op in (OP_POW, OP_INTG) else 0.0
        if u.op in (OP_LINCOMB, OP_SUMSQ):
            o["b"] = len(terms)
            o["n"] = len(u.terms)
            for (s, p, c) in u.terms:
                terms.append((ref(s), p, c, 0, 0))
        elif u.op == OP_MULSH:
            o["a"] = ref(u.args[0])  # the shared operand
            o["b"] = len(terms)
            o["n"] = len(u.mterms)
            o["dst"] = ref(u.mterms[0][3])
            for (s, p, c, d) in u.mterms:
                terms.append((ref(s), p, c, ref(d), 0))
        elif u.op == OP_SINCOS:
            o["a"] = ref(u.args[0])
            o["dst2"] = ref(u.pair)
        elif u.op == OP_TIME:
            pass
        elif u.op == OP_ADDSUB:
            o["flags"] = int(o["flags"]) | u.signs
            o["a"] = ref(u.args[0])
            o["b"] = ref(u.args[1])
        else:
            o["a"] = ref(u.args[0])
            if len(u.args) > 1:
                o["b"] = ref(u.args[1])
            if u.op in (OP_DIV, OP_POW, OP_SQRT, OP_LOG):
                o["dst2"] = u.inv_row
    while cur_level < n_lev:
        level_start.append(len(emit))
        cur_level += 1
    # SVD level: x_i[k+1] = src_i[k] / (k+1)
    for j, i in enumerate(sv_plain):
        o = ops[len(emit) + j]
        o["opcode"] = OP_SVD
        o["dst"] = ref(i)
        o["a"] = ref(sv_src[i])
    level_start.append(len(emit) + len(sv_plain))

    d = Decomposition()
    d.n_state = n
    d.n_par = n_par
    d.order = order
    d.uvars = uv
    d.ops = ops
    d.terms = (
        np.array(terms, dtype=term_dtype) if terms else np.zeros(0, dtype=term_dtype)
    )
    d.level_start = np.array(level_start, dtype=np.uint32)
    d.ev_ref = np.array([ref(a) for a in ev_u], dtype=np.uint32)
    d.n_rows = n_rows
    d.n_events = len(ev_u)
    d.var_names = names
    d.sv_src = sv_src
    return d


def _live_set(uv, n, sv_src, ev_u):
    """u-variables reachable from the state derivatives / events."""
    live, stack = set(), list(sv_src) + list(ev_u)
    while stack:
        a = stack.pop()
        if a == ONE or a in live:
            continue
        live.add(a)
        u = uv[a]
        stack.extend(x for x in u.args if x != ONE)
        if u.terms:
            stack.extend(t[0] for t in u.terms if t[0] != ONE)
        if u.pair is not None:
            stack.append(u.pair)
        if u.op == "mulsh_out":
            stack.append(u.imm)
        if u.op == OP_MULSH:
            stack.extend(t[3] for t in u.mterms)
    return live


def _fuse(uv, n, sv_src, ev_u):
    """Peephole fusion of generic register-reuse patterns.

    * SUMSQ : a LINCOMB whose terms are all ``+1 * square(a_i)`` and whose
              squares have no other consumer -> one op with a single
              accumulator (N-body: r^2 = dx^2 + dy^2 + dz^2).
    * MULSH : >= 2 MULs sharing one operand b -> one op that loads b[k-j]
              once per j for all products (N-body: dx*w, dy*w, dz*w).
    """
    uses = {}
    for u in uv:
        srcs = set(u.args)
        if u.terms:
            srcs |= set(t[0] for t in u.terms)
        for a in srcs:
            uses[a] = uses.get(a, 0) + 1
    for a in list(sv_src) + list(ev_u):
        uses[a] = uses.get(a, 0) + 1

    # SUMSQ
    for u in uv[n:]:
        if u.op != OP_LINCOMB or len(u.terms) < 2:
            continue
        ok = all(
            s != ONE and p == -1 and c == 1.0 and uv[s].op == OP_SQUARE and uses.get(s, 0) == 1
            for (s, p, c) in u.terms
        )
        if ok:
            srcs = [uv[s].args[0] for (s, _, _) in u.terms]
            u.op = OP_SUMSQ
            u.terms = [(s, -1, 1.0) for s in srcs]
            u.args = tuple(srcs)

    # MULSH: group MULs by shared operand (greedy, most-shared first).  Only products of the SAME
    # dependency depth are grouped: a product that feeds another one sharing the operand
    # ((s * d) * d in second-order variational equations) must not end up in the same op.
    depth = {}
    for u in uv:  # (u-variables are created in dependency order)
        srcs = [a for a in u.args if a != ONE]
        if u.terms:
            srcs += [t[0] for t in u.terms if t[0] != ONE]
        depth[u.id] = 1 + max([depth[a] for a in srcs] + [0]) if u.id >= n else 0
    muls = [u for u in uv[n:] if u.op == OP_MUL]
    by_operand = {}
    for u in muls:
        for a in set(u.args):
            by_operand.setdefault((a, depth[u.id]), []).append(u)
    taken = set()
    for (b, _), lst in sorted(by_operand.items(), key=lambda kv: (-len(kv[1]), kv[0])):
        grp = [u for u in lst if u.id not in taken]
        # Bound the register footprint of the fused op.
        while len(grp) >= 2:
            chunk, grp = grp[:4], grp[4:]
            if len(chunk) < 2:
                break
            owner = chunk[0]
            mterms = []
            for u in chunk:
                other = u.args[0] if u.args[1] == b else u.args[1]
                mterms.append((other, -1, 1.0, u.id))
                taken.add(u.id)
            # The first MUL's u-var stays the first output; a new u-var owns the op.
            op_u = _UVar(len(uv), OP_MULSH, (b,) + tuple(t[0] for t in mterms))
            op_u.mterms = mterms
            uv.append(op_u)
            for u in chunk:
                u.op = "mulsh_out"
                u.imm = op_u.id
                u.args = ()


# --------------------------------------------------------------------------------------
# Event tape: the event functions ALONE, as functions of the state jets.
# --------------------------------------------------------------------------------------
class EventTape:
    """Second tape for the register-resident kernels (csrc/hy_evtape.cuh): those keep the jets
    of the ODE's sub-expressions in registers, so event functions cannot share u-variables with
    the ODE; they are decomposed on their own, one event at a time (no sub-expression is shared
    between two events: event e is evaluated by lane e mod G without synchronisation).

    Row references: a state variable i is ``(i * (order + 1)) | REF_JET`` as in the main tape;
    every other row lives in the "event workspace", numbered from ``n_state * (order + 1)``.
    """

    def __init__(self):
        self.ops = np.zeros(0, dtype=op_dtype)
        self.terms = np.zeros(0, dtype=term_dtype)
        self.ev_ref = np.zeros(0, dtype=np.uint32)
        self.op_start = np.zeros(1, dtype=np.uint32)
        self.n_rows = 0       # rows of the event workspace
        self.n_par = 0
        self.n_events = 0


def decompose_event_tape(events, names, order):
    """Lower ``events`` (expressions in the state variables ``names``) to an :class:`EventTape`,
    or return None if they use runtime parameters (the register-resident kernels hold none)."""
    n, P1 = len(names), order + 1
    base0 = n * P1
    dummy = [(E.expression(nm), E.expression(0.0)) for nm in names]
    ops_all, terms_all, ev_ref, op_start = [], [], [], [0]
    next_row = base0
    for ev in events:
        dc = decompose(dummy, order, events=[ev])
        if dc.n_par:
            return None
        rowmap = {}

        def remap(ref, jet=None):
            nonlocal next_row
            ref = int(ref)
            if ref == REF_ONE:
                return REF_ONE
            b = ref & 0x7FFFFFFF
            is_jet = bool(ref & REF_JET) if jet is None else jet
            if b < base0:
                return ref  # a state jet
            if b not in rowmap:
                rowmap[b] = next_row
                next_row += P1 if is_jet else 1
            return rowmap[b] | (REF_JET if is_jet else 0)

        for o in dc.ops:
            oc, fl = int(o["opcode"]), int(o["flags"])
            if not (fl & OPF_EVENT) or (fl & OPF_SVD) or oc == OP_SVD:
                continue
            if oc == OP_INTG:
                return None  # (not in the event evaluator of the register-resident kernels: shared tape)
            q = np.zeros(1, dtype=op_dtype)[0]
            q["opcode"], q["flags"], q["n"], q["imm"] = oc, fl & (OPF_NEGA | OPF_NEGB), o["n"], o["imm"]
            if oc in (OP_LINCOMB, OP_SUMSQ, OP_MULSH):
                q["b"] = len(terms_all)
                for t in dc.terms[int(o["b"]): int(o["b"]) + int(o["n"])]:
                    if int(t["par"]) >= 0:
                        return None
                    terms_all.append((remap(t["src"]), -1, float(t["coef"]),
                                      remap(t["dst"]) if oc == OP_MULSH else 0, 0))
                if oc == OP_MULSH:
                    q["a"] = remap(o["a"])
                else:
                    q["dst"] = remap(o["dst"])
            else:
                q["dst"] = remap(o["dst"])
                if oc != OP_TIME:
                    q["a"] = remap(o["a"])
                if oc in (OP_MUL, OP_DIV, OP_ADDSUB):
                    q["b"] = remap(o["b"])
                if oc == OP_SINCOS:
                    q["dst2"] = remap(o["dst2"])
                elif oc in (OP_DIV, OP_POW, OP_SQRT, OP_LOG):
                    q["dst2"] = remap(int(o["dst2"]) & 0x7FFFFFFF, jet=False)
            ops_all.append(q)
        r = int(dc.ev_ref[0])
        if (r & 0x7FFFFFFF) < base0:
            # an event on a bare state variable: copy the jet into the workspace so that the root
            # finder sees a unit-stride polynomial like any other
            q = np.zeros(1, dtype=op_dtype)[0]
            q["opcode"], q["n"], q["b"] = OP_LINCOMB, 1, len(terms_all)
            terms_all.append((r, -1, 1.0, 0, 0))
            dst = next_row | REF_JET
            next_row += P1
            q["dst"] = dst
            ops_all.append(q)
            ev_ref.append(dst)
        else:
            ev_ref.append(remap(r, jet=True))
        op_start.append(len(ops_all))
    et = EventTape()
    et.ops = np.array(ops_all, dtype=op_dtype) if ops_all else np.zeros(0, dtype=op_dtype)
    et.terms = np.array(terms_all, dtype=term_dtype) if terms_all else np.zeros(0, dtype=term_dtype)
    et.ev_ref = np.array(ev_ref, dtype=np.uint32)
    et.op_start = np.array(op_start, dtype=np.uint32)
    et.n_rows = next_row - base0
    et.n_events = len(events)
    return et
