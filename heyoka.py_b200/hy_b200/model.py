"""ODE builders needed by the hot path's configurations.

Mirrors the call surface of the reference's ``heyoka.model`` for the three
models BASELINE.json's configs use (/root/reference/heyoka/expose_models.cpp:
237-272 ``nbody``/``nbody_energy``, :395-400 ``cr3bp``/``cr3bp_jacobi``,
``pendulum`` in the same file; structure pinned by
/root/reference/heyoka/_test_model.py:153-278).
"""

from . import _expression as E

__all__ = [
    "pendulum",
    "pendulum_energy",
    "nbody",
    "nbody_energy",
    "cr3bp",
    "cr3bp_jacobi",
    "np1body",
    "np1body_energy",
    "np1body_potential",
    "fixed_centres",
    "fixed_centres_energy",
    "fixed_centres_potential",
]


def pendulum(gconst=1.0, length=1.0):
    """x' = v, v' = -(g/l) sin x  (_test_model.py:213-233)."""
    x, v = E.make_vars("x", "v")
    return [(x, v), (v, -(gconst / length) * E.sin(x))]


def pendulum_energy(gconst=1.0, length=1.0):
    """Energy per unit mass (_test_model.py:235-259)."""
    x, v = E.make_vars("x", "v")
    return (0.5 * length * length) * v**2 + (gconst * length) * (1.0 - E.cos(x))


def _nbody_vars(n):
    xs = []
    for i in range(n):
        xs.append(
            E.make_vars(
                "x_{}".format(i),
                "y_{}".format(i),
                "z_{}".format(i),
                "vx_{}".format(i),
                "vy_{}".format(i),
                "vz_{}".format(i),
            )
        )
    return xs


def nbody(n, masses=None, Gconst=1.0):
    """Newtonian N-body problem in Cartesian coordinates; state variables
    ordered (x, y, z, vx, vy, vz) per body (_test_model.py:153-180).

    For every pair i<j: d = r_j - r_i, w = (d.d)^(-3/2); body i gains
    G m_j w d, body j loses G m_i w d.
    """
    if n < 2:
        raise ValueError("At least 2 bodies are needed to construct an N-body system")
    if masses is None:
        masses = [1.0] * n
    # masses are numbers or expressions (the reference accepts e.g. par[i]: runtime masses)
    masses = [m if isinstance(m, E.expression) else float(m) for m in masses]
    if len(masses) > n:
        raise ValueError("Too many masses for an N-body system")
    masses = masses + [0.0] * (n - len(masses))
    G = float(Gconst)
    b = _nbody_vars(n)
    acc = [[[] for _ in range(3)] for _ in range(n)]

    def _zero(m):
        return not isinstance(m, E.expression) and m == 0.0

    for i in range(n):
        for j in range(i + 1, n):
            if _zero(masses[i]) and _zero(masses[j]):
                continue
            d = [b[j][c] - b[i][c] for c in range(3)]
            w = E.pow(d[0] * d[0] + d[1] * d[1] + d[2] * d[2], -1.5)
            for c in range(3):
                f = d[c] * w
                if not _zero(masses[j]):
                    acc[i][c].append((G * masses[j]) * f)
                if not _zero(masses[i]):
                    acc[j][c].append((-G * masses[i]) * f)
    sys = []
    for i in range(n):
        for c in range(3):
            sys.append((b[i][c], b[i][3 + c]))
        for c in range(3):
            sys.append((b[i][3 + c], E.sum(acc[i][c]) if acc[i][c] else E.expression(0.0)))
    return sys


def nbody_energy(n, masses=None, Gconst=1.0):
    if masses is None:
        masses = [1.0] * n
    masses = [float(m) for m in masses] + [0.0] * (n - len(masses))
    G = float(Gconst)
    b = _nbody_vars(n)
    terms = []
    for i in range(n):
        if masses[i] != 0.0:
            terms.append(
                (0.5 * masses[i]) * (b[i][3] * b[i][3] + b[i][4] * b[i][4] + b[i][5] * b[i][5])
            )
    for i in range(n):
        for j in range(i + 1, n):
            if masses[i] * masses[j] == 0.0:
                continue
            d = [b[j][c] - b[i][c] for c in range(3)]
            terms.append(
                (-G * masses[i] * masses[j])
                * E.pow(d[0] * d[0] + d[1] * d[1] + d[2] * d[2], -0.5)
            )
    return E.sum(terms) if terms else E.expression(0.0)


def cr3bp(mu=1e-3):
    """Circular restricted three-body problem in the rotating frame,
    Hamiltonian form with momenta px = vx - y, py = vy + x
    (_test_model.py:261-266: dyn[0] == (x, px + y);
    doc/notebooks/The restricted three-body problem.ipynb:44-62)."""
    mu = float(mu)
    if not (0.0 < mu < 0.5):
        raise ValueError(
            "The 'mu' parameter in a CR3BP must be in the range (0, 0.5), but a value of {} "
            "was provided instead".format(mu)
        )
    x, y, z, px, py, pz = E.make_vars("x", "y", "z", "px", "py", "pz")
    xa = x - mu
    xb = x - mu + 1.0
    yz2 = y * y + z * z
    g1 = (1.0 - mu) * E.pow(xa * xa + yz2, -1.5)
    g2 = mu * E.pow(xb * xb + yz2, -1.5)
    g = g1 + g2
    return [
        (x, px + y),
        (y, py - x),
        (z, pz),
        (px, py - g1 * xa - g2 * xb),
        (py, -px - g * y),
        (pz, -g * z),
    ]


def cr3bp_jacobi(mu=1e-3):
    """Jacobi constant C = -2 H in the variables of :func:`cr3bp`."""
    mu = float(mu)
    x, y, z, px, py, pz = E.make_vars("x", "y", "z", "px", "py", "pz")
    xa = x - mu
    xb = x - mu + 1.0
    yz2 = y * y + z * z
    r1 = E.sqrt(xa * xa + yz2)
    r2 = E.sqrt(xb * xb + yz2)
    vx, vy = px + y, py - x
    kin = 0.5 * (vx * vx + vy * vy + pz * pz)
    eff = 0.5 * (x * x + y * y) + (1.0 - mu) / r1 + mu / r2
    return 2.0 * eff - 2.0 * kin


# ---------------------------------------------------------------------------------------------
# (N+1)-body problem: the N-body problem in the coordinates RELATIVE to body 0
# (reference: expose_models.cpp:256-270; structure pinned by _test_model.py:186-208 - n counts
# the central body, the variables are x_1 ... vz_{n-1}).
#   r_i'' = -G (m_0 + m_i) r_i / |r_i|^3 - sum_{j != i} G m_j [ (r_i - r_j) / |r_i - r_j|^3 + r_j / |r_j|^3 ]
# ---------------------------------------------------------------------------------------------
def _np1_masses(n, masses):
    if masses is None:
        masses = [1.0] * n
    masses = [m if isinstance(m, E.expression) else float(m) for m in masses]
    if len(masses) > n:
        raise ValueError("Too many masses for an (N+1)-body system")
    return masses + [0.0] * (n - len(masses))


def _is_zero(m):
    return not isinstance(m, E.expression) and m == 0.0


def np1body(n, masses=None, Gconst=1.0):
    if n < 2:
        raise ValueError("At least 2 bodies are needed to construct an (N+1)-body system")
    m = _np1_masses(n, masses)
    G = float(Gconst)
    b = _nbody_vars(n)[1:]  # bodies 1 .. n-1 (named x_1, ...)
    nb = n - 1
    rm3 = [E.pow(b[i][0] * b[i][0] + b[i][1] * b[i][1] + b[i][2] * b[i][2], -1.5) for i in range(nb)]
    acc = [[[] for _ in range(3)] for _ in range(nb)]
    for i in range(nb):
        mu = m[0] + m[i + 1] if not (_is_zero(m[0]) and _is_zero(m[i + 1])) else 0.0
        if not _is_zero(mu):
            for c in range(3):
                acc[i][c].append((-G * mu) * (b[i][c] * rm3[i]))
    for i in range(nb):
        for j in range(i + 1, nb):
            d = [b[i][c] - b[j][c] for c in range(3)]
            w = E.pow(d[0] * d[0] + d[1] * d[1] + d[2] * d[2], -1.5)
            for c in range(3):
                f = d[c] * w
                if not _is_zero(m[j + 1]):   # body j on body i: direct + indirect term
                    acc[i][c].append((-G * m[j + 1]) * f)
                    acc[i][c].append((-G * m[j + 1]) * (b[j][c] * rm3[j]))
                if not _is_zero(m[i + 1]):
                    acc[j][c].append((G * m[i + 1]) * f)
                    acc[j][c].append((-G * m[i + 1]) * (b[i][c] * rm3[i]))
    sys = []
    for i in range(nb):
        for c in range(3):
            sys.append((b[i][c], b[i][3 + c]))
        for c in range(3):
            sys.append((b[i][3 + c], E.sum(acc[i][c]) if acc[i][c] else E.expression(0.0)))
    return sys


def np1body_potential(n, masses=None, Gconst=1.0):
    m = _np1_masses(n, masses)
    G = float(Gconst)
    b = _nbody_vars(n)[1:]
    nb = n - 1
    terms = []
    for i in range(nb):
        if not (_is_zero(m[0]) or _is_zero(m[i + 1])):
            terms.append((-G * m[0] * m[i + 1]) * E.pow(b[i][0] * b[i][0] + b[i][1] * b[i][1] + b[i][2] * b[i][2], -0.5))
        for j in range(i + 1, nb):
            if _is_zero(m[i + 1]) or _is_zero(m[j + 1]):
                continue
            d = [b[i][c] - b[j][c] for c in range(3)]
            terms.append((-G * m[i + 1] * m[j + 1]) * E.pow(d[0] * d[0] + d[1] * d[1] + d[2] * d[2], -0.5))
    return E.sum(terms) if terms else E.expression(0.0)


def np1body_energy(n, masses=None, Gconst=1.0):
    """Total energy in the barycentric frame from the relative state: u_0 = -sum m_i v_i / M is the
    barycentric velocity of body 0, u_i = v_i + u_0."""
    m = _np1_masses(n, masses)
    b = _nbody_vars(n)[1:]
    nb = n - 1
    if all(_is_zero(x) for x in m):
        return E.expression(0.0)
    M = m[0]
    for x in m[1:]:
        M = M + x
    u0 = [E.sum([(-1.0 * m[i + 1]) * b[i][3 + c] for i in range(nb) if not _is_zero(m[i + 1])] or [E.expression(0.0)]) / M
          for c in range(3)]
    terms = []
    if not _is_zero(m[0]):
        terms.append((0.5 * m[0]) * (u0[0] * u0[0] + u0[1] * u0[1] + u0[2] * u0[2]))
    for i in range(nb):
        if _is_zero(m[i + 1]):
            continue
        u = [b[i][3 + c] + u0[c] for c in range(3)]
        terms.append((0.5 * m[i + 1]) * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]))
    kin = E.sum(terms) if terms else E.expression(0.0)
    return kin + np1body_potential(n, masses, Gconst)


# ---------------------------------------------------------------------------------------------
# Fixed centres: a test particle attracted by masses at fixed positions
# (reference: expose_models.cpp:174-232, :288-305; _test_model.py:101-151).
# ---------------------------------------------------------------------------------------------
def _fc_args(masses, positions):
    import numpy as np

    try:
        pos = np.array(positions, dtype=object)
    except Exception:
        pos = None
    if pos is None or pos.ndim != 2:
        raise ValueError(
            "Invalid positions array in a fixed centres model: the number of dimensions must be 2, "
            "but it is {} instead".format(0 if pos is None else pos.ndim))
    if pos.shape[1] != 3:
        raise ValueError(
            "Invalid positions array in a fixed centres model: the number of columns must be 3, "
            "but it is {} instead".format(pos.shape[1]))
    out = []
    for row in pos:
        r = []
        for v in row:
            if isinstance(v, E.expression):
                r.append(v)
            else:
                try:
                    r.append(E.expression(float(v)))
                except Exception:
                    raise TypeError(
                        "The positions array in a fixed centres model could not be converted into an array "
                        "of expressions - please make sure that the array's values can be converted into "
                        "heyoka expressions")
        out.append(r)
    masses = [mm if isinstance(mm, E.expression) else float(mm) for mm in masses]
    if len(masses) != len(out):
        raise ValueError(
            "Mismatched sizes detected in a fixed centres model: the number of masses ({}) differs from "
            "the number of position vectors ({})".format(len(masses), len(out)))
    return masses, out


def fixed_centres(Gconst=1.0, masses=(), positions=()):
    masses, pos = _fc_args(masses, positions)
    G = float(Gconst)
    x, y, z, vx, vy, vz = E.make_vars("x", "y", "z", "vx", "vy", "vz")
    r = (x, y, z)
    acc = [[], [], []]
    for mj, pj in zip(masses, pos):
        d = [r[c] - pj[c] for c in range(3)]
        w = E.pow(d[0] * d[0] + d[1] * d[1] + d[2] * d[2], -1.5)
        for c in range(3):
            acc[c].append((-G * mj) * (d[c] * w))
    return [(x, vx), (y, vy), (z, vz)] + [
        (v, E.sum(acc[c]) if acc[c] else E.expression(0.0)) for c, v in enumerate((vx, vy, vz))]


def fixed_centres_potential(Gconst=1.0, masses=(), positions=()):
    masses, pos = _fc_args(masses, positions)
    G = float(Gconst)
    x, y, z = E.make_vars("x", "y", "z")
    r = (x, y, z)
    terms = []
    for mj, pj in zip(masses, pos):
        d = [r[c] - pj[c] for c in range(3)]
        terms.append((-G * mj) * E.pow(d[0] * d[0] + d[1] * d[1] + d[2] * d[2], -0.5))
    return E.sum(terms) if terms else E.expression(0.0)


def fixed_centres_energy(Gconst=1.0, masses=(), positions=()):
    vx, vy, vz = E.make_vars("vx", "vy", "vz")
    return 0.5 * (vx * vx + vy * vy + vz * vz) + fixed_centres_potential(Gconst, masses, positions)
