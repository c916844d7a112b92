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
