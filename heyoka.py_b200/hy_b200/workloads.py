"""Synthetic workloads of BASELINE.json's configs (systems + seeded initial
conditions), shared by bench.py and the tests."""

import numpy as np

from . import model
from ._expression import make_vars, sin, cos, par, time
from . import _expression as _E


class _hy:  # tiny namespace so the builders below read like user code
    model = model
    make_vars = staticmethod(make_vars)
    sin = staticmethod(sin)
    cos = staticmethod(cos)
    par = par
    time = time


hy = _hy

# Outer Solar System (doc/notebooks/Outer Solar System.ipynb:42-113).
OSS_MASSES = np.array(
    [1.00000597682, 1 / 1047.355, 1 / 3501.6, 1 / 22869.0, 1 / 19314.0, 7.4074074e-09]
)
OSS_G = 0.01720209895 * 0.01720209895 * 365 * 365
OSS_IC = np.array([
    -4.06428567034226e-3, -6.08813756435987e-3, -1.66162304225834e-6,
    +6.69048890636161e-6 * 365, -6.33922479583593e-6 * 365, -3.13202145590767e-9 * 365,
    +3.40546614227466e0, +3.62978190075864e0, +3.42386261766577e-2,
    -5.59797969310664e-3 * 365, +5.51815399480116e-3 * 365, -2.66711392865591e-6 * 365,
    +6.60801554403466e0, +6.38084674585064e0, -1.36145963724542e-1,
    -4.17354020307064e-3 * 365, +3.99723751748116e-3 * 365, +1.67206320571441e-5 * 365,
    +1.11636331405597e1, +1.60373479057256e1, +3.61783279369958e-1,
    -3.25884806151064e-3 * 365, +2.06438412905916e-3 * 365, -2.17699042180559e-5 * 365,
    -3.01777243405203e1, +1.91155314998064e0, -1.53887595621042e-1,
    -2.17471785045538e-4 * 365, -3.11361111025884e-3 * 365, +3.58344705491441e-5 * 365,
    -2.13858977531573e1, +3.20719104739886e1, +2.49245689556096e0,
    -1.76936577252484e-3 * 365, -2.06720938381724e-3 * 365, +6.58091931493844e-4 * 365,
])


def oss_sys():
    return hy.model.nbody(6, masses=OSS_MASSES, Gconst=OSS_G)


def oss_ensemble(B, seed=20251019, amp=1e-12):
    """Config 2 ICs: base IC x (1 + U(-amp, amp)), recentred on the centre of
    mass as the notebook's gen() does (Outer Solar System.ipynb:200-252)."""
    rng = np.random.default_rng(seed)
    st = OSS_IC[:, None] * (1.0 + rng.uniform(-amp, amp, (36, B)))
    m = OSS_MASSES[:, None]
    for c in range(6):
        com = np.sum(st[c::6] * m, axis=0) / np.sum(m)
        st[c::6] -= com
    return np.ascontiguousarray(st)


def oss_energy(st):
    m = OSS_MASSES
    pos = np.stack([st[0::6], st[1::6], st[2::6]], axis=1)  # [body, 3, B]
    vel = np.stack([st[3::6], st[4::6], st[5::6]], axis=1)
    kin = 0.5 * np.sum(m[:, None] * np.sum(vel * vel, axis=1), axis=0)
    pot = 0.0
    for i in range(6):
        for j in range(i + 1, 6):
            r = np.sqrt(np.sum((pos[i] - pos[j]) ** 2, axis=0))
            pot = pot - OSS_G * m[i] * m[j] / r
    return kin + pot


def pendulum_sys():
    x, v = hy.make_vars("x", "v")
    return [(x, v), (v, -9.8 * hy.sin(x))]


PEND_IC = np.array([[0.0, 0.1, 0.2, 0.3], [0.25, 0.26, 0.27, 0.28]])


def cr3bp_sys(mu=0.01):
    return hy.model.cr3bp(mu=mu)


CR3BP_IC = np.array([-0.45, 0.80, 0.00, -0.80, -0.45, 0.58])


def cr3bp_ensemble(B, seed=20251020, amp=1e-3):
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(CR3BP_IC[:, None] + rng.uniform(-amp, amp, (6, B)))


def forced_pendulum_sys():
    x, v = hy.make_vars("x", "v")
    return [(x, v), (v, hy.cos(hy.time) - hy.par[0] * v - hy.sin(x))]


# Kepler + J2 single satellite (config 4): RHS as in
# doc/notebooks/Box control for Formation Flying Satellites.ipynb:179-195.
J2_MU = 398600.4418   # km^3/s^2
J2_J2 = 1082.645e-6
J2_RE = 6371.0        # km


def kepler_j2_sys():
    from ._expression import sqrt

    x, y, z, vx, vy, vz = make_vars("x", "y", "z", "vx", "vy", "vz")
    c = 1.5 * J2_J2 * J2_MU * J2_RE**2
    r2 = x**2 + y**2 + z**2
    ax = (-J2_MU * x / r2 - c * x / r2**2 * (1.0 - 5.0 * z**2 / r2)) / sqrt(r2)
    ay = (-J2_MU * y / r2 - c * y / r2**2 * (1.0 - 5.0 * z**2 / r2)) / sqrt(r2)
    az = (-J2_MU * z / r2 - c * z / r2**2 * (3.0 - 5.0 * z**2 / r2)) / sqrt(r2)
    return [(x, vx), (y, vy), (z, vz), (vx, ax), (vy, ay), (vz, az)]


def kepler_j2_ensemble(B, seed=20251021):
    """LEO initial conditions: a in U(6800, 7800) km, e in U(0, 0.02), i in U(0, pi),
    Omega, omega, M in U(0, 2 pi) -> Cartesian state [6, B]."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(6800.0, 7800.0, B)
    e = rng.uniform(0.0, 0.02, B)
    inc = rng.uniform(0.0, np.pi, B)
    Om, om, M = (rng.uniform(0.0, 2 * np.pi, B) for _ in range(3))
    E = M.copy()
    for _ in range(20):
        E = E - (E - e * np.sin(E) - M) / (1 - e * np.cos(E))
    xp = a * (np.cos(E) - e)
    yp = a * np.sqrt(1 - e * e) * np.sin(E)
    r = a * (1 - e * np.cos(E))
    vxp = -np.sqrt(J2_MU * a) / r * np.sin(E)
    vyp = np.sqrt(J2_MU * a) / r * np.sqrt(1 - e * e) * np.cos(E)
    cO, sO, co, so, ci, si = np.cos(Om), np.sin(Om), np.cos(om), np.sin(om), np.cos(inc), np.sin(inc)
    R11, R12 = cO * co - sO * so * ci, -cO * so - sO * co * ci
    R21, R22 = sO * co + cO * so * ci, -sO * so + cO * co * ci
    R31, R32 = so * si, co * si
    st = np.stack([R11 * xp + R12 * yp, R21 * xp + R22 * yp, R31 * xp + R32 * yp,
                   R11 * vxp + R12 * vyp, R21 * vxp + R22 * vyp, R31 * vxp + R32 * vyp])
    return np.ascontiguousarray(st)


def kepler_j2_energy(st):
    x, y, z, vx, vy, vz = st
    r2 = x * x + y * y + z * z
    r = np.sqrt(r2)
    c = 0.5 * J2_J2 * J2_MU * J2_RE**2
    return 0.5 * (vx * vx + vy * vy + vz * vz) - J2_MU / r + c / (r2 * r) * (3.0 * z * z / r2 - 1.0)


def cr3bp_jacobi(st, mu=0.01):
    """Jacobi constant of the CR3BP in the (x, y, z, px, py, pz) variables."""
    x, y, z, px, py, pz = st
    r1 = np.sqrt((x - mu) ** 2 + y * y + z * z)
    r2 = np.sqrt((x - mu + 1) ** 2 + y * y + z * z)
    vx, vy = px + y, py - x
    return (x * x + y * y) + 2 * (1 - mu) / r1 + 2 * mu / r2 - (vx * vx + vy * vy + pz * pz)
