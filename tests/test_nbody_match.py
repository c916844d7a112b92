"""CPU tests of the tape matcher that selects the register-resident N-body
kernel (csrc/hy_nbody_match.hpp, through the C ABI hy_tape_kernel_variant; no
device needed).  Anything that is not exactly the tape of the reference's
model.nbody (expose_models.cpp:237-272) must stay on the tape interpreter."""

import numpy as np

import hy_b200 as hy
from hy_b200 import _cabi, decompose as D

import common


def _variant(sys_, order=20, **kw):
    return _cabi.tape_kernel_variant(D.decompose(sys_, order, **kw))


def test_outer_solar_system_matches():
    assert _variant(common.oss_sys()) == 6
    # any masses / G, and any order up to 20
    sys_ = hy.model.nbody(6, masses=[3.0, 1e-3, 2e-4, 5e-5, 1e-5, 7e-9], Gconst=0.5)
    assert _variant(sys_) == 6
    assert _variant(sys_, order=12) == 6
    assert _variant(sys_, order=9) == 6


def test_smaller_body_counts_match():
    for nb in (3, 4, 5):
        assert _variant(hy.model.nbody(nb, masses=[1.0] + [1e-3] * (nb - 1), Gconst=0.7)) == nb


def test_non_matching_tapes_keep_the_interpreter():
    # orders above the unrolled maximum (20; 22 for the 6-body high-accuracy build, variant 226)
    assert _variant(common.oss_sys(), order=22) == 226
    assert _variant(common.oss_sys(), order=21) == 226
    assert _variant(common.oss_sys(), order=23) == 0
    assert _variant(hy.model.nbody(5, masses=[1.0, 1e-3, 1e-3, 1e-3, 1e-3]), order=22) == 0
    # other systems
    assert _variant(common.pendulum_sys()) == 0
    # body counts without a kernel (3..6 are precompiled, 7 and 8 are built at hy_create time on 32-lane groups)
    assert _variant(hy.model.nbody(2)) == 0
    assert _variant(hy.model.nbody(7)) == 7
    assert _variant(hy.model.nbody(8)) == 8
    assert _variant(hy.model.nbody(9)) == 0
    assert _variant(hy.model.nbody(7), order=22) == 0
    # a massless body drops terms from the sums: not the full pattern
    assert _variant(hy.model.nbody(6, masses=[1.0, 1e-3, 1e-3, 1e-3, 1e-3, 0.0])) == 0


def test_parametric_masses_match():
    # masses as runtime parameters (the reference's model.nbody accepts expressions): the acceleration
    # terms carry par[j]; the matcher accepts them (the kernel is then built at hy_create time with
    # HY_NBR_PAR: tests/test_gpu_jit.py)
    sys_ = hy.model.nbody(6, masses=[hy.par[i] for i in range(6)], Gconst=0.5)
    assert D.decompose(sys_, 20).n_par == 6
    assert _variant(sys_) == 6
    mixed = hy.model.nbody(4, masses=[1.0, hy.par[0], 1e-3, hy.par[1]])
    assert _variant(mixed) == 4


def test_perturbed_nbody_is_rejected():
    # an extra force term / a different exponent must not be swallowed by the matcher
    sys_ = common.oss_sys()
    x0 = sys_[0][0]
    pert = list(sys_)
    pert[3] = (pert[3][0], pert[3][1] + 1e-9 * x0)
    assert _variant(pert) == 0
    xs = hy.make_vars(*["q{}".format(i) for i in range(36)])
    # same structure with r^-2 forces (exponent -1.0 instead of -1.5)
    eqs = []
    acc = {i: [0, 0, 0] for i in range(6)}
    for i in range(6):
        for j in range(i + 1, 6):
            dd = [xs[6 * j + c] - xs[6 * i + c] for c in range(3)]
            w = hy.pow(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2], -1.0)
            for c in range(3):
                acc[i][c] = acc[i][c] + 0.3 * (dd[c] * w)
                acc[j][c] = acc[j][c] + (-0.2) * (dd[c] * w)
    for i in range(6):
        for c in range(3):
            eqs.append((xs[6 * i + c], xs[6 * i + 3 + c]))
        for c in range(3):
            eqs.append((xs[6 * i + 3 + c], acc[i][c]))
    assert _variant(eqs) == 0


# ---- the CR3BP matcher (csrc/hy_cr3bp_match.hpp): exactly the tape of model.cr3bp
# (expose_models.cpp:395-400) at the compiled orders (20 in FP64, 9 in FP32) ----
CRB = 203


def test_cr3bp_matches():
    for mu in (0.01, 1e-3, 0.3):
        assert _variant(hy.model.cr3bp(mu=mu)) == CRB
        assert _variant(hy.model.cr3bp(mu=mu), order=9) == CRB
    assert _variant(common.cr3bp_sys()) == CRB


def test_cr3bp_lookalikes_keep_the_interpreter():
    # orders above the unrolled maximum (lower ones take the order-checked path)
    assert _variant(common.cr3bp_sys(), order=12) == CRB
    assert _variant(common.cr3bp_sys(), order=22) == 222  # (FP64 order-22 build)
    assert _variant(common.cr3bp_sys(), order=23) == 0
    # same structure with one altered equation
    sys_ = hy.model.cr3bp(mu=0.01)
    (x, fx), (y, fy), (z, fz), (px, fpx), (py, fpy), (pz, fpz) = sys_
    assert _variant([(x, fx), (y, fy), (z, fz), (px, fpx), (py, fpy), (pz, fpz + 1e-3 * z)]) == 0
    assert _variant([(x, fx), (y, fy), (z, 2.0 * fz), (px, fpx), (py, fpy), (pz, fpz)]) == 0
    assert _variant([(x, fx), (y, fy + x), (z, fz), (px, fpx), (py, fpy), (pz, fpz)]) == 0
    # variables in another order
    assert _variant([(y, fy), (x, fx), (z, fz), (px, fpx), (py, fpy), (pz, fpz)]) == 0
