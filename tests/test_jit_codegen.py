"""CPU: the run-time kernel generator (csrc/hy_jit.hpp).  No device is needed to generate and
compile a kernel: hy_jit_precompile lowers the tape to CUDA source, compiles it with NVRTC for
sm_100a and stores the image in the kernel cache (what __graft_entry__.build() does for the
BASELINE tapes).  Running the kernels is covered by tests/test_gpu_jit.py."""

import glob
import os
import struct

import numpy as np
import pytest

import hy_b200 as hy
from hy_b200 import _cabi
from hy_b200 import decompose as D
from hy_b200 import workloads as W


@pytest.fixture()
def jit_cache(tmp_path, monkeypatch):
    monkeypatch.setenv("HY_CUDA_JIT_CACHE", str(tmp_path))
    monkeypatch.setenv("HY_CUDA_JIT", "1")  # compile whatever the size of the tape
    return tmp_path


def _images(path):
    out = []
    for f in glob.glob(os.path.join(str(path), "*.hyjit")):
        b = open(f, "rb").read()
        assert b[:4] == b"HYJ1"
        nl = struct.unpack("<I", b[4:8])[0]
        out.append((b[8:8 + nl].decode(), b[8 + nl:]))
    return out


def test_every_op_kind_compiles_and_is_cached(jit_cache):
    x, v, s = hy.make_vars("x", "v", "s")
    sys_ = [(x, v),
            (v, hy.cos(hy.time) - hy.par[0] * v - hy.sin(x) + 0.01 * hy.exp(-s * s) / (1.0 + x * x)),
            (s, hy.log(2.0 + hy.cos(x)) * hy.sqrt(1.0 + v * v) - hy.par[1] * s + (1.5 + x * x) ** -1.5 + (x * v) * (s * v))]
    dc = D.decompose(sys_, 20, events=[v * v - 1.0])
    kinds = {D.OP_NAMES[int(o["opcode"])] for o in dc.ops}
    assert {"lincomb", "mul", "square", "div", "pow", "sqrt", "exp", "log", "sincos", "time"} <= kinds
    for fp_bits in (64, 32):
        fc, secs = _cabi.jit_precompile(dc, fp_bits, batch=4096, n_tevents=1)
        assert fc == 0 and secs > 0.0          # compiled now
        fc, secs = _cabi.jit_precompile(dc, fp_bits, batch=4096, n_tevents=1)
        assert fc == 1                          # second time: from the cache
    imgs = _images(jit_cache)
    assert len(imgs) == 2
    names = sorted(n for n, _ in imgs)
    # the persistent propagate kernel, one thread per trajectory (G = 1), FP32 and FP64
    assert all("propagate_kernel" in n for n in names), names
    assert any("IdLi1E" in n for n in names) and any("IfLi1E" in n for n in names), names
    for _, cubin in imgs:
        assert cubin[:4] == b"\x7fELF"


def test_small_tapes_stay_on_the_interpreter_by_default(jit_cache, monkeypatch):
    monkeypatch.setenv("HY_CUDA_JIT", "2")
    dc = D.decompose(W.pendulum_sys(), 20)
    fc, _ = _cabi.jit_precompile(dc, 64)
    assert fc == -1
    assert _images(jit_cache) == []


def test_order_blocking_variant_compiles(jit_cache, monkeypatch):
    # HY_CUDA_JIT_BLOCK = 4: products / quotients in blocks of four orders (another kernel image)
    vs = hy.var_ode_sys(W.kepler_j2_sys(), hy.var_args.vars)
    dc = D.decompose(vs.sys, 20)
    assert _cabi.jit_precompile(dc, 64, batch=64)[0] == 0
    monkeypatch.setenv("HY_CUDA_JIT_BLOCK", "4")
    assert _cabi.jit_precompile(dc, 64, batch=64)[0] == 0
    assert len(_images(jit_cache)) == 2


def test_event_code_generation_without_a_device(tmp_path, monkeypatch):
    # hy_jit_precompile_events: the event functions of a system on a register-resident kernel, generated
    # and compiled device-less.  The source shows the lane-parallel plan: config 5's nine squares are five
    # distinct ones ((x - mu)^2, (x - mu + 1)^2, x^2, y^2, z^2), "x - mu" never becomes a jet of its own.
    import hy_b200 as hy
    from hy_b200 import _cabi, decompose as D, workloads as W

    monkeypatch.setenv("HY_CUDA_JIT_CACHE", str(tmp_path / "cache"))
    monkeypatch.setenv("HY_CUDA_JIT_DUMP", str(tmp_path))
    mu, order = 0.01, 20
    x, y, z = hy.make_vars("x", "y", "z")
    evs = [(x - mu) ** 2 + y * y + z * z - 0.012 ** 2, (x - mu + 1.0) ** 2 + y * y + z * z - 0.012 ** 2,
           x * x + y * y + z * z - 25.0]
    sys5 = W.cr3bp_sys(mu)
    args = (D.decompose(sys5, order, events=evs), D.decompose(sys5, order),
            D.decompose_event_tape(evs, [l.name for l, _ in sys5], order))
    fc, secs = _cabi.jit_precompile_events(*args, 64, n_tevents=3)
    assert fc == 0 and secs > 0
    src = "".join(p.read_text() for p in tmp_path.glob("hy_jit_*.cu"))
    assert "5 product units on the lanes of the group" in src
    assert src.count("evt_unit_sq<R, XS, 20>") == 3            # 5 squares on 2 lanes: 3 rounds
    norms = src.split("hy_gen_evt_order")[0]
    assert "evt_exec_all" not in norms                         # the folded operands are not materialised ...
    assert src.split("hy_gen_evt_order")[1].count("evt_exec_all") == 2   # ... unless the root finder runs
    assert _cabi.jit_precompile_events(*args, 64, n_tevents=3)[0] == 1     # second time: from the cache
    # a system with no register-resident kernel has no such build
    xx, vv = hy.make_vars("x", "v")
    pend = [(xx, vv), (vv, -9.8 * hy.sin(xx))]
    pe = [xx - 0.1]
    assert _cabi.jit_precompile_events(D.decompose(pend, order, events=pe), D.decompose(pend, order),
                                       D.decompose_event_tape(pe, ["x", "v"], order), 64, n_tevents=1)[0] == -1


def test_event_code_generation_for_a_16_lane_group(tmp_path, monkeypatch):
    # N-body register kernel (16 lanes per trajectory): differences of state jets are every-order ops (lane 0,
    # then a warp sync), their squares / the products of state jets are spread over up to 8 lanes per round,
    # "x + c" operands of a product are folded, the event function that IS a product is stored as a jet
    import hy_b200 as hy
    from hy_b200 import _cabi, decompose as D, workloads as W

    monkeypatch.setenv("HY_CUDA_JIT_CACHE", str(tmp_path / "cache"))
    monkeypatch.setenv("HY_CUDA_JIT_DUMP", str(tmp_path))
    sys_ = W.oss_sys()
    v = lambda s: hy.expression(s)
    dx, dy = v("x_1") - v("x_2"), v("y_1") - v("y_2")
    evs = [dx * dx + dy * dy - 60.0, v("x_1") * v("vy_1") - v("y_1") * v("vx_1") - 2.0, (v("x_1") + 0.5) * (v("y_1") - 0.25)]
    order = 20
    fc, secs = _cabi.jit_precompile_events(D.decompose(sys_, order, events=evs), D.decompose(sys_, order),
                                           D.decompose_event_tape(evs, [l.name for l, _ in sys_], order), 64)
    assert fc == 0
    src = "".join(p.read_text() for p in tmp_path.glob("hy_jit_*.cu"))
    assert "propagate_kernel" not in src.split("namespace hy {")[1].split("hy_gen_evt_norms")[0]
    assert "5 product units on the lanes of the group" in src
    norms = src.split("hy_gen_evt_order")[0]
    assert norms.count("evt_exec_all") == 2 and "__syncwarp();" in norms      # dx, dy: every-order ADDSUBs on lane 0
    assert "evt_unit_sq<R, 1, 20>" in norms                                   # squares of unit-stride jets (dx, dy)
    assert "evt_unit_mul<R, XS, XS, 20>" in norms and "sub == 2u ?" not in norms.split("evt_unit_sq")[0]
    assert "const bool ala = " in norms and "const bool alb = " in norms      # (x_1 + 0.5) * (y_1 - 0.25): both folded
    assert "sub < 3u" in norms                                                # three products in one round
