"""CPU: the model builders the path uses, scenario by scenario as in the reference's own test
(/root/reference/heyoka/_test_model.py:101-278: fixed_centres, nbody / np1body, pendulum, cr3bp)."""

import numpy as np
import pytest

import hy_b200 as hy
from hy_b200 import expression as ex
from hy_b200 import make_vars, model


def test_fixed_centres():
    x, y, z, vx, vy, vz = make_vars("x", "y", "z", "vx", "vy", "vz")
    dyn = model.fixed_centres(Gconst=1.5, masses=[1.1], positions=[[1.0, 2.0, 3.0]])
    assert dyn[0][0] == x and dyn[0][1] == ex("vx")
    model.fixed_centres_energy(Gconst=1.5, masses=[1.1], positions=[[1.0, 2.0, 3.0]])
    model.fixed_centres_potential(Gconst=1.5, masses=[1.1], positions=[[1.0, 2.0, 3.0]])
    with pytest.raises(ValueError) as cm:
        model.fixed_centres(Gconst=1.5, masses=[1.1], positions=[1.0, 2.0, 3.0])
    assert ("Invalid positions array in a fixed centres model: the number of dimensions must be 2, but it is 1 "
            "instead") in str(cm.value)
    with pytest.raises(ValueError) as cm:
        model.fixed_centres(Gconst=1.5, masses=[1.1], positions=[[1.0, 2.0, 3.0, 4.0]])
    assert ("Invalid positions array in a fixed centres model: the number of columns must be 3, but it is 4 "
            "instead") in str(cm.value)
    with pytest.raises(TypeError) as cm:
        model.fixed_centres(Gconst=1.5, masses=[1.1], positions=[[{}, {}, {}]])
    assert ("The positions array in a fixed centres model could not be converted into an array of expressions - please "
            "make sure that the array's values can be converted into heyoka expressions") in str(cm.value)
    dyn = model.fixed_centres(Gconst=np.single(1.5), masses=[np.single(1.1)], positions=[[1.0, 2.0, 3.0]])
    assert dyn[0][0] == x and dyn[0][1] == ex("vx")
    # (the reference keeps G and the single-precision mass 1.10000002... as separate factors; here G * m is one
    #  constant: 1.5 * 1.1000000238... = 1.65000003...)
    assert "1.65000003" in repr(dyn)


def test_nbody():
    dyn = model.nbody(2, masses=[0.0, 0.0])
    assert len(dyn) == 12
    for i in (3, 4, 5, 9, 10, 11):
        assert dyn[i][1] == ex(0.0)
    dyn = model.nbody(2, Gconst=5.0)
    for i in (3, 4, 5, 9, 10, 11):
        assert "5.0000000000000" in str(dyn[i][1])
    assert model.nbody_energy(2, masses=[0.0, 0.0]) == ex(0.0)
    assert "5.0000000000000" in str(model.nbody_energy(2, Gconst=5.0))
    dyn = model.np1body(2, masses=[0.0, 0.0])
    assert len(dyn) == 6
    for i in (3, 4, 5):
        assert dyn[i][1] == ex(0.0)
    dyn = model.np1body(2, Gconst=5.0)
    for i in (3, 4, 5):
        assert "10.0000000000000" in str(dyn[i][1])
    assert model.np1body_energy(2, masses=[]) == ex(0.0)
    assert "5.0000000000000" in str(model.np1body_energy(2, Gconst=5.0))


def test_pendulum():
    x, v = make_vars("x", "v")
    sin, cos = hy.sin, hy.cos
    dyn = model.pendulum()
    assert dyn[0][0] == x and dyn[0][1] == v and dyn[1][0] == v and dyn[1][1] == -sin(x)
    dyn = model.pendulum(gconst=2.0)
    assert dyn[1][1] == -2.0 * sin(x)
    dyn = model.pendulum(gconst=4.0, length=2.0)
    assert dyn[1][1] == -2.0 * sin(x)
    assert model.pendulum_energy() == ((0.5 * v**2) + (1.0 - cos(x)))
    assert model.pendulum_energy(gconst=2.0) == ((0.5 * v**2) + (2.0 * (1.0 - cos(x))))
    assert model.pendulum_energy(length=2.0, gconst=4.0) == ((2.0 * v**2) + (8.0 * (1.0 - cos(x))))


def test_cr3bp():
    x, px, y = make_vars("x", "px", "y")
    dyn = model.cr3bp()
    assert dyn[0][0] == x and dyn[0][1] == px + y
    dyn = model.cr3bp(mu=1.0 / 2**4)
    assert "0.06250000000" in str(dyn[3][1])
    assert "0.00100000" in str(model.cr3bp_jacobi())
    assert "0.06250000000" in str(model.cr3bp_jacobi(mu=1.0 / 2**4))
