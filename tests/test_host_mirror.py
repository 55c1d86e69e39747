"""The host-side mirror (qse_b200.host) against the reference generators' golden output."""
import numpy as np
import pytest

from qse_b200 import host


def test_lattice_positions_bit_exact(golden):
    g = golden("lattices.npz")
    for c in range(int(g["n_cases"])):
        name = str(g[f"{c}_name"])
        args = [float(a) if i == 0 and name != "torus" else a for i, a in enumerate(g[f"{c}_args"])]
        if name == "torus":
            args = [int(args[0]), int(args[1]), float(args[2]), float(args[3])]
        else:
            args = [args[0]] + [int(a) for a in args[1:]]
        q = getattr(host.lattices, name)(*args)
        assert np.array_equal(q.positions, g[f"{c}_positions"]), (name, args)
        if f"{c}_cell" in g:
            assert np.array_equal(q.cell.lattice_vectors, g[f"{c}_cell"])
            assert np.allclose(q.cell.reciprocal(), g[f"{c}_reciprocal"], rtol=0, atol=1e-15)


def test_repeat_checks():
    with pytest.raises(Exception, match="must be an integer greater than 1"):
        host.lattices.square(1.0, 1, 3)
    with pytest.raises(Exception, match="must be an integer greater than 1"):
        host.lattices.chain(1.0, 2.5)


def test_signal_schedule(golden):
    g = golden("signals.npz")
    s1 = host.Signal(np.linspace(0, 1, 5), 20)
    s2 = host.Signal([3.0], 400)
    ss = host.Signals([s1, s2])
    assert s1.time_per_value() == int(g["s1_tpv"]) and s2.time_per_value() == int(g["s2_tpv"])
    assert np.array_equal(s1.expand(), g["s1_expand"])
    assert ss.duration == int(g["total"]) and len(ss) == int(g["n"])
    assert np.array_equal(ss.expand(), g["expand"])
    with pytest.raises(ValueError, match="must divide"):
        host.Signal(np.ones(6), 400)  # the stale reference script case (400 % 6 != 0)
    with pytest.raises(ValueError, match="must be an ints"):
        host.Signal([1.0]).duration = 2.5
    assert (2 * s2 + 1.0).values[0] == 7.0
    with pytest.raises(TypeError):
        s1 + s2


def test_blockade_radius(golden):
    g = golden("interaction.npz")
    assert host.blockade_radius(2 * np.pi) == float(g["blockade_radius_2pi"])


def test_qbits_attach_hook():
    class Calc:
        def set_qbits(self, q):
            self.got = q

    c = Calc()
    q = host.Qbits(np.zeros((2, 2)), calculator=c)
    assert c.got is q and len(q) == 2 and q.dim == 2


def test_random_register_is_deterministic():
    a = host.lattices.random_register(12, 5.0)
    b = host.lattices.random_register(12, 5.0)
    assert np.array_equal(a.positions, b.positions)
    d = a.get_all_distances()
    assert d[np.triu_indices(12, 1)].min() >= 0.6 * 5.0
