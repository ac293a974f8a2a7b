"""CPU: the C oracle (oracle/mcpm_oracle.c) against the golden vectors made from the compiled reference."""
import numpy as np
import pytest

from conftest import O, rel_err, system_from_meta

TOL = 1e-12  # oracle restates the reference's arithmetic: agreement is at rounding level


def test_philox_known_answers():
    # Random123 known-answer vectors for philox4x32-10
    assert O.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = [O.philox_uniform(1, 2, s, k) for s in range(50) for k in range(8)]
    assert 0.0 <= min(u) and max(u) < 0.9999


def test_reference_rng_stream(golden):
    arrays, _ = golden
    got = O.mt_uniforms(12345, 4096)
    assert np.array_equal(got, arrays["rng_12345"])  # bit-exact: mt19937_64 + quirk Q1 mapping


def test_known_answers(golden):
    _, meta = golden
    ka = meta["known_answers"]
    s = O.argon_slit()
    c = O.COracle(s, 64)
    r0 = 2.0 ** (1.0 / 6.0) * 3.405
    c.set_positions([10.0, 10.0 + r0], [10.0, 10.0], [10.0, 10.0])
    assert c.particle_energy(0)[0] == pytest.approx(ka["two_ar_pair"], rel=TOL)
    assert ka["two_ar_pair"] == pytest.approx(-119.8, rel=1e-14)
    for z, key in ((3.4025, "wall_z_sigma"), (10.0, "wall_z_center"), (20.5, "wall_outside")):
        c.set_positions([10.0], [10.0], [z])
        assert c.particle_energy(0)[2] == pytest.approx(ka[key], rel=TOL)
    assert ka["wall_z_sigma"] == pytest.approx(-1094.356254192, rel=1e-12)
    assert ka["wall_z_center"] == pytest.approx(-62.2412425733286, rel=1e-13)
    assert c.swap_zz() == pytest.approx(ka["zz_c1"], rel=1e-15)
    assert c.volume() == ka["volume_c1"]


@pytest.mark.parametrize("name", ["c1", "c1raw", "ch4", "bulk"])
def test_energy_vectors(golden, name):
    arrays, meta = golden
    s = system_from_meta(meta[name]["system"])
    x, y, z = arrays[name + "_x"], arrays[name + "_y"], arrays[name + "_z"]
    c = O.COracle(s, len(x) + 8)
    c.set_positions(x, y, z)
    box, pp, pw = c.box_energy()
    assert rel_err(box, np.where(arrays[name + "_box"] == 0, 1e-300, arrays[name + "_box"])) < TOL or np.allclose(box, arrays[name + "_box"], rtol=TOL, atol=0)
    assert rel_err(pp, arrays[name + "_pp"]) < TOL
    assert np.allclose(pw, arrays[name + "_pw"], rtol=TOL, atol=0)
    for k, (i, t) in enumerate(zip(arrays[name + "_idx"], arrays[name + "_trials"])):
        assert np.allclose(c.particle_energy(int(i)), arrays[name + "_cur"][k], rtol=TOL, atol=0)
        assert np.allclose(c.moved_energy(int(i), *t), arrays[name + "_moved"][k], rtol=TOL, atol=0)
        assert np.allclose(c.trial_energy(*t), arrays[name + "_ghost"][k], rtol=TOL, atol=0)


@pytest.mark.parametrize("name,start", [("c1t", "c1"), ("lowt", "c1"), ("bulkt", "bulk")])
def test_trace_vectors(golden, name, start):
    """Same mt19937_64 seed -> identical accept/reject sequence, counters, dr and final configuration."""
    arrays, meta = golden
    m = meta[{"c1t": "c1_trace", "lowt": "low_trace", "bulkt": "bulk_trace"}[name]]
    s = system_from_meta(m["system"])
    n0 = m["n0"]
    x, y, z = arrays[start + "_x"][:n0], arrays[start + "_y"][:n0], arrays[start + "_z"][:n0]
    c = O.COracle(s, 1024)
    c.set_positions(x, y, z)
    c.seed(m["seed"])
    c.box_energy()
    traces, per_set = [], []
    for k in range(1, m["sets"] + 1):
        c.reset_stats()
        _, tr = c.run(m["steps_per_set"], 0, True)
        st = c.state()
        c.adjust(k)
        traces.append(tr)
        per_set.append([st["n"], st["acceptance"], st["rejection"], st["n_displacements"], st["accept_swap"],
                        st["reject_swap"], st["n_swaps"], c.state()["dr"]])
    assert np.array_equal(np.concatenate(traces), arrays[name + "_trace"])
    assert np.array_equal(np.array(per_set), arrays[name + "_per_set"])
    fx, fy, fz = c.get_positions()
    assert np.array_equal(fx, arrays[name + "_fx"]) and np.array_equal(fy, arrays[name + "_fy"]) and np.array_equal(fz, arrays[name + "_fz"])
    box, _, _ = c.box_energy()
    assert np.allclose(box, arrays[name + "_fbox"], rtol=TOL, atol=0)


def test_exact_bookkeeping_tracks_recompute(golden):
    arrays, meta = golden
    s = system_from_meta(meta["c1"]["system"])
    c = O.COracle(s, 1024)
    c.set_positions(arrays["c1_x"], arrays["c1_y"], arrays["c1_z"])
    c.seed(5)
    c.box_energy()
    c.run(20000, 0)
    st = c.state()
    box, _, _ = c.box_energy()
    assert st["exact_energy"] == pytest.approx(box[3], rel=1e-10)
    assert st["exact_pair"] == pytest.approx(box[0], rel=1e-10)
    assert st["exact_wall"] == pytest.approx(box[2], rel=1e-10)


def test_replay_equals_seeded(golden):
    arrays, meta = golden
    s = system_from_meta(meta["c1"]["system"])
    a = O.COracle(s, 1024); b = O.COracle(s, 1024)
    for c in (a, b):
        c.set_positions(arrays["c1_x"], arrays["c1_y"], arrays["c1_z"])
    a.seed(99)
    b.replay(O.mt_uniforms(99, 8 * 3000))
    _, ta = a.run(3000, 0, True); _, tb = b.run(3000, 0, True)
    assert np.array_equal(ta, tb)
    assert a.draws() == b.draws() <= 8 * 3000
