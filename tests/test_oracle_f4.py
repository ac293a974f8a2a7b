"""CPU: the C oracle's restatement of row f4 (spherical pore, hard spheres, NPT volume moves) against golden vectors of the compiled
reference (tests/golden/reference_f4.npz, made by tests/golden/make_golden_f4.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, O

RTOL = 1e-12
RTOL_CYL = 2e-10      # the reference's 2F1 (Numerical-Recipes hypgeo: series / ODE integration) is itself good to ~2e-12; sums of ~300 wall terms


@pytest.fixture(scope="module")
def f4():
    """reference_f4 (sphere, hard spheres, NPT) and reference_cyl (cylindrical walls) merged."""
    arrays = dict(np.load(os.path.join(GOLDEN, "reference_f4.npz")))
    arrays.update(dict(np.load(os.path.join(GOLDEN, "reference_cyl.npz"))))
    meta = json.load(open(os.path.join(GOLDEN, "reference_f4.json")))
    meta.update(json.load(open(os.path.join(GOLDEN, "reference_cyl.json"))))
    return arrays, meta


def close(a, b, rtol=RTOL):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return bool(np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(b), np.mean(np.abs(b)) + 1e-300)))


@pytest.mark.parametrize("name", ["sph", "hs", "hsov", "cyl3", "cyl4"])
def test_energies(f4, name):
    arrays, meta = f4
    s = O.make_system(**meta[name]["system"])
    x, y, z = arrays[name + "_x"], arrays[name + "_y"], arrays[name + "_z"]
    o = O.COracle(s, len(x) + 8)
    o.set_positions(x, y, z)
    box, pp, pw = o.box_energy()
    rt = RTOL_CYL if name.startswith("cyl") else RTOL
    assert close(box, arrays[name + "_box"], rt) and close(pp, arrays[name + "_pp"], rt) and close(pw, arrays[name + "_pw"], rt)
    for k, (i, t) in enumerate(zip(arrays[name + "_idx"], arrays[name + "_trials"])):
        assert close(o.particle_energy(int(i)), arrays[name + "_cur"][k], rt)
        assert close(o.moved_energy(int(i), *t), arrays[name + "_moved"][k], rt)
        assert close(o.trial_energy(*t), arrays[name + "_ghost"][k], rt)


def run_trace(o, m):
    tr, per_set, vol = [], [], []
    for s in range(1, m["sets"] + 1):
        o.reset_stats()
        _, t = o.run(m["steps_per_set"], m["correct"], True)
        tr.append(t)
        st = o.state(); vs = o.volume_state()
        o.adjust(s)
        per_set.append([st["n"], st["acceptance"], st["rejection"], st["n_displacements"], st["accept_swap"], st["reject_swap"], st["n_swaps"], o.state()["dr"]])
        vol.append([vs["volume"], vs["width"][2], o.volume_state()["dv"], vs["acceptance_vol"], vs["rejection_vol"], vs["n_vol_changes"], vs["as_shipped_energy"]])
    return np.concatenate(tr), np.array(per_set, dtype=np.float64), np.array(vol)


@pytest.mark.parametrize("name,mkey,start", [("spht", "sph_trace", "sph"), ("hst", "hs_trace", "hs"), ("hsgt", "hsg_trace", "hs"), ("nptt", "npt_trace", None),
                                              ("cyl3t", "cyl3_trace", "cyl3"), ("cyl4t", "cyl4_trace", "cyl4")])
def test_traces(f4, golden, name, mkey, start):
    arrays, meta = f4
    m = meta[mkey]
    s = O.make_system(**m["system"])
    if start is None:
        g, _ = golden
        x, y, z = g["bulk_x"], g["bulk_y"], g["bulk_z"]
    else:
        x, y, z = arrays[start + "_x"], arrays[start + "_y"], arrays[start + "_z"]
    o = O.COracle(s, 1024)
    o.set_positions(x, y, z)
    o.seed(m["seed"])
    o.box_energy()
    tr, per_set, vol = run_trace(o, m)
    d = np.nonzero(tr != arrays[name + "_trace"])[0]
    assert len(d) == 0, f"first divergence at step {d[0]}"
    assert np.array_equal(per_set, arrays[name + "_per_set"])
    assert np.allclose(vol[:, :3], arrays[name + "_vol"][:, :3], rtol=1e-13, atol=0) and np.array_equal(vol[:, 3:6], arrays[name + "_vol"][:, 3:6])
    if name == "nptt":
        assert np.allclose(vol[:, 6], arrays[name + "_vol"][:, 6], rtol=1e-10, atol=0)     # the reference's running box energy, as shipped
    fx, fy, fz = o.get_positions()
    assert np.allclose(fx, arrays[name + "_fx"], rtol=1e-13, atol=1e-13) and np.allclose(fz, arrays[name + "_fz"], rtol=1e-13, atol=1e-13)
    if name.startswith("cyl"):
        assert close(o.box_energy()[0], arrays[name + "_fbox"], RTOL_CYL)
        return
    assert close(o.box_energy()[0], arrays[name + "_fbox"], 1e-10)
