"""CPU: the C oracle's restatement of the two-box (Gibbs ensemble) branches — MoveParticle's box selection (src/moves.cpp:127-137) and
SwapParticle's Gibbs branch (src/moves.cpp:342-400) — against golden vectors of the compiled reference
(tests/golden/reference_gibbs.npz, made by tests/golden/make_golden_gibbs.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, O


@pytest.fixture(scope="module")
def gibbs():
    return dict(np.load(os.path.join(GOLDEN, "reference_gibbs.npz"))), json.load(open(os.path.join(GOLDEN, "reference_gibbs.json")))


def code_mask(trace):
    """The direction of a transfer attempt whose donor box is empty (kind 6) cannot be observed in the reference: no box bit there."""
    return np.where(((trace >> 1) & 7) == 6, 0x0F, 0xFF).astype(np.uint8)


def oracle_pair(arrays, m, name, cap=1024):
    a, b = O.make_system(**m["box0"]), O.make_system(**m["box1"])
    oa, ob = O.COracle(a, cap), O.COracle(b, cap)
    oa.link(ob)
    for ib, o in ((0, oa), (1, ob)):
        o.set_positions(arrays[f"{name}_x{ib}"], arrays[f"{name}_y{ib}"], arrays[f"{name}_z{ib}"])
        box = o.box_energy()[0]
        assert np.allclose(box, arrays[f"{name}_box{ib}"], rtol=1e-12, atol=0)
    return oa, ob


@pytest.mark.parametrize("name", ["vle", "pore"])
def test_gibbs_trace_matches_the_reference(gibbs, name):
    arrays, meta = gibbs
    m = meta[name]
    oa, ob = oracle_pair(arrays, m, name)
    oa.seed(m["seed"])
    tr, per_set = [], []
    for s in range(1, m["sets"] + 1):
        oa.reset_stats(); ob.reset_stats()
        tr.append(oa.gibbs_run(m["steps_per_set"], m["correct"]))
        st = [oa.state(), ob.state()]
        wd = [oa.get_widom(), ob.get_widom()]
        oa.adjust(s); ob.adjust(s)
        row = []
        for ib, o in ((0, oa), (1, ob)):
            row += [st[ib]["n"], st[ib]["acceptance"], st[ib]["rejection"], st[ib]["n_displacements"], o.state()["dr"],
                    wd[ib][0][0], wd[ib][0][1], wd[ib][1][0], wd[ib][1][1]]
        row += [st[0]["accept_swap"], st[0]["reject_swap"], st[0]["n_swaps"]]
        per_set.append(row)
    tr = np.concatenate(tr)
    ref = arrays[name + "_trace"]
    mask = code_mask(ref)
    assert np.array_equal(tr & mask, ref & mask)
    got, want = np.array(per_set, dtype=np.float64), arrays[name + "_per_set"]
    assert np.allclose(got, want, rtol=1e-11, atol=0, equal_nan=True)
    for ib, o in ((0, oa), (1, ob)):
        x, y, z = o.get_positions()
        assert np.array_equal(x, arrays[f"{name}_fx{ib}"]) and np.array_equal(y, arrays[f"{name}_fy{ib}"]) and np.array_equal(z, arrays[f"{name}_fz{ib}"])
        assert np.allclose(o.box_energy()[0], arrays[f"{name}_fbox{ib}"], rtol=1e-12, atol=0)
    # the transfer conserves the total count
    assert oa.state()["n"] + ob.state()["n"] == len(arrays[f"{name}_x0"]) + len(arrays[f"{name}_x1"])


def test_gibbs_run_sets_is_the_same_loop(gibbs):
    """orc_gibbs_run_sets (ResetStats, steps, AdjustMCMoves on both boxes) == the per-set calls above."""
    arrays, meta = gibbs
    m = meta["vle"]
    oa, ob = oracle_pair(arrays, m, "vle")
    oa.seed(m["seed"])
    tr = oa.gibbs_run_sets(1, m["sets"], m["steps_per_set"], m["correct"], True)
    ref = arrays["vle_trace"]
    assert np.array_equal(tr & code_mask(ref), ref & code_mask(ref))
    assert abs(oa.state()["dr"] - arrays["vle_per_set"][-1][4]) < 1e-12 and abs(ob.state()["dr"] - arrays["vle_per_set"][-1][13]) < 1e-12
