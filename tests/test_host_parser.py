"""CPU: the C++ host driver's input parser (mcpm_read_input_file) against the reference's ReadInputFile."""
import ctypes as C
import os

import pytest

from conftest import GOLDEN


def parse(mcpm, path):
    cfg = mcpm.SimConfig()
    rc = mcpm.lib().mcpm_read_input_file(path.encode(), C.byref(cfg))
    return rc, cfg


def test_parser_matches_reference_golden(mcpm, golden):
    _, meta = golden
    want = meta["parser_c1"]
    rc, cfg = parse(mcpm, os.path.join(GOLDEN, "ar_cfg1.in"))
    assert rc == 0
    assert cfg.project.decode() == want["project"] and cfg.fluid.decode() == want["fluid"] and cfg.box.decode() == want["box"]
    assert [cfg.n_sets, cfg.n_equil_sets, cfg.steps_per_set, cfg.print_every] == want["sets"]
    assert cfg.n_particles == want["n0"] and cfg.n_boxes == 1 and cfg.n_species == 1
    s = cfg.system
    for key, val in want["system"].items():
        got = list(s.width) if key == "width" else getattr(s, key)
        assert got == val, key          # exact: same parsing (stod/stoi) and the same derived widths, dr


def test_parser_live_against_reference(mcpm, ref, tmp_path):
    """Case-insensitive keywords, trailing ';', names lower-cased, unknown lines ignored (src/read.cpp:50-121)."""
    from conftest import O
    r2 = O.RefOracle()      # ReadInputFile accumulates into counters: fresh object
    try:
        path = os.path.join(GOLDEN, "ar_small.in")
        rc_ref, ps, n0, proj, fname, bname, sets = r2.read_input(path)
        rc, cfg = parse(mcpm, path)
        assert rc == 0 and rc_ref == 0
        assert (cfg.project.decode(), cfg.fluid.decode(), cfg.box.decode()) == (proj, fname, bname) == ("ar_small", "ar", "pore")
        assert cfg.n_particles == n0
        assert [cfg.n_sets, cfg.n_equil_sets, cfg.steps_per_set, cfg.print_every] == list(sets)
        for name, _ in ps._fields_:
            if name in ("pair_kind", "pressure", "dv"):      # compared below / not part of every file
                continue
            a = getattr(ps, name); b = getattr(cfg.system, name)
            assert (list(a) == list(b)) if name == "width" else (a == b), name
        assert cfg.has_seed == 1 and cfg.seed == 4242 and cfg.print_trajectory == 1
    finally:
        r2.close()


def test_parser_quirks(mcpm, tmp_path):
    text = open(os.path.join(GOLDEN, "ar_cfg1.in")).read()
    # a slit with "steele10-4-3" matches no wall potential in the reference's dispatch (src/energy.cpp:69-77)
    p = tmp_path / "steele.in"
    p.write_text(text.replace("pore ar VdwPotential lj", "pore ar VdwPotential steele10-4-3"))
    rc, cfg = parse(mcpm, str(p))
    assert rc == 0 and cfg.system.wall_kind == mcpm.WALL_NONE
    p = tmp_path / "bulk.in"
    p.write_text(text.replace("Geometry slit", "GEOMETRY Bulk;").replace("Size 20.0", "size 50.0"))
    rc, cfg = parse(mcpm, str(p))
    assert rc == 0 and cfg.system.geometry == mcpm.GEOM_BULK and list(cfg.system.width) == [50.0, 50.0, 50.0]
    assert cfg.system.dr == pytest.approx(5.0)
    rc, _ = parse(mcpm, str(tmp_path / "missing.in"))
    assert rc == 5
    p = tmp_path / "ext.in"
    p.write_text(text + "Replicas 8\nChemicalPotentialSweep -1800 -1134 40\nDevices 8\nCapacity 768\n")
    rc, cfg = parse(mcpm, str(p))
    assert rc == 0 and (cfg.replicas, cfg.sweep_points, cfg.devices, cfg.capacity) == (8, 40, 8, 768)
    assert (cfg.sweep_mu0, cfg.sweep_mu1) == (-1800.0, -1134.0)


def test_malformed_values_return_a_status(mcpm, tmp_path):
    """mcpm.h promises status codes, never an exception across the C boundary: a value std::stod / stoi cannot parse must
    come back as MCPM_ERR_ARG with a message (through ctypes an escaping C++ exception is std::terminate)."""
    text = open(os.path.join(GOLDEN, "ar_cfg1.in")).read()
    for bad in ("ProductionSets\n", "Seed abc\n", "ar ar Sigma x3.4\n", "NSwapAttempts\n"):
        p = tmp_path / "bad.in"
        p.write_text(text + bad)
        rc, _ = parse(mcpm, str(p))
        assert rc == 1, bad
        assert b"malformed" in mcpm.lib().mcpm_last_error()
        assert mcpm.lib().mcpm_simulate(str(p).encode(), str(tmp_path).encode(), 2) == 1
    rc, _ = parse(mcpm, str(tmp_path / "nope.in"))
    assert rc == 5 and b"cannot open" in mcpm.lib().mcpm_last_error()


def test_unsupported_inputs_are_refused_before_any_device_work(mcpm, tmp_path):
    text = open(os.path.join(GOLDEN, "ar_cfg1.in")).read()
    p = tmp_path / "restart.in"
    p.write_text(text + "ContinueAfterCrash\n")
    assert mcpm.lib().mcpm_simulate(str(p).encode(), str(tmp_path).encode(), 2) == 4          # MCPM_ERR_UNSUPPORTED
    assert b"ContinueAfterCrash" in mcpm.lib().mcpm_last_error()


@pytest.mark.parametrize("fname,project", [("ar_sphere.in", "ar_sphere"), ("lj_npt.in", "lj_npt")])
def test_parser_f4_inputs_live_against_reference(mcpm, ref, fname, project):
    """Spherical pore (widths = Size on all axes, "lj" + sphere = SphericalLJ_Pot) and an NPT input (ExternalPressure, NVolumeAttempts)."""
    from conftest import O
    r2 = O.RefOracle()
    try:
        path = os.path.join(GOLDEN, fname)
        rc_ref, ps, n0, proj, fname_, bname, sets = r2.read_input(path)
        rc, cfg = parse(mcpm, path)
        assert rc == 0 and rc_ref == 0 and cfg.project.decode() == proj == project
        assert cfg.n_particles == n0 and [cfg.n_sets, cfg.n_equil_sets, cfg.steps_per_set, cfg.print_every] == list(sets)
        for name, _ in ps._fields_:
            a = getattr(ps, name); b = getattr(cfg.system, name)
            assert (list(a) == list(b)) if name == "width" else (a == b), name
    finally:
        r2.close()


def test_parser_recognises_what_the_engine_refuses(mcpm, tmp_path):
    text = open(os.path.join(GOLDEN, "ar_sphere.in")).read()
    p = tmp_path / "cyl.in"
    p.write_text(text.replace("Geometry sphere", "Geometry cylinder").replace("Cavity Ar VdwPotential lj", "Cavity Ar VdwPotential steele10-4-3"))
    rc, cfg = parse(mcpm, str(p))
    assert rc == 0 and cfg.system.geometry == mcpm.GEOM_CYLINDER and cfg.system.wall_kind == 4      # CylindricalSteele10_4_3: supported
    assert list(cfg.system.width) == [30.0, 30.0, 30.0]            # PBC length 2*rcut along x, diameter on y and z
    p = tmp_path / "eam.in"
    p.write_text(text.replace("Ar Ar VdwPotential lj", "Ar Ar VdwPotential eam_na"))
    assert mcpm.lib().mcpm_simulate(str(p).encode(), str(tmp_path).encode(), 2) == 4
    assert b"EAM" in mcpm.lib().mcpm_last_error()
    p = tmp_path / "gibbs.in"
    p.write_text(open(os.path.join(GOLDEN, "lj_npt.in")).read().replace("ExternalPressure 3.0e7\n", ""))
    assert mcpm.lib().mcpm_simulate(str(p).encode(), str(tmp_path).encode(), 2) == 4
    assert b"Gibbs" in mcpm.lib().mcpm_last_error()


def test_two_box_input(mcpm):
    """A second BoxName block: box 1 of a two-box (Gibbs) system with its own geometry / size / wall; fluid, temperature and move weights shared."""
    import ctypes as C
    cfg = mcpm.SimConfig()
    assert mcpm.lib().mcpm_read_input_file(os.path.join(GOLDEN, "gibbs_ar.in").encode(), C.byref(cfg)) == 0
    assert (cfg.n_boxes, cfg.n_species) == (2, 1)
    assert (cfg.box.decode(), cfg.box2.decode()) == ("liquid", "vapour") and (cfg.n_particles, cfg.n_particles2) == (230, 60)
    assert list(cfg.fix_volume) == [1, 1]
    assert list(cfg.system.width) == [24.0, 24.0, 24.0] and list(cfg.system2.width) == [40.0, 40.0, 40.0]
    assert cfg.system.dr == pytest.approx(2.4) and cfg.system2.dr == pytest.approx(4.0)               # 0.1 * width, src/read.cpp:168
    for s in (cfg.system, cfg.system2):
        assert (s.n_disp, s.n_vol, s.n_swap, s.temperature, s.rcut, s.eps_ff, s.sig_ff) == (4, 0, 1, 120.0, 10.0, 119.8, 3.405)
