"""`.ini` case files stay unchanged: microhh_b200.case reads the reference's shipped cases with the reference's defaults
(needs /root/reference; skipped on the GPU box) and maps them onto the hot path's switches."""
import glob
import os

import numpy as np
import pytest

from microhh_b200 import capi
from microhh_b200.case import CaseConfig, read_ini

REF = "/root/reference/cases"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


def case(name):
    return CaseConfig.from_file(os.path.join(REF, name, name + ".ini"))


def test_every_shipped_case_parses():
    inis = [p for p in glob.glob(os.path.join(REF, "*", "*.ini"))
            if os.path.basename(p)[:-4] == os.path.basename(os.path.dirname(p))]
    assert len(inis) >= 20
    for p in inis:
        c = CaseConfig.from_file(p)
        assert c.itot > 0 and c.ktot > 0 and c.order in ("2", "4")
        assert isinstance(c.unsupported(), list)


def test_drycblles_as_shipped():
    """the shipped file has swadvec=2 (2i5 is a test permutation), smag2, thermo dry, surface model."""
    c = case("drycblles")
    assert (c.itot, c.jtot, c.ktot) == (128, 128, 128) and c.order == "2"
    assert (c.swadvec, c.swdiff, c.swpres, c.swthermo) == ("2", "smag2", "2", "dry")
    assert c.unsupported() == [] and c.scalars == ["th"]
    p = c.make_params()
    assert (p.swadvec, p.swdiff, p.swthermo, p.surface_model) == (2, 1, 1, 1)
    assert (p.mbcbot, p.mbctop) == (capi.BC_DIRICHLET, capi.BC_NEUMANN)       # noslip / freeslip
    assert p.sbcbot[0] == capi.BC_NEUMANN and p.sbctop[0] == capi.BC_NEUMANN    # flux / neumann
    assert abs(c.dnmax - 0.3) < 1e-12 and abs(c.cflmax - 1.2) < 1e-12 and c.visc == 1e-5
    gd = c.grid_data()
    assert (gd.igc, gd.jgc, gd.kgc) == (1, 1, 1) and gd.icells == 130
    c.ini["advec"]["swadvec"] = "2i5"                                            # the reference's test permutation
    c2 = CaseConfig(c.ini)
    assert c2.ghost_cells() == (3, 3, 1) and c2.make_params().swadvec == 25


def test_defaults_follow_the_grid_order():
    tg = case("taylorgreen")       # no [advec]/[diff]/[pres] switches: defaults = swspatialorder
    assert (tg.order, tg.swadvec, tg.swdiff, tg.swpres, tg.swthermo) == ("2", "2", "2", "2", "0")
    assert tg.jtot == 1 and tg.unsupported() == []
    m5 = case("moser590")
    assert (m5.order, m5.swadvec, m5.swdiff, m5.swpres) == ("4", "4", "4", "4") and m5.unsupported() == []
    assert m5.ghost_cells() == (3, 3, 3) and m5.grid_data().order == 4
    assert m5.make_params().swadvec == 4


def test_out_of_path_switches_are_named():
    m180 = case("moser180")                                                      # ships with swadvec=4m
    assert m180.swadvec == "4m" and m180.unsupported() == [] and m180.make_params().swadvec == 41
    bx = case("bomex")                                                           # moist is on the path, mbcbot=ustar is not
    assert bx.scalars == ["thl", "qt"] and bx.unsupported() == ["mbcbot/mbctop=ustar/freeslip"]
    assert bx.thermo_moist_params() == dict(pbot=101500., swupdatebasestate=True)
    with pytest.raises(ValueError):
        bx.make_params()
    assert "swmicro=2mom_warm" in case("dycoms").unsupported()
    g4 = case("gabls4s3")                                                        # swadvec=2i4 + smag2 + dry
    assert g4.unsupported() == [] and g4.make_params().swadvec == 24 and g4.ghost_cells() == (2, 2, 2)
    assert case("weisman_klemp").unsupported() == ["swmicro=nsw6"] and case("weisman_klemp").ghost_cells()[:2] == (3, 3)
    arm = case("arm")
    assert arm.unsupported() == [] and arm.make_params().swthermo == 3
    ws = case("weakscaling")                                                     # 4th-order DNS with swthermo=buoy
    assert ws.scalars == ["b"] and ws.unsupported() == [] and ws.make_params().swthermo == 2
    assert ws.thermo_buoy_params() == dict(alpha=0., n2=0., utrans=0., swbaroclinic=False, dbdy_ls=0.)
    ps = case("prandtlslope")
    tb = ps.thermo_buoy_params()
    assert abs(tb["alpha"] - 0.5235) < 1e-12 and tb["n2"] == 1. and ps.unsupported() == []
    ini = {k: dict(v) for k, v in ws.ini.items()}
    ini.setdefault("diff", {})["swdiff"] = "smag2"
    assert any("buoy" in b for b in CaseConfig(ini).unsupported())
    assert case("andren1994").unsupported() == []                                # neutral LES: calc_evisc_neutral


def test_per_scalar_keys_win():
    """`key[scalar]` overrides (src/boundary.cxx:234-235, src/fields.cxx:412): gabls1 has sbcbot[th]=dirichlet, sbctop[th]=flux."""
    c = case("gabls1")
    assert c.scalars == ["th"] and c.sbcbot["th"] == "dirichlet" and c.sbctop["th"] == "flux"
    p = c.make_params()
    assert p.sbcbot[0] == capi.BC_DIRICHLET and p.sbctop[0] == capi.BC_NEUMANN
    assert c.svisc["th"] == 1e-5
    # neither form present -> required, as in the reference
    ini = {k: dict(v) for k, v in c.ini.items()}
    del ini["boundary"]["sbcbot[th]"]
    with pytest.raises(KeyError):
        CaseConfig(ini)


def test_ini_syntax(tmp_path):
    p = tmp_path / "x.ini"
    p.write_text("[grid]\nitot=8 # comment\n\n# full-line comment\njtot = 4\n[fields]\nrndamp[th]=0.1\n")
    ini = read_ini(str(p))
    assert ini["grid"] == {"itot": "8", "jtot": "4"} and ini["fields"]["rndamp[th]"] == "0.1"


def test_every_shipped_case_is_classified():
    """The table of DESIGN.md section 8: which shipped cases lie inside the accelerated path, and the one switch that keeps each
    of the others out."""
    import glob
    expect_out = {"bomex": ["mbcbot/mbctop=ustar/freeslip"], "conservation": ["swdiff=0"], "dycoms": ["swmicro=2mom_warm"],
                  "rcemip": ["swmicro=nsw6"], "rico": ["swmicro=2mom_warm"], "rico_radiation": ["swmicro=2mom_warm"],
                  "weisman_klemp": ["swmicro=nsw6"]}
    seen = 0
    for d in sorted(glob.glob(os.path.join(REF, "*"))):
        name = os.path.basename(d)
        f = os.path.join(d, name + ".ini")
        if not os.path.exists(f):
            continue
        seen += 1
        c = CaseConfig.from_file(f)
        assert c.unsupported() == expect_out.get(name, []), (name, c.unsupported())
        if name not in expect_out:
            p = c.make_params()
            assert p.swthermo == {"0": 0, "dry": 1, "buoy": 2, "moist": 3}[c.swthermo]
    assert seen >= 25
