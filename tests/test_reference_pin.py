"""Pinning the oracle on the reference ITSELF.

The reference cannot be compiled here (no Fortran compiler), so its own source
is executed: oracle/refrun translates the routines of the MOLOCH step --
`moloch`, `sound`, `advection`, `wafone`, `boundary`, `bdyval`, `mkslice`, ...
-- mechanically from /root/reference/Main/*.F90 to Python (statement by
statement, no knowledge of the algorithm) and runs them on one rank.

* `test_reference_source_matches_oracle`: reference source vs hand-written
  oracle from the same initial state, every field BIT FOR BIT (needs
  /root/reference; skipped where it is absent, e.g. on the GPU box).
* `test_oracle_matches_reference_golden`: the digests of those reference runs
  are committed (tests/golden/reference_moloch.json, written by
  `python -m oracle.refrun.run_moloch`); the oracle must reproduce them
  anywhere.  tests/test_gpu_parity.py checks the CUDA path against the same
  digests.
"""
import json
import os

import numpy as np
import pytest

from oracle.refrun import run_moloch as R

from util import make_oracle_bdy

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "reference_moloch.json")))
CASES = R.golden_cases()
# fields whose value goes through pow/exp: compared through their sum, not their bytes, where another
# math library (the CUDA one) or another libm build could be involved
TRANSCENDENTAL = {"p", "rho", "qsat", "ps", "pf3d", "th3d", "rhb3d", "wpx3d", "rhox2d", "tp2d", "th700", "ptrop"}


SETUP = R.setup_cases()


def test_golden_file_covers_every_case():
    assert set(GOLDEN) == set(CASES) | set(SETUP)
    for name, (wl, nsteps) in CASES.items():
        assert GOLDEN[name]["steps"] == nsteps and GOLDEN[name]["grid"] == [wl.jx, wl.iy, wl.kz]
        assert set(GOLDEN[name]["fields"]) == set(R.case_fields(wl))


@pytest.mark.skipif(not R.available(), reason="needs the reference sources under /root/reference")
@pytest.mark.parametrize("case", list(CASES))
def test_reference_source_matches_oracle(case):
    wl, nsteps = CASES[case]
    r, o = R.run_case(wl, nsteps)
    bad = [f for f in R.case_fields(wl) if not np.array_equal(r.get(f), o.get(f))]
    assert not bad, f"oracle differs from the executed reference source in {bad}"
    for f in R.case_fields(wl):      # and the committed digests are those of this run
        assert R.digest(r.get(f))["sha256"] == GOLDEN[case]["fields"][f]["sha256"], f


@pytest.mark.skipif(not R.available(), reason="needs the reference sources under /root/reference")
def test_reference_phases_match_oracle():
    """sound and advection on their own, and the work arrays they leave behind."""
    wl, _ = CASES["limited_area"]
    o, B = make_oracle_bdy(wl)
    r = R.ReferenceRun(wl, o, B)
    r.call("reset_tendencies"); o.reset_tendencies()
    r.call("sound", r.ns["dtsound"]); o.sound()
    for f in ("u", "v", "w", "pai", "s", "zdiv2"):
        assert np.array_equal(r.get(f), o.get(f)), f"sound: {f}"
    r.call("advection", r.ns["dtstepa"]); o.advection()
    for f in ("u", "v", "w", "pai", "tetav", "ux", "vx", "wx", "qx", "trac", "wz", "p0"):
        assert np.array_equal(r.get(f), o.get(f)), f"advection: {f}"


@pytest.mark.skipif(not R.available(), reason="needs the reference sources under /root/reference")
@pytest.mark.parametrize("case", list(SETUP))
def test_reference_setup_matches_oracle(case):
    """The set-up chain -- model_zitaf/h and the metric functions (Share/mod_zita.F90),
    compute_moloch_static (Main/mod_params.F90:3316-3395), init_moloch (Main/mod_moloch.F90:201-308),
    setup_bdywt and paicompute (Main/mod_bdycod.F90) -- executed from the reference source against the
    oracle's restatement of it: every static field and the initial Exner function bit for bit."""
    sr, o = R.run_setup_case(SETUP[case])
    bad = [f for f in R.SETUP_FIELDS if not np.array_equal(sr.get(f), o.get(f))]
    assert not bad, bad
    for f in R.SETUP_FIELDS:
        assert R.digest(sr.get(f))["sha256"] == GOLDEN[case]["fields"][f]["sha256"], f


@pytest.mark.skipif(not R.available(), reason="needs the reference sources under /root/reference")
@pytest.mark.parametrize("band", [0, 1])
def test_reference_boundary_setup_matches(band):
    """setup_boundaries (ba%ibnd of the three staggerings), setup_bdycon's MOLOCH branch with the Lehmann
    coefficients, relax_coefficients and lowpass_init executed from the reference source, against the
    oracle's and the host model's restatements: ibnd, hefc, gmeanz, nztop, tnudge, km, lm, bvx, bvy, cnudge
    bit for bit."""
    from regcm_b200 import synthetic as S
    wl = S.small(S.WORKLOADS["cordex25"], 40, 36, 8, ntr=1, nspgx=6, do_bdy=1, mo_top_nudge=1, mo_spectral_nudge=1,
                 ds_km=100.0, dtrad=150.0, dt=150.0, i_band=band, oro="sine" if band else "gauss")
    o, _ = make_oracle_bdy(wl)
    ns = R.BdySetupRun(wl, o.get("zeta")).run().ns
    T = S.bdycon_setup(wl, o.get("zeta"))
    for which, nm in (("cr", "ba_cr"), ("ud", "ba_ud"), ("vd", "ba_vd")):
        assert np.array_equal(ns[nm].ibnd.a.astype(int), T["ibnd"][which]), nm
    assert (ns["nztop"], ns["km"], ns["lm"]) == (o.get_int("nztop"), o.get_int("km"), o.get_int("lm")) == \
        (T["nztop"], T["km"], T["lm"])
    for n in ("gmeanz", "tnudge", "cnudge"):
        assert np.array_equal(ns[n].a, o.get(n)), n
    km, lm = ns["km"], ns["lm"]
    assert np.array_equal(ns["bvx"].a, np.array(o.get("bvx")).reshape(2 * km, wl.jx))
    assert np.array_equal(ns["bvy"].a, np.array(o.get("bvy")).reshape(2 * lm, wl.iy))
    assert np.array_equal(ns["hefc"].a, S.hefc_lehmann(wl))


@pytest.mark.skipif(not R.available(), reason="needs the reference sources under /root/reference")
@pytest.mark.parametrize("case,px,py", [("limited_area_boundary", 2, 2), ("band_boundary", 3, 1)] + (
    [("periodic_hills", 2, 2), ("limited_area", 1, 3)] if os.environ.get("REF_FULL", "0") == "1" else []))
def test_reference_multirank_matches_oracle(case, px, py):
    """The reference's `moloch` on px x py ranks (one thread per rank), halos through the reference's OWN
    exchange routines (real8_3d_exchange_left_right[_bottom_top], MPI-3 variants,
    Main/mpplib/mod_mppparam.F90:3809-3878, 4257-4309, 4661-4712) on an emulated mpi_neighbor_alltoallv,
    against the single-domain oracle: bit for bit.  Pins the decomposition and halo semantics that the
    oracle's exchange emulation and the CUDA library's halo plan restate."""
    wl, nsteps = CASES[case]
    o, B = make_oracle_bdy(wl)
    mr = R.MultiRankReference(wl, o, B, px, py)
    mr.step(nsteps)
    o.step(nsteps)
    bad = [f for f in R.case_fields(wl) if not np.array_equal(mr.get(f), o.get(f))]
    assert not bad, bad


@pytest.mark.skipif(not R.available(), reason="needs the reference sources under /root/reference")
@pytest.mark.parametrize("px,py", [(2, 2), (3, 1)] if os.environ.get("REF_FULL", "0") == "1" else [(2, 2)])
def test_reference_multirank_spectral_nudging_matches_decomposed_oracle(px, py):
    """mospectral_nudge on px x py ranks with the reference's OWN row_reduce / column_reduce
    (Main/mpplib/mod_mppparam.F90:20618-20664) executed on an emulated mpi_allreduce over the row / column
    communicators, sums in rank order: equal to the equally decomposed oracle bit for bit.  Pins the
    `count = nk*(i2-i1+1)` semantics of the two reductions across ranks (incl. the stale tails on the top row
    of ranks), which the CUDA library's halo_group_sum restates."""
    wl, nsteps = CASES["limited_area_spectral"]
    nsteps = 1      # one step = 3 variables x kz levels x 2 reductions, incl. the short (stale-tail) calls
    o, B = make_oracle_bdy(wl, px=px, py=py)
    # the reference ranks take their inputs -- incl. cnudge/tnudge, whose global means (sumall) depend on the
    # decomposition in the last bits -- from the equally decomposed oracle
    mr = R.MultiRankReference(wl, o, B, px, py)
    mr.step(nsteps)
    o.step(nsteps)
    bad = [f for f in R.case_fields(wl) if not np.array_equal(mr.get(f), o.get(f))]
    assert not bad, bad


@pytest.mark.skipif(not R.available(), reason="needs the reference sources under /root/reference")
@pytest.mark.parametrize("kz,nspgx", [(8, 4), (41, 12), (30, 6)])
def test_reference_exponential_sponge_profile_matches(kz, nspgx):
    """The exponential branch of setup_bdycon (Main/mod_bdycod.F90:536-543) with exponential_nudging
    (Main/mpplib/mod_runparams.F90:611-636, incl. its internal findwhere) and spline1d (Share/mod_spline.F90:387-459)
    executed from source on MOLOCH's sigma/hsigma: the sponge table hefc the workloads use
    (regcm_b200.synthetic.hefc_table, a restatement of those three routines) equals it bit for bit."""
    from regcm_b200 import synthetic as S
    wl = S.small(S.WORKLOADS["cordex25"], 2 * nspgx + 8, 2 * nspgx + 6, kz, ntr=0, nspgx=nspgx, do_bdy=1)
    o, _ = make_oracle_bdy(wl)
    run = R.BdySetupRun(wl, o.get("zeta"), lehmann=False).run()
    assert np.array_equal(run.ns["hefc"].a, S.hefc_table(wl))
    assert np.array_equal(o.get("hefc").reshape(kz, nspgx), S.hefc_table(wl))     # what the oracle relaxes with


@pytest.mark.skipif(not R.available(), reason="needs the reference sources under /root/reference")
def test_reference_runs_use_the_reference_allocation():
    """The arrays of a ReferenceRun have exactly the bounds allocate_atmosphere
    (Main/mod_atm_interface.F90:579-624) and allocate_moloch (Main/mod_moloch.F90:159-199) give when THEY are
    executed from source -- so the bounds checking of those runs is the reference's own allocation -- and the
    host model's table (regcm_b200/hostmodel.py) agrees with them."""
    from regcm_b200 import hostmodel as H
    from regcm_b200 import synthetic as S
    wl = S.small(S.WORKLOADS["cordex25"], 16, 14, 8, ntr=1, nspgx=4, mo_nsound=3, ibltyp=2, idiag=1, ichdiag=1, do_bdy=1)
    o, B = make_oracle_bdy(wl)
    comm = R._Comm(4)
    alias = {"mo_atm%zeta": "z", "mo_atm%qs": "qsat"}
    hname = {"z": "zeta"}
    for rank in range(4):
        r = R.ReferenceRun(wl, o, B, px=2, py=2, rank=rank, comm=comm)
        ref = R.reference_allocation_bounds(r)
        n = 0
        for k, b in ref.items():
            name = alias.get(k, k.replace("mo_atm%", ""))
            a = r.ns.get(name)
            if a is None:
                assert name in ("pf", "dz"), name          # not used by the path
                continue
            assert a.bounds() == b, (rank, k, b, a.bounds())
            hn = hname.get(name, name)
            # (arrays the reference allocates on the interior only cross the ABI on the owned box: skipped)
            if hn in H.ALLOC and hn not in ("tten", "uten", "vten", "qxten", "chiten", "tketen", "bdywtu", "bdywtv",
                                            "bdywtw", "tetavf", "ten0", "qen0", "chiten0"):
                hb = H.bounds(r.g, hn)                      # the host hand-off boxes: same horizontal bounds
                assert (b[0][0], b[0][1], b[1][0], b[1][1]) == hb, (rank, hn, b, hb)
            n += 1
        assert n >= 45


@pytest.mark.skipif(not R.available(), reason="needs the reference sources under /root/reference")
def test_reference_set_nproc_matches_decomp():
    """set_nproc (Main/mpplib/mod_mppparam.F90:1250-1641) and setup_model_indexes
    (Main/mod_atm_interface.F90:182-382) executed from source, rank by rank, against regcm_b200/decomp.py:
    process grid, dot/cross ranges, neighbours, boundary flags, interior and ghost ranges."""
    from regcm_b200.decomp import default_cpus_per_dim, make_geom
    n = 0
    for jx, iy, band, crm in ((400, 400, 0, 0), (47, 53, 0, 0), (46, 50, 1, 0), (40, 24, 1, 1), (1536, 1536, 0, 0),
                              (100, 300, 0, 0), (300, 100, 0, 0)):
        for nproc in (1, 2, 3, 4, 6, 8):
            for r in range(nproc):
                ref = R.reference_set_nproc(jx, iy, 41, band, crm, nproc, r)
                if nproc > 1:
                    assert default_cpus_per_dim(nproc, jx, iy) == tuple(ref["dims"])
                g = make_geom(jx, iy, 41, band, crm, ref["dims"][0], ref["dims"][1], r)
                mine = dict(global_dot_jstart=g.jde1, global_dot_jend=g.jde2, global_dot_istart=g.ide1,
                            global_dot_iend=g.ide2, global_cross_jstart=g.jce1, global_cross_jend=g.jce2,
                            global_cross_istart=g.ice1, global_cross_iend=g.ice2, has_bdyleft=g.bl, has_bdyright=g.br,
                            has_bdybottom=g.bb, has_bdytop=g.bt)
                if nproc > 1:
                    mine.update(left=g.left, right=g.right, bottom=g.bottom, top=g.top)
                assert {k: ref[k] for k in mine} == mine, (jx, iy, band, crm, nproc, r)
                ix = ref["indexes"]
                for k in ("jde1", "jde2", "ide1", "ide2", "jdi1", "jdi2", "idi1", "idi2", "jce1", "jce2", "ice1", "ice2",
                          "jci1", "jci2", "ici1", "ici2"):
                    assert ix[k] == getattr(g, k), (k, jx, iy, nproc, r)
                assert (ix["jce1ga"], ix["jce2ga"], ix["ice1ga"], ix["ice2ga"]) == g.ext("cross", 1, 1)
                assert (ix["jde1gb"], ix["jde2gb"], ix["ide1gb"], ix["ide2gb"]) == g.ext("dot", 2, 2)
                n += 1
    assert n > 150


@pytest.mark.skipif(not R.available(), reason="needs the reference sources under /root/reference")
def test_reference_massck_matches_oracle():
    """massck's atmosphere sums (Main/mod_massck.F90:77-185) executed from source == the oracle's, bit for bit
    (same single running sums); the device's row-wise sums are compared with these to 1e-12."""
    from regcm_b200 import synthetic as S
    wl = S.small(S.WORKLOADS["cordex25"], 20, 18, 8, ntr=1, nspgx=4, do_massck=1)
    o, _ = make_oracle_bdy(wl)
    o.step(1)
    ref = R.reference_massck(wl, o)
    dry, dadv, wat, wadv = o.massck()
    assert (ref["tdrym"], ref["tdadv"], ref["tqmass"], ref["tqadv"]) == (dry, dadv, wat, wadv)
    assert ref["tcrai"] == ref["tncrai"] == ref["tqeva"] == 0.0


@pytest.mark.parametrize("case", list(SETUP))
def test_oracle_setup_matches_reference_golden(case):
    o, _ = make_oracle_bdy(SETUP[case])
    for f, want in GOLDEN[case]["fields"].items():
        got = R.digest(o.get(f))
        if f in ("pai", "coru", "corv", "zeta", "fmz", "fmzf", "zetaf", "rfmzu", "rfmzv", "bdywtu", "bdywtv", "bdywtw",
                 "p", "qsat", "rho", "tvirt", "tetav", "ffilt"):
            # exp/sin/pow inside: bytes on this libm, sums anywhere
            assert abs(got["sum"] - want["sum"]) <= 1e-12 * abs(want["sum"]), f
        else:
            assert got["sha256"] == want["sha256"], f


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_matches_reference_golden(case):
    wl, nsteps = CASES[case]
    o, _ = make_oracle_bdy(wl)
    o.step(nsteps)
    for f, want in GOLDEN[case]["fields"].items():
        got = R.digest(o.get(f))
        if f in TRANSCENDENTAL:
            assert abs(got["sum"] - want["sum"]) <= 1e-12 * abs(want["sum"]), f
        else:
            assert got["sha256"] == want["sha256"], f"{case}: {f} differs from the executed reference source"
