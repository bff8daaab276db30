"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the
same seeded inputs.  Prognostic fields must be BIT-EXACT (the kernels keep the
reference's operation order and are built -fmad=false; the oracle is built
-ffp-contract=off).  Diagnostics that go through pow/exp (p, rho, qsat, ps)
are compared to 1e-13 relative (libm vs CUDA math library)."""
import numpy as np
import pytest

from regcm_b200 import synthetic as S

from util import DIAGNOSTIC, PROGNOSTIC, compare, make_gpu, make_oracle, oracle_inputs

pytestmark = pytest.mark.gpu

CASES = {
    "periodic_flat": S.small(S.WORKLOADS["isc24_small"], 40, 24, 12),
    "periodic_hills": S.small(S.WORKLOADS["isc24_small"], 36, 28, 10, oro="sine", oro_h=800.0, msf_amp=0.03,
                              clat=30.0),
    "limited_area": S.small(S.WORKLOADS["cordex25"], 44, 40, 14, ntr=3, nspgx=6),
    "band": S.small(S.WORKLOADS["cordex25"], 40, 32, 11, ntr=1, nspgx=5, i_band=1, oro="sine"),
    "rotllr": S.small(S.WORKLOADS["cordex25"], 38, 30, 9, ntr=2, nspgx=5, lrotllr=1),
    "no_divdamp": S.small(S.WORKLOADS["cordex25"], 40, 36, 10, ntr=1, nspgx=5, mo_divdamp=0),
    "no_divfilter": S.small(S.WORKLOADS["isc24_small"], 36, 28, 10, oro="sine", oro_h=500.0, mo_divfilter=0),
    "no_damp_no_filter": S.small(S.WORKLOADS["cordex25"], 40, 36, 10, ntr=0, nspgx=5, mo_divdamp=0, mo_divfilter=0),
    # wide enough for the 32-column / 256-thread variants of the column kernels
    "wide": S.small(S.WORKLOADS["cordex25"], 420, 340, 8, ntr=1, nspgx=6, mo_nsound=2),
    "tall": S.small(S.WORKLOADS["isc24_small"], 34, 26, 70, oro="sine", oro_h=500.0, mo_nsound=2),
}


@pytest.mark.parametrize("case", list(CASES))
def test_phases_bit_exact(case):
    wl = CASES[case]
    o, _ = make_oracle(wl)
    fields, profiles = oracle_inputs(o, wl)
    if wl.lrotllr:
        profiles["rlat"] = S.make_primary(wl)["rlat"]
    m = make_gpu(wl, fields, profiles)
    o.reset_tendencies(); m.reset_tendencies()
    o.sound(); m.sound()
    compare(o, m, ["u", "v", "w", "pai", "s"], label="sound: ")
    o.advection(); m.advection()
    compare(o, m, ["u", "v", "w", "pai", "tetav", "ux", "vx", "wx", "qx", "trac"], label="advection: ")
    o.sound(); m.sound()
    o.advection(); m.advection()
    compare(o, m, ["u", "v", "w", "pai", "tetav", "qx", "trac"], label="2nd nadv: ")
    m.close()


@pytest.mark.parametrize("case", list(CASES))
def test_steps_bit_exact(case):
    wl = CASES[case]
    o, _ = make_oracle(wl)
    fields, profiles = oracle_inputs(o, wl)
    if wl.lrotllr:
        profiles["rlat"] = S.make_primary(wl)["rlat"]
    m = make_gpu(wl, fields, profiles)
    for n in (1, 4):
        o.step(n); m.moloch(n)
        compare(o, m, PROGNOSTIC + ["trac"], label=f"after {n} more steps: ")
        compare(o, m, DIAGNOSTIC, exact=False, rtol=1e-13, label=f"after {n} more steps: ")
    m.close()


def test_wafone_single_field():
    wl = CASES["limited_area"]
    o, _ = make_oracle(wl)
    fields, profiles = oracle_inputs(o, wl)
    m = make_gpu(wl, fields, profiles)
    o.reset_tendencies(); m.reset_tendencies()
    o.sound(); m.sound()
    for f, n in (("tetav", 0), ("qx", 1), ("trac", 2)):
        o.wafone(f, max(n, 1)); m.wafone(f, n)
        compare(o, m, [f, "wz"], label=f"wafone({f}): ")
    m.close()


def _golden():
    import json
    import os
    from oracle.refrun import run_moloch as R
    here = os.path.dirname(os.path.abspath(__file__))
    return R, json.load(open(os.path.join(here, "golden", "reference_moloch.json")))


@pytest.mark.parametrize("case", ["periodic_hills", "limited_area", "limited_area_rotllr", "limited_area_boundary",
                                  "limited_area_spectral", "limited_area_tke"])
def test_reference_golden(case):
    """The CUDA path against the digests of the reference's OWN source, executed through the mechanical
    translator of oracle/refrun (tests/golden/reference_moloch.json; see tests/test_reference_pin.py):
    prognostic fields bit for bit (SHA-256 of the bytes), pow/exp-based diagnostics through their sum."""
    from util import make_gpu_bdy, make_oracle_bdy
    R, golden = _golden()
    wl, nsteps = R.golden_cases()[case]
    o, B = make_oracle_bdy(wl)          # inputs only: the oracle is not stepped
    m = make_gpu_bdy(wl, o, B)
    m.moloch(nsteps)
    trans = {"p", "rho", "qsat", "ps", "pf3d", "th3d", "rhb3d", "wpx3d", "rhox2d", "tp2d", "th700", "ptrop"}
    for f, want in golden[case]["fields"].items():
        got = R.digest(m.get_global(f))
        if f in trans:
            assert abs(got["sum"] - want["sum"]) <= 1e-12 * abs(want["sum"]), f
        else:
            assert got["sha256"] == want["sha256"], f"{case}: {f} differs from the executed reference source"
    m.close()
