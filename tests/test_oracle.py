"""CPU tests of the oracle (test infrastructure).  The reference ships no
tests or golden vectors for the MOLOCH path, so the
oracle is pinned by known-answer properties derived from the reference source
 and by committed golden fixtures generated from it."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle.oracle import Oracle
from regcm_b200 import synthetic as S

from util import make_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FLAT = S.small(S.WORKLOADS["isc24_small"], 24, 20, 10)
HILLS = S.small(S.WORKLOADS["isc24_small"], 28, 22, 10, oro="sine", oro_h=900.0, msf_amp=0.04, clat=35.0)
LAM = S.small(S.WORKLOADS["cordex25"], 30, 26, 12, ntr=2, nspgx=5)


def test_bounds_checked_build_runs_all_cases():
    """Every index the restatement touches lies inside the reference's
    allocation bounds (the checked build aborts otherwise)."""
    for wl in (FLAT, HILLS, LAM, S.small(LAM, 30, 26, 12, lrotllr=1), S.small(LAM, 30, 26, 12, i_band=1)):
        for px, py in ((1, 1), (2, 2)):
            o, _ = make_oracle(wl, px=px, py=py, checked=True)
            o.step(1)
            assert np.isfinite(o.get("pai")).all()


@pytest.mark.parametrize("wl", [S.small(HILLS, 28, 22, 10, msf_amp=0.0), S.small(LAM, 30, 26, 12, lrotllr=1)],
                         ids=["periodic_msf1", "limited_area_rotllr"])
def test_wafone_preserves_constants(wl):
    """wafone(pp == c) returns c: the zdv terms cancel the flux divergence
    (Main/mod_moloch.F90:890-891, 950-952, 979-981).  In the non-ROTLLR branch
    the cancellation is exact only for unit map factors (the zdv terms carry
    rmu/rmv, the fluxes do not, :1006-1009), hence msf = 1 there."""
    o, _ = make_oracle(wl)
    o.reset_tendencies()
    o.sound()
    c = 3.25
    q = np.full_like(o.get("tetav"), c)
    o.set("tetav", q)
    o.wafone("tetav")
    out = o.get("tetav")
    own = o.get("fmz") != 0
    assert np.abs(out[own] - c).max() < 5e-13


def test_resting_atmosphere_stays_at_rest():
    """moloch_static_test1-like state (Main/mod_bdycod.F90:3798-3816): flat
    terrain, no wind, horizontally uniform hydrostatic profile from paicompute:
    u, v stay exactly zero, w stays at round-off, pai/tetav are stationary."""
    wl = S.small(FLAT, 20, 16, 12, u0=0.0, v0=0.0)
    P = S.make_primary(wl)
    P["t"] = np.broadcast_to(P["t"].mean(axis=(1, 2), keepdims=True), P["t"].shape).copy()
    P["u"][...] = 0.0
    P["v"][...] = 0.0
    o = Oracle(wl)
    o.load_primary(P)
    pai0, th0 = o.get("pai").copy(), o.get("tetav").copy()
    o.step(20)
    assert np.abs(o.get("u")).max() == 0.0 and np.abs(o.get("v")).max() == 0.0
    assert np.abs(o.get("w")).max() < 1e-9
    assert np.abs(o.get("pai") - pai0).max() < 1e-12
    assert np.abs(o.get("tetav") - th0).max() / th0.max() < 1e-12


def test_uniform_flow_conserves_tracer_mass():
    """Doubly periodic flat domain, msf = 1, s = 0: the flux form conserves
    sum(pp / fmz) to round-off."""
    wl = S.small(FLAT, 32, 24, 10, ntr=1)
    o, _ = make_oracle(wl)
    o.reset_tendencies()           # s = 0: purely horizontal advection
    o.set("u", np.full_like(o.get("u"), 10.0))   # non-divergent: the zdv terms vanish
    o.set("v", np.full_like(o.get("v"), -3.0))
    fmz = o.get("fmz")
    m0 = (o.get("trac")[0] / fmz).sum()
    for _ in range(6):
        o.wafone("trac", 1)
    m1 = (o.get("trac")[0] / fmz).sum()
    assert abs(m1 - m0) / abs(m0) < 1e-12
    assert not np.array_equal(o.get("trac"), S.make_primary(wl)["trac"])


def test_implicit_w_solves_the_tridiagonal_system():
    """One sound sub-step on a single column: the oracle's Thomas sweeps
    (Main/mod_moloch.F90:634-664) agree with a dense solve of
    -zd w(k+1) + (1+zu+zd) w(k) - zu w(k-1) = zwexpl, w(1)=0, w(kz+1)=-s(kz+1)."""
    wl = S.small(HILLS, 12, 10, 14, mo_nsound=1, mo_divdamp=0, mo_divfilter=0)
    o, _ = make_oracle(wl)
    cpd, rdrcv, egrav = S.cpd, S.rdrcv, S.egrav
    kz = wl.kz
    dts = wl.dt / wl.mo_nadv / wl.mo_nsound
    dtrdz = dts / (wl.mo_ztop / kz)
    zcs2 = dtrdz ** 2 * rdrcv
    pai, th, w0 = o.get("pai"), o.get("tetav"), o.get("w")
    fmz, fmzf, ffilt = o.get("fmz"), o.get("fmzf"), o.get("ffilt")
    o.reset_tendencies()
    o.sound()
    w1, s1 = o.get("w"), o.get("s")
    ji = (5, 4)
    # rebuild zdiv2 (post K7) from the Exner update pai1 = pai*(1 - rdrcv*(zdiv2 + dtrdz*fmz*(w(k)-w(k+1))))
    pai1 = o.get("pai")
    col = lambda a: a[:, ji[0], ji[1]]
    zdiv = (1.0 - col(pai1) / col(pai)) / rdrcv - dtrdz * col(fmz) * (col(w1)[:-1] - col(w1)[1:])
    A = np.zeros((kz + 1, kz + 1)); b = np.zeros(kz + 1)
    A[0, 0] = 1.0
    A[kz, kz] = 1.0; b[kz] = col(w1)[kz]
    tf = 0.5 * (col(th)[:-1] + col(th)[1:])          # tetavf(k), k=2..kz  -> index k-2
    for k in range(2, kz + 1):                        # 1-based level
        kk = k - 1
        tfk = tf[k - 2] - col(w0)[kk] * col(fmzf)[kk] * dtrdz * (col(th)[kk - 1] - col(th)[kk])
        r1 = cpd * tfk * col(fmzf)[kk]
        we = col(w0)[kk] - r1 * dtrdz * (col(pai)[kk - 1] - col(pai)[kk]) - egrav * dts
        we += rdrcv * r1 * dtrdz * (col(pai)[kk - 1] * zdiv[kk - 1] - col(pai)[kk] * zdiv[kk])
        zu = zcs2 * col(fmz)[kk - 1] * r1 * col(pai)[kk - 1] + ffilt[kk]
        zd = zcs2 * col(fmz)[kk] * r1 * col(pai)[kk] + ffilt[kk]
        A[kk, kk] = 1.0 + zu + zd; A[kk, kk + 1] = -zd; A[kk, kk - 1] = -zu; b[kk] = we
    wd = np.linalg.solve(A, b)
    # s was finished with the new w (:728-730); the dense solve reproduces w itself
    assert np.abs(wd - col(w1)).max() < 1e-9 * max(1.0, np.abs(col(w1)).max())
    assert s1.shape[0] == kz + 1


@pytest.mark.parametrize("wl,grids", [(HILLS, [(2, 1), (2, 2), (3, 2)]), (LAM, [(2, 1), (1, 2), (2, 4), (3, 3)])],
                         ids=["periodic", "limited_area"])
def test_decomposition_invariance(wl, grids):
    """1 subdomain == px x py subdomains with emulated exchange_*, bit for bit
    (no reductions inside the dycore; ffilt, a global mean, is shared)."""
    o1, _ = make_oracle(wl)
    ff = o1.get("ffilt")
    o1.step(3)
    for px, py in grids:
        o2, _ = make_oracle(wl, px=px, py=py, ffilt=ff)
        o2.step(3)
        for n in ("u", "v", "w", "pai", "tetav", "t", "qx", "ps"):
            assert np.array_equal(o1.get(n), o2.get(n)), f"{n} differs for {px}x{py}"


def test_numpy_setup_matches_oracle_setup():
    """regcm_b200.synthetic (NumPy stand-in of compute_moloch_static,
    init_moloch, paicompute used by bench.py) agrees with the oracle's own
    restatement of the same set-up code to round-off."""
    for wl in (HILLS, LAM):
        o, _ = make_oracle(wl)
        F, prof = S.model_inputs(wl)
        for n in ("fmz", "fmzf", "rfmzu", "rfmzv", "zeta", "hx", "hy", "coru", "corv", "bdywtu", "bdywtv",
                  "bdywtw", "pai", "tetav", "p", "rho", "qsat"):
            a, b = o.get(n), F[n]
            own = a != 0
            assert (np.abs(a - b)[own] <= 1e-12 * np.abs(a[own])).all(), n
        for n in ("gzitak", "gzitakh", "ffilt", "xkdamp", "xknu"):
            assert np.allclose(o.get(n), prof[n], rtol=1e-12, atol=1e-15), n


# ---- golden fixtures -----------------------------------------------------------
GOLD = os.path.join(HERE, "golden", "oracle_golden.json")


def _digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def golden_cases():
    return {"periodic_hills": HILLS, "limited_area": LAM}


def compute_golden():
    out = {}
    for name, wl in golden_cases().items():
        o, _ = make_oracle(wl)
        out[name] = {"ffilt": o.get("ffilt").tolist(), "steps": {}}
        done = 0
        for n in (1, 10):
            o.step(n - done)
            done = n
            out[name]["steps"][str(n)] = {
                f: {"sha256": _digest(o.get(f)), "sum": float(o.get(f).sum()), "absmax": float(np.abs(o.get(f)).max())}
                for f in ("u", "v", "w", "pai", "tetav", "t", "qx")}
    return out


def test_oracle_reproduces_golden_fixtures():
    """tests/golden/oracle_golden.json was generated by tests/golden/make_golden.py
    from this oracle; it pins the oracle (and, through the GPU parity tests,
    the CUDA path) against silent changes.  Digests are of the exact bits."""
    gold = json.load(open(GOLD))
    now = compute_golden()
    for name in gold:
        for n, fields in gold[name]["steps"].items():
            for f, g in fields.items():
                c = now[name]["steps"][n][f]
                assert abs(c["sum"] - g["sum"]) <= 1e-9 * max(1.0, abs(g["sum"])), (name, n, f)
                assert c["sha256"] == g["sha256"], f"{name} step {n} {f}: bits changed"


def test_oracle_reproduces_bench_parity_golden():
    """tests/golden/bench_parity.json (what `bench.py --gpus N` checks its N ranks against before it times them)
    is the oracle's: regenerate the digests from the host model's inputs and compare, inputs included."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("make_bench_golden", os.path.join(ROOT, "scripts", "make_bench_golden.py"))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    import bench
    gold = json.load(open(os.path.join(HERE, "golden", "bench_parity.json")))
    wl = bench.parity_workload()
    assert gold["workload"] == bench.parity_descriptor(wl) and gold["steps"] == bench.PARITY_STEPS
    o, F, prof = G.oracle_from_host_inputs(wl)
    assert {n: bench.digest(a) for n, a in {**F, **prof}.items()} == gold["inputs"]
    o.step(bench.PARITY_STEPS)
    assert {n: bench.digest(o.get(n)) for n in bench.PARITY_FIELDS} == gold["fields"]
