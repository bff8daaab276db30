"""The product's CUDA sources executed on the CPU (TEST INFRASTRUCTURE).

tests/emu_lib.py compiles regcm_b200/csrc/*.cu -- the files nvcc builds for
sm_100a -- with g++ against a SIMT stand-in (threads are fibers; __syncthreads,
warp shuffles and votes are real rendez-vous points; cp.async lands only at its
wait_group; device and shared memory are NaN-poisoned; ranks of a multi-GPU run
are threads whose peer mappings are plain pointers).  These tests then run the
*bodies of the GPU parity tests* (tests/test_gpu_*.py) against that build: same
C ABI, same Python mirror, same oracle, same bit-exact bar.  They do not replace
the GPU tests (no SASS, no real memory model, the division seed is exact instead
of MUFU.RCP64H); they make sure that every kernel, launcher, halo plan and ABI
entry has been executed before it meets a B200 -- in particular code written
when the round's GPU budget was already spent.

A representative subset runs by default (about two minutes); EMU_FULL=1 runs
every case of the GPU test files, forwards and in reverse CTA/thread order.
"""
import os

import pytest

import emu_lib
import util

FULL = os.environ.get("EMU_FULL", "0") == "1"
from regcm_b200.synthetic import small as _small  # noqa: E402


@pytest.fixture(autouse=True)
def _emulated_library(monkeypatch):
    lib = emu_lib.load()
    lib.emu_set_order(0)
    lib.emu_set_jitter(0)
    monkeypatch.setattr(util, "LIB", lib)
    yield lib
    lib.emu_set_order(0)
    lib.emu_set_jitter(0)


def _gpu_tests():
    import test_gpu_multi as M
    import test_gpu_parity as P
    import test_gpu_zbdy as Z
    return P, Z, M


def _pick(cases, default):
    return [c for c in cases if c != "wide"] if FULL else default


P_CASES = ["periodic_flat", "periodic_hills", "limited_area", "band", "rotllr", "no_divdamp", "no_divfilter",
           "no_damp_no_filter", "tall"]
Z_CASES = ["lam_full", "lam_flux_tracers", "lam_no_icbc_condensate", "lam_ipptls1", "band", "lam_tke", "no_sponge",
           "lam_slice", "spectral", "spectral_band"]


@pytest.mark.parametrize("case", _pick(P_CASES, ["periodic_hills", "limited_area"]))
def test_dycore_phases(case):
    _gpu_tests()[0].test_phases_bit_exact(case)


@pytest.mark.parametrize("case", _pick(P_CASES, ["band", "rotllr", "no_damp_no_filter"]))
def test_dycore_steps(case):
    _gpu_tests()[0].test_steps_bit_exact(case)


def test_dycore_steps_reverse_order(_emulated_library):
    """CTAs and threads in the opposite order: a dependence between threads that no barrier orders shows up."""
    _emulated_library.emu_set_order(1)
    _gpu_tests()[0].test_steps_bit_exact("limited_area")


@pytest.mark.skipif(not FULL, reason="EMU_FULL=1: 70 levels, the wide shared-memory variants of the column kernels")
def test_dycore_tall():
    _gpu_tests()[0].test_steps_bit_exact("tall")


def test_wafone_single_field():
    _gpu_tests()[0].test_wafone_single_field()


@pytest.mark.parametrize("case", ["periodic_hills", "limited_area", "limited_area_rotllr", "limited_area_boundary",
                                  "limited_area_spectral", "limited_area_tke"])
def test_reference_golden(case):
    """The emulated CUDA path against the digests of the executed reference source."""
    _gpu_tests()[0].test_reference_golden(case)


@pytest.mark.parametrize("case", ["band_boundary", "no_damp_no_filter", "vapour_only", "limited_area_diag"])
def test_reference_golden_more(case):
    _gpu_tests()[1].test_reference_golden_more(case)


@pytest.mark.parametrize("case", _pick(Z_CASES, ["lam_full", "lam_ipptls1"]))
def test_boundary(case):
    _gpu_tests()[1].test_boundary_bit_exact(case)


@pytest.mark.parametrize("case", _pick(Z_CASES, ["lam_tke", "spectral_band", "lam_slice"]))
def test_steps_with_boundary(case):
    _gpu_tests()[1].test_steps_with_boundary_bit_exact(case)


def test_boundary_reverse_order(_emulated_library):
    _emulated_library.emu_set_order(1)
    _gpu_tests()[1].test_steps_with_boundary_bit_exact("lam_full")


def test_mkslice_massck_tke_misc():
    Z = _gpu_tests()[1]
    Z.test_mkslice()
    Z.test_massck_and_ps_guard()
    Z.test_tke_steps_periodic()
    Z.test_bdy_shift_swaps_buffers()
    Z.test_boundary_needs_configuration()


def _multi_cases():
    M = _gpu_tests()[2]
    keep = None if FULL else {("periodic", "p2p"), ("limited_area_2x2", "p2p"), ("limited_area_2x4", "p2p"),
                              ("band_2x4", "p2p_unfused"), ("limited_area_2x2", "nccl"), ("limited_area_2x2", "p2p_sound"),
                              ("limited_area_1x4", "p2p"), ("limited_area_1x2", "p2p_psignal")}
    out = []
    for tr in ("p2p", "p2p_sound", "p2p_unfused", "nccl", "p2p_psignal", "p2p_nowz"):
        for c in M.CASES:
            if tr in ("p2p_psignal", "p2p_nowz"):
                if c[0] not in M.ROWS_ONLY:
                    continue
            elif tr != "p2p" and c[0] not in ("periodic", "limited_area", "limited_area_2x2", "band_2x4"):
                continue
            if keep is None or (c[0], tr) in keep:
                out.append(pytest.param(*c, tr, id=f"{c[0]}-{tr}"))
    return out


@pytest.mark.parametrize("name,wl,px,py,transport", _multi_cases())
def test_decomposed(name, wl, px, py, transport, monkeypatch):
    """Ranks are threads: the peer-store transport (fused into the sound kernels or as separate rounds) and
    the NCCL path (mailbox) against the single-domain oracle, bit for bit."""
    _gpu_tests()[2].test_decomposed_bit_exact(name, wl, px, py, transport, monkeypatch)


@pytest.mark.parametrize("px,py", [(2, 1), (1, 2), (2, 2)] if FULL else [(2, 2)])
def test_decomposed_boundary(px, py, monkeypatch):
    _gpu_tests()[2].test_decomposed_boundary_bit_exact(px, py, monkeypatch)


@pytest.mark.parametrize("nslabs", [1, 3, 1000])
def test_handoff(nslabs):
    import test_gpu_zz_handoff as Hf
    Hf.test_handoff_matches_field_transfers(nslabs)


def test_handoff_bad_arguments():
    import test_gpu_zz_handoff as Hf
    Hf.test_handoff_rejects_bad_arguments()


def test_bench_e2e_path(_emulated_library):
    """bench.py's end-to-end leg (measure_e2e: pipelined hand-off with its self-check) on the emulated library:
    it must choose the pipelined path, and K of its steps with zero tendencies equal K moloch() steps."""
    import numpy as np
    import bench
    from regcm_b200 import synthetic as S
    from regcm_b200.moloch import MolochB200
    wl = S.small(S.WORKLOADS["cordex25"], 40, 36, 9, ntr=2, nspgx=5)
    ms = []
    for _ in range(2):
        m = MolochB200(wl, lib=_emulated_library).allocate_moloch()
        fields, profiles, boxes = S.model_inputs_local(wl, m.g)
        m.init_moloch(fields, profiles, boxes)
        ms.append(m)
    r = bench.measure_e2e(ms[0], wl, 1)
    assert r["handoff_note"] is None and set(r["ms_per_step_by_handoff"]) == {"sequential", "pipelined"}, r
    assert r["handoff"] in ("pipelined", "sequential")      # the faster of the two (on a GPU: the link decides)
    assert r["d2h_bytes_per_step"] > 0 and r["h2d_bytes_per_step"] > 0
    ms[1].moloch(5)      # self-check step + (warm-up + 1 timed step) per hand-off mode
    for f in ("u", "v", "w", "pai", "t", "qx", "trac"):
        assert np.array_equal(ms[0].get_global(f), ms[1].get_global(f)), f
    for m in ms:
        m.close()


@pytest.mark.parametrize("px,py,transport", [(2, 1, "nccl"), (1, 2, "p2p+nccl"), (2, 2, "p2p+nccl"), (2, 4, "nccl")]
                         if FULL else [(2, 2, "p2p+nccl"), (2, 1, "nccl")])
def test_decomposed_spectral_nudging(px, py, transport):
    """row_reduce / column_reduce (all-gather over the NCCL stand-in + rank-ordered sum) == the decomposed oracle."""
    _gpu_tests()[2].test_decomposed_spectral_nudging_bit_exact(px, py, transport)


def test_spectral_nudging_needs_nccl():
    _gpu_tests()[2].test_spectral_nudging_needs_nccl_for_reductions()


def test_restart_from_the_save_set():
    import test_gpu_zz_handoff as Hf
    Hf.test_restart_from_the_save_set_is_bit_exact()


@pytest.mark.parametrize("name,px,py", [("limited_area_2x2", 2, 2), ("band_2x4", 2, 4), ("limited_area_1x4", 1, 4)] if FULL
                         else [("limited_area_1x4", 1, 4)])
def test_decomposed_with_drifting_ranks(name, px, py, monkeypatch, _emulated_library):
    """The peer-store transport under rank drift: every launch of every rank thread first sleeps a pseudo-random
    time (one launch in 16, up to 20 ms: a rank is at times several kernels behind its neighbours).  The
    protocol must still deliver every ghost cell before its reader and never overwrite one that is still read:
    bit-exact against the single-domain oracle."""
    M = _gpu_tests()[2]
    wl = [c for c in M.CASES if c[0] == name][0][1]
    _emulated_library.emu_set_jitter(20000)
    M.test_decomposed_bit_exact(name, wl, px, py, "p2p", monkeypatch)      # "p2p" = every fusable round fused


def test_bench_main_dry_run(_emulated_library, monkeypatch, capsys, tmp_path):
    """bench.py's whole N=1 flow (model construction with the fusion-level check, timed steps, per-kernel profile,
    roofline, end-to-end leg, CPU baseline, the JSON line) on the emulated library with a stand-in for the few
    torch.cuda calls it makes: catches script errors before the script meets a GPU box."""
    import json
    import sys
    import time
    import types

    import bench
    from regcm_b200 import moloch as M
    from regcm_b200 import synthetic as S

    class _Event:
        def __init__(self, enable_timing=False):
            self.t = 0.0

        def record(self, stream=None):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

    class _Tensor:
        def __init__(self, v):
            self.v = v

        def item(self):
            return self.v[0]

    class _Ctx:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    cuda = types.SimpleNamespace(is_available=lambda: True, set_device=lambda d: None, synchronize=lambda: None,
                                 Stream=lambda: types.SimpleNamespace(cuda_stream=0), Event=_Event,
                                 stream=lambda s: _Ctx())
    torch = types.ModuleType("torch")
    torch.cuda = cuda
    torch.tensor = lambda v, device=None: _Tensor(v)
    torch.device = lambda *a: None
    dist = types.ModuleType("torch.distributed")
    torch.distributed = dist
    monkeypatch.setitem(sys.modules, "torch", torch)
    monkeypatch.setitem(sys.modules, "torch.distributed", dist)
    monkeypatch.setattr(M, "_lib", _emulated_library)          # what load_library() hands to MolochB200
    monkeypatch.setitem(S.WORKLOADS, "tiny", S.small(S.WORKLOADS["cordex25"], 40, 36, 9, ntr=2, nspgx=5))
    monkeypatch.setattr(bench, "cpu_sample_workload", lambda wl: wl)
    # the parity block against a golden file of the same tiny case, made here by the oracle the way
    # scripts/make_bench_golden.py makes the committed one
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_bench_golden", os.path.join(bench.ROOT, "scripts", "make_bench_golden.py"))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    pwl = S.small(S.WORKLOADS["cordex25"], 24, 20, 9, ntr=2, nspgx=5)
    monkeypatch.setattr(bench, "parity_workload", lambda: pwl)
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    os.makedirs(tmp_path / "tests" / "golden")
    o, F, prof = G.oracle_from_host_inputs(pwl)
    o.step(bench.PARITY_STEPS)
    with open(tmp_path / "tests" / "golden" / "bench_parity.json", "w") as f:
        json.dump({"inputs": {n: bench.digest(a) for n, a in {**F, **prof}.items()},
                   "fields": {n: bench.digest(o.get(n)) for n in bench.PARITY_FIELDS}}, f)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "tiny", "--steps", "2", "--warmup", "1"])
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.setenv("MOLOCH_B200_FUSE_HALO", "2")
    monkeypatch.delenv("MOLOCH_B200_WSOLVE", raising=False)
    monkeypatch.setenv("BENCH_TUNE_REPS", "1")           # one timing per variant is enough on the emulator
    # argparse built its choices from S.WORKLOADS at call time, so "tiny" is accepted
    assert bench.main() == 0
    line = json.loads([x for x in capsys.readouterr().out.splitlines() if x.startswith("{")][-1])
    assert line["metric"] == "MOLOCH dycore cell-updates/s" and line["n_gpus"] == 1 and line["finite"]
    assert line["gpu_launches"] > 0 and line["value"] > 0
    assert line["roofline"]["bound"] == "hbm" and line["roofline"]["frac"] > 0
    assert set(line["e2e"]["ms_per_step_by_handoff"]) == {"sequential", "pipelined"}
    assert line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] > 0
    assert line["parity"]["bit_exact"] is True and line["parity"]["inputs_match_golden"] is True, line["parity"]
    tune = line["config"]["variant_tuning"]["wsolve"]
    # the tiny grid is a "small per-rank grid": variant 2 is among the candidates there
    assert tune["new_bit_exact_vs_v5"] is True and tune["candidates"] == [5, 8, 9, 12, 2], tune
    assert line["config"]["wsolve_variant"] in (5, 8, 9, 12, 2) and set(tune["ms_per_step"]) == {"5", "8", "9", "12", "2"}
    lg = line["e2e"]["link_gbs"]     # None only when the hand-off window is not positive (emulator timing noise; never on a GPU)
    assert (lg is None and line["e2e"]["ms_per_step"] <= line["ms_per_step"]) or lg["d2h"] > 0


def test_halo_timeout_is_reported(_emulated_library):
    """Failure detection of the peer-store transport: a rank whose neighbour never arrives gives up its wait after
    the timeout, every later wait gives up at once, and moloch_b200_sync reports the round (RegCM's fatal())."""
    import threading

    from regcm_b200 import synthetic as S
    from regcm_b200.moloch import MolochError
    from multirank import MultiRank
    from util import make_oracle, oracle_inputs
    wl = S.small(S.WORKLOADS["cordex25"], 40, 36, 8, ntr=1, nspgx=5, mo_nsound=2)
    o, _ = make_oracle(wl)
    fields, profiles = oracle_inputs(o, wl)
    mr = MultiRank(wl, 2, 1, fields, profiles, transport="p2p")
    try:
        err = []

        mr.ranks[0].set_option("halo_timeout_ms", 1500)   # the library default is 30 s

        def lonely():
            try:
                mr.ranks[0].moloch(1)      # rank 1 never steps
                mr.ranks[0].sync()
            except MolochError as e:
                err.append(str(e))
        t = threading.Thread(target=lonely)
        t.start()
        t.join(timeout=120)
        assert not t.is_alive(), "the lonely rank is still waiting"
        assert err and "timed out in round" in err[0], err
    finally:
        mr.close()


def test_full_size_property_checks_at_reduced_size():
    """The bodies of tests/test_gpu_zz_fullsize.py (oracle parity on the doubly periodic `ideal` shape; tracer
    homogeneity, copy equality and run-to-run reproducibility) at sizes the emulated build finishes in seconds."""
    import test_gpu_zz_fullsize as Fz
    from regcm_b200 import synthetic as S
    Fz.check_oracle_parity(S.small(S.WORKLOADS["ideal"], 36, 20, 12), 2)
    Fz.check_tracer_properties(S.small(S.WORKLOADS["cordex25"], 40, 36, 9, ntr=3, nspgx=5), 2)


@pytest.mark.parametrize("impl,case", [("6", "limited_area"), ("6", "tall"), ("2", "limited_area"), ("8", "tall"),
                                       ("9", "tall"), ("10", "limited_area"), ("8", "limited_area"), ("9", "limited_area")]
                         if FULL else [("6", "limited_area"), ("8", "limited_area"), ("9", "tall"), ("11", "limited_area")])
def test_wsolve_variants(impl, case, monkeypatch):
    import test_gpu_zz_variants as V
    V.test_wsolve_variants_bit_exact(impl, case, monkeypatch)


@pytest.mark.parametrize("skip", ["1", "0"])
def test_waf_zero_field_skip(skip, monkeypatch):
    import test_gpu_zz_variants as V
    monkeypatch.setattr(V.P.S, "small", lambda wl, jx, iy, kz, **kw: _small(wl, min(jx, 72), min(iy, 40), min(kz, 9), **kw))
    V.test_waf_zero_field_skip_is_bit_identical(skip, monkeypatch)


def test_waf_per_loop_kernels(monkeypatch):
    import test_gpu_zz_variants as V
    V.test_waf_per_loop_kernels_bit_exact("limited_area", monkeypatch)


def test_set_option():
    import test_gpu_zz_variants as V
    V.test_set_option_switches_variants_of_a_live_context()


def test_graft_entry_smoke(_emulated_library, monkeypatch, capsys):
    """__graft_entry__.smoke() -- what the driver runs on the B200 before the bench -- on the emulated library."""
    import __graft_entry__ as G
    from regcm_b200 import moloch as M
    monkeypatch.setattr(M, "_lib", _emulated_library)
    G.smoke()
    assert "smoke ok" in capsys.readouterr().out
