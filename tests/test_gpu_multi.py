"""Multi-GPU parity (needs >= 2 B200s: run with `gpurun --gpus 2|4|8`): the
2-D decomposed CUDA run, halos over NCCL, against the single-domain oracle.
Decomposition invariance is bit-exact by construction (no reductions in the
dycore)."""
import numpy as np
import pytest

from regcm_b200 import synthetic as S
from regcm_b200.moloch import load_library

from multirank import MultiRank
from util import DIAGNOSTIC, PROGNOSTIC, bdy_tables_from_oracle, make_oracle, make_oracle_bdy, oracle_inputs

pytestmark = pytest.mark.gpu


def ndev():
    import util
    return (util.LIB if util.LIB is not None else load_library()).moloch_b200_device_count()


CASES = [
    ("periodic", S.small(S.WORKLOADS["isc24_small"], 40, 24, 12, oro="sine", oro_h=600.0), 2, 1),
    ("limited_area", S.small(S.WORKLOADS["cordex25"], 44, 40, 14, ntr=3, nspgx=6), 2, 1),
    ("limited_area_1x2", S.small(S.WORKLOADS["cordex25"], 44, 40, 14, ntr=3, nspgx=6), 1, 2),
    ("periodic_2x2", S.small(S.WORKLOADS["isc24_small"], 40, 24, 12, oro="sine", oro_h=600.0), 2, 2),
    ("limited_area_2x2", S.small(S.WORKLOADS["cordex25"], 45, 41, 14, ntr=3, nspgx=6), 2, 2),
    ("band_2x4", S.small(S.WORKLOADS["cordex25"], 46, 50, 11, ntr=1, nspgx=5, i_band=1, oro="sine"), 2, 4),
    ("limited_area_2x4", S.small(S.WORKLOADS["cordex25"], 47, 53, 12, ntr=2, nspgx=6), 2, 4),
    # the bench's decomposition (rows only): interior ranks with two neighbours; every round of the step that can
    # be fused is (sound loop, advection's u,v / ux,vx, wz between the two WAF kernels)
    ("limited_area_1x4", S.small(S.WORKLOADS["cordex25"], 44, 57, 14, ntr=3, nspgx=6), 1, 4),
]
ROWS_ONLY = ("limited_area_1x2", "limited_area_1x4")


@pytest.mark.parametrize("transport", ["p2p", "p2p_sound", "p2p_unfused", "nccl", "p2p_psignal", "p2p_nowz"])
@pytest.mark.parametrize("name,wl,px,py", CASES, ids=[c[0] for c in CASES])
def test_decomposed_bit_exact(name, wl, px, py, transport, monkeypatch):
    if ndev() < px * py:
        pytest.skip(f"needs {px * py} GPUs")
    if transport in ("p2p_psignal", "p2p_nowz"):
        if name not in ROWS_ONLY:
            pytest.skip("the signalling side and the wz fusion are A/B-tested on the rows-only decompositions")
    elif transport != "p2p" and name not in ("periodic", "limited_area", "limited_area_2x2", "band_2x4"):
        pytest.skip("NCCL and the unfused peer transport are covered by representative cases")
    # p2p: every exchange that has a producer and a consumer kernel fused into them (sound loop, advection's
    # u,v / ux,vx); p2p_sound: sub-steps 2.. of the sound loop only; p2p_unfused: one exchange launch per
    # round (MOLOCH_B200_FUSE_HALO is read when the context is created)
    monkeypatch.setenv("MOLOCH_B200_FUSE_HALO", {"p2p_unfused": "0", "p2p_sound": "1"}.get(transport, "2"))
    # p2p_psignal: fused rounds signalled by the producer's last edge CTA instead of the consumer's first (an option
    # that measured slower on 8 GPUs and is off by default);
    # p2p_nowz: wz between the two WAF kernels travels in a stand-alone round
    monkeypatch.setenv("MOLOCH_B200_PSIGNAL", "1" if transport == "p2p_psignal" else "0")
    monkeypatch.setenv("MOLOCH_B200_FUSE_WZ", "0" if transport == "p2p_nowz" else "1")
    transport = "p2p" if transport.startswith("p2p") else transport
    o, _ = make_oracle(wl)
    fields, profiles = oracle_inputs(o, wl)
    mr = MultiRank(wl, px, py, fields, profiles, transport=transport)
    try:
        for n in (1, 3):
            o.step(n)
            mr.call("moloch", n)
            bad = []
            for f in PROGNOSTIC + (["trac"] if wl.ntr else []):
                a, b = o.get(f), mr.get_global(f)
                if not np.array_equal(a, b):
                    d = np.abs(a - b)
                    bad.append(f"{f}: {np.count_nonzero(d)} cells differ, max {d.max():.3e} at "
                               f"{np.unravel_index(d.argmax(), d.shape)}")
            for f in DIAGNOSTIC:
                a, b = o.get(f), mr.get_global(f)
                r = np.abs(a - b) / np.maximum(np.abs(a), 1e-300)
                if r.max() > 1e-13:
                    bad.append(f"{f}: rel {r.max():.3e}")
            assert not bad, f"after {n} more steps ({px}x{py}):\n" + "\n".join(bad)
    finally:
        mr.close()


BDY = S.small(S.WORKLOADS["cordex25"], 46, 42, 12, ntr=2, nspgx=6, do_bdy=1, present_qc=1, present_qi=1,
              mo_top_nudge=1, ichebdy=1, do_slice=1)


@pytest.mark.parametrize("px,py", [(2, 1), (1, 2), (2, 2)])
def test_decomposed_boundary_bit_exact(px, py, monkeypatch):
    """moloch() with the lateral boundary and mkslice on px x py GPUs: bdyval, the
    relaxation and mkslice are rank-local; `boundary` adds one u/v halo round."""
    if ndev() < px * py:
        pytest.skip(f"needs {px * py} GPUs")
    monkeypatch.setenv("MOLOCH_B200_FUSE_HALO", "2")     # every fusable round fused
    wl = BDY
    o, B = make_oracle_bdy(wl)
    fields, profiles = oracle_inputs(o, wl)
    fields["zetaf"] = o.get("zetaf")
    fields["xlat"] = o.get("xlat")
    mr = MultiRank(wl, px, py, fields, profiles, bdy=bdy_tables_from_oracle(wl, o), boundary=B)
    try:
        o.step(3)
        mr.call("moloch", 3)
        bad = [f for f in PROGNOSTIC + ["trac"] if not np.array_equal(o.get(f), mr.get_global(f))]
        assert not bad, bad
        for f in ("pf3d", "th3d", "ps"):
            a, b = o.get(f), mr.get_global(f)
            assert (np.abs(a - b) <= 1e-13 * np.abs(a)).all(), f
    finally:
        mr.close()


SPEC = S.small(S.WORKLOADS["cordex25"], 48, 40, 8, ntr=2, nspgx=6, do_bdy=1, present_qc=1, present_qi=1, mo_top_nudge=1,
               ichebdy=1, mo_spectral_nudge=1, ds_km=100.0, dtrad=150.0, dt=150.0)


@pytest.mark.parametrize("px,py,transport", [(2, 1, "nccl"), (1, 2, "p2p+nccl"), (2, 2, "p2p+nccl"), (2, 4, "nccl")])
def test_decomposed_spectral_nudging_bit_exact(px, py, transport):
    """mospectral_nudge on px x py ranks: row_reduce / column_reduce as an all-gather over NCCL with the partial
    sums added in rank order -- what the oracle's emulation of the two MPI_Allreduce calls does, so the
    decomposed CUDA run equals the equally decomposed oracle bit for bit (a different decomposition differs
    in the last bits: the reductions are the one place where the dycore's result depends on it)."""
    if ndev() < px * py:
        pytest.skip(f"needs {px * py} GPUs")
    wl = SPEC
    o, B = make_oracle_bdy(wl, px=px, py=py)
    fields, profiles = oracle_inputs(o, wl)
    mr = MultiRank(wl, px, py, fields, profiles, transport=transport, bdy=bdy_tables_from_oracle(wl, o), boundary=B,
                   xbctime=o.get_xbctime())
    try:
        o.step(3)
        mr.call("moloch", 3)
        bad = [f for f in PROGNOSTIC + ["trac"] if not np.array_equal(o.get(f), mr.get_global(f))]
        assert not bad, bad
    finally:
        mr.close()


def test_spectral_nudging_needs_nccl_for_reductions():
    if ndev() < 2:
        pytest.skip("needs 2 GPUs")
    wl = SPEC
    o, B = make_oracle_bdy(wl, px=2, py=1)
    fields, profiles = oracle_inputs(o, wl)
    mr = MultiRank(wl, 2, 1, fields, profiles, transport="p2p", bdy=bdy_tables_from_oracle(wl, o), boundary=B,
                   xbctime=o.get_xbctime())
    try:
        with pytest.raises(RuntimeError, match="needs the NCCL communicator"):
            mr.call("moloch", 1)
    finally:
        mr.close()
