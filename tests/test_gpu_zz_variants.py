"""Selectable kernel variants (environment switches read when a context is created) against the oracle:
every A/B candidate must be bit-exact before it is timed.

    MOLOCH_B200_WSOLVE = 5 (default: thread per column, cp.async ring, three sweep arrays in shared memory, 4 warps
                            per SM: the variant profiles/ measured)
                         6 (two sweep arrays, the finished divergence recomputed in the upward pass; 7 warps per SM;
                            bench.py times the variants and keeps the fastest), 7 (as 6 with a 6-deep ring: 6 warps)
                         2 (CTA = 32 columns x all levels, one-warp sweeps)
                         8, 9, 10 (round 2: row tiles, 16-byte cp.async.cg ring, re-partitioned ring in the upward
                            pass; 8 = three sweep arrays + ring of 6, 9 = two + ring of 9, 10 = two + ring of 12)
                         11, 12 (sweep arrays in tensor memory via tcgen05.st/ld, eight warps per SM; ring of 8 / 6;
                            kz <= 41, taller grids run variant 8)
    MOLOCH_B200_WAF    = 2 (default: field-batched fused WAF kernels) | 1 (one kernel per reference loop nest)

(Sorts after the other GPU test files; the same bodies run on the CPU build of the CUDA sources.)"""
import pytest

import test_gpu_parity as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["limited_area", "tall"])
@pytest.mark.parametrize("impl", ["6", "7", "2", "8", "9", "10", "11", "12"])
def test_wsolve_variants_bit_exact(impl, case, monkeypatch):
    monkeypatch.setenv("MOLOCH_B200_WSOLVE", impl)
    P.test_steps_bit_exact(case)


@pytest.mark.parametrize("case", ["limited_area", "periodic_hills"])
def test_waf_per_loop_kernels_bit_exact(case, monkeypatch):
    monkeypatch.setenv("MOLOCH_B200_WAF", "1")
    P.test_steps_bit_exact(case)


@pytest.mark.parametrize("skip", ["1", "0"])
def test_waf_zero_field_skip_is_bit_identical(skip, monkeypatch):
    """The fused WAF kernels do not advect a field that is exactly +0 in the whole window of a CTA / warp
    (MOLOCH_B200_WAF_ZEROSKIP, default on).  A tracer that is zero except for a blob exercises the skipped,
    the computed and the mixed windows; the result must equal the oracle's BYTE for byte (the sign of a zero
    included), with the skip and without it."""
    import numpy as np
    from util import make_gpu, make_oracle, oracle_inputs
    monkeypatch.setenv("MOLOCH_B200_WAF_ZEROSKIP", skip)
    wl = P.S.small(P.S.WORKLOADS["cordex25"], 132, 70, 14, ntr=3, nspgx=6)
    o, _ = make_oracle(wl)
    fields, profiles = oracle_inputs(o, wl)
    tr = np.array(fields["trac"])
    tr[1] = 0.0
    j1, j2 = wl.jx // 3, wl.jx // 3 + 18
    tr[1, 3:8, wl.iy // 3:wl.iy // 3 + 9, j1:j2] = 1.0e-6 * (1.0 + np.arange(j2 - j1)[None, None, :] * 0.01)
    tr[2] = 0.0                       # identically zero: skipped everywhere
    fields = dict(fields, trac=tr)
    o.set("trac", tr)
    m = make_gpu(wl, fields, profiles)
    o.step(3); m.moloch(3)
    for f in ("trac", "qx", "tetav", "pai"):
        a, b = np.ascontiguousarray(o.get(f)), np.ascontiguousarray(m.get_global(f))
        assert a.tobytes() == b.tobytes(), f"{f} differs (zero-field skip = {skip})"
    assert (m.get_global("trac")[2] == 0.0).all() and not np.signbit(m.get_global("trac")[2]).any()
    m.close()


def test_set_option_switches_variants_of_a_live_context():
    """moloch_b200_set_option: the variants can be switched between steps of one context (what bench.py's
    autotuning does) and the run stays bit-exact; unknown names and values are refused."""
    import numpy as np
    from regcm_b200.moloch import MolochError
    from util import PROGNOSTIC, compare, make_gpu, make_oracle, oracle_inputs
    wl = P.CASES["limited_area"]
    o, _ = make_oracle(wl)
    fields, profiles = oracle_inputs(o, wl)
    m = make_gpu(wl, fields, profiles)
    # (the halo options are accepted on one rank too: they only change what a decomposed run does)
    for opt, v in (("wsolve", 6), ("waf", 1), ("wsolve", 2), ("waf", 2), ("wsolve", 7), ("wsolve", 5), ("fuse_wz", 0),
                   ("fuse_status", 0), ("halo_psignal", 1), ("wsolve", 12)):
        m.set_option(opt, v)
        o.step(1); m.moloch(1)
        compare(o, m, PROGNOSTIC + ["trac"], label=f"after set_option({opt}, {v}): ")
    with pytest.raises(MolochError, match="unknown option"):
        m.set_option("nonsense", 1)
    with pytest.raises(MolochError, match="wsolve must be"):
        m.set_option("wsolve", 3)
    with pytest.raises(MolochError, match="wsolve must be"):
        m.set_option("wsolve", 14)
    assert np.isfinite(m.get_global("pai")).all()
    m.close()
