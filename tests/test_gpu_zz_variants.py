"""Selectable kernel variants (environment switches read when a context is created) against the oracle:
every A/B candidate must be bit-exact before it is timed.

    MOLOCH_B200_WSOLVE = 6 (default: thread per column, cp.async ring, two sweep arrays in shared memory, the
                            finished divergence recomputed in the upward pass; 7 warps per SM)
                         5 (three sweep arrays in shared memory, 4 warps per SM: the variant profiles/ measured)
                         2 (CTA = 32 columns x all levels, one-warp sweeps)
    MOLOCH_B200_WAF    = 2 (default: field-batched fused WAF kernels) | 1 (one kernel per reference loop nest)

(Sorts after the other GPU test files; the same bodies run on the CPU build of the CUDA sources.)"""
import pytest

import test_gpu_parity as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["limited_area", "tall"])
@pytest.mark.parametrize("impl", ["5", "2"])
def test_wsolve_variants_bit_exact(impl, case, monkeypatch):
    monkeypatch.setenv("MOLOCH_B200_WSOLVE", impl)
    P.test_steps_bit_exact(case)


@pytest.mark.parametrize("case", ["limited_area", "periodic_hills"])
def test_waf_per_loop_kernels_bit_exact(case, monkeypatch):
    monkeypatch.setenv("MOLOCH_B200_WAF", "1")
    P.test_steps_bit_exact(case)
