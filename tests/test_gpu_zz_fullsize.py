"""GPU tests at BASELINE.json's full sizes.

* config 1 (`Testing/ideal.in` shape, 500x100x60, doubly periodic): the CUDA path against the CPU oracle,
  bit for bit, after a short integration (the oracle needs about half a second per step there);
* config 3 (cordex25, 400x400x41, F = 20) -- the benchmark configuration -- against the CPU oracle at FULL size,
  bit for bit, after two steps (the oracle needs about a second per step on the GPU box's host cores), and
  through size-independent properties of the path: scaling a tracer by a power of two scales its solution
  exactly (WAF fluxes are homogeneous of degree one in the advected field, the limiter only sees ratios, and
  status_update's clipping at zero is scale-free), an untouched copy of a tracer stays equal to the original,
  all fields stay finite, and a second run from the same inputs reproduces the first bit for bit.

(Sorts after the other GPU test files: written without GPU access; the same bodies run at reduced size on the CPU
build of the CUDA sources, tests/test_emu_full.py.)"""
import numpy as np
import pytest

from regcm_b200 import hostmodel as H
from regcm_b200 import synthetic as S
from regcm_b200.moloch import MolochB200

import util
from util import PROGNOSTIC, compare, make_gpu, make_oracle, oracle_inputs

pytestmark = pytest.mark.gpu


def check_oracle_parity(wl, nsteps):
    o, _ = make_oracle(wl)
    fields, profiles = oracle_inputs(o, wl)
    m = make_gpu(wl, fields, profiles)
    o.step(nsteps)
    m.moloch(nsteps)
    compare(o, m, PROGNOSTIC + (["trac"] if wl.ntr else []), label=f"{wl.name} after {nsteps} steps: ")
    m.close()


def test_config1_ideal_full_size_bit_exact():
    check_oracle_parity(S.WORKLOADS["ideal"], 6)


def test_config3_cordex25_full_size_bit_exact():
    """The benchmark configuration itself (bench.py's default workload), every prognostic field and all ten
    tracers, 400x400x41, against the oracle."""
    check_oracle_parity(S.WORKLOADS["cordex25"], 2)


def check_tracer_properties(wl, nsteps):
    assert wl.ntr >= 3
    runs = []
    for _ in range(2):
        m = MolochB200(wl, lib=util.LIB).allocate_moloch()
        fields, profiles, boxes = S.model_inputs_local(wl, m.g)
        m.init_moloch(fields, profiles, boxes)
        box = H.bounds(m.g, "trac")
        t1 = m.get_local("trac", box, 1)
        m.set_local("trac", t1 * 1024.0, box, 2)       # species 2 = 2**10 x species 1
        m.set_local("trac", t1, box, 3)                # species 3 = species 1
        m.moloch(nsteps)
        out = {n: m.get_local("trac", box, n) for n in (1, 2, 3)}
        out["moved"] = np.array([float(not np.array_equal(out[1], t1))])
        out["pai"], out["u"], out["qv"] = m.get_local("pai"), m.get_local("u"), m.get_local("qx", None, 1)
        runs.append(out)
        m.close()
    a = runs[0]
    assert all(np.isfinite(v).all() for v in a.values())
    assert a["moved"][0] == 1.0, "the tracer did not change at all: nothing was advected"
    assert np.array_equal(a[2], a[1] * 1024.0), "WAF advection is not homogeneous: 1024 x tracer != tracer x 1024"
    assert np.array_equal(a[3], a[1]), "two identical tracers diverged"
    for k in a:
        assert np.array_equal(a[k], runs[1][k]), f"{k}: a second run from the same inputs differs"


def test_config3_full_size_tracer_properties():
    check_tracer_properties(S.WORKLOADS["cordex25"], 3)
