#!/usr/bin/env python
"""Generates tests/golden/bench_parity.json: SHA-256 digests of the prognostic fields after
`steps` MOLOCH steps of bench.py's parity case (a reduced cordex25: same levels, species,
sponge, dt, dx; 64x64 columns), computed by the CPU ORACLE started from exactly the arrays
bench.py hands to the library (regcm_b200.synthetic.model_inputs, the NumPy host stand-in),
plus the digests of those inputs.  bench.py --gpus N runs the same case on its N ranks with
the transport and kernel variants it is about to time, and compares (line key "parity").

    python scripts/make_bench_golden.py          # rewrites the fixture
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from regcm_b200 import synthetic as S  # noqa: E402


def oracle_from_host_inputs(wl):
    """An oracle world whose static fields, tables and initial state are the host model's."""
    F, prof = S.model_inputs(wl)
    o = Oracle(wl)
    o.load_primary(S.make_primary(wl))      # allocates and derives everything once ...
    for n, a in F.items():                  # ... then every array the library receives is overridden
        o.set(n, a)
    for n, v in prof.items():
        o.set(n, v)
    return o, F, prof


def main():
    wl = bench.parity_workload()
    o, F, prof = oracle_from_host_inputs(wl)
    o.step(bench.PARITY_STEPS)
    out = {"what": "bench.py parity case: oracle (oracle/moloch_oracle.cpp) from the host model's inputs",
           "generator": "python scripts/make_bench_golden.py",
           "workload": bench.parity_descriptor(wl), "steps": bench.PARITY_STEPS,
           "inputs": {n: bench.digest(a) for n, a in sorted({**F, **prof}.items())},
           "fields": {n: bench.digest(o.get(n)) for n in bench.PARITY_FIELDS}}
    path = os.path.join(ROOT, "tests", "golden", "bench_parity.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
