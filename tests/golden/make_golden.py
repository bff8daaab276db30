"""Generates tests/golden/oracle_golden.json from the oracle (run from the repo
root: python tests/golden/make_golden.py).  The reference itself cannot be
built here (no Fortran compiler), so these vectors pin the oracle, not the
reference: PARITY UNPINNED by the reference's own tests."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_oracle  # noqa: E402

if __name__ == "__main__":
    out = test_oracle.compute_golden()
    with open(test_oracle.GOLD, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", test_oracle.GOLD)
