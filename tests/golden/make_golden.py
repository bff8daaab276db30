"""Generates tests/golden/oracle_golden.json from the oracle (run from the repo
root: python tests/golden/make_golden.py): regression vectors of the oracle
itself after 1 and 10 steps.  The vectors that tie the oracle to the REFERENCE
are tests/golden/reference_moloch.json, written by
`python -m oracle.refrun.run_moloch` from runs of the reference's own source
(see tests/test_reference_pin.py)."""
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
