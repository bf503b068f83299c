"""Write tests/golden/parity_bounds.json from the numbers the GPU parity tests measured (gpurun_out/parity_r02.json).
Every non-bit-exact parity test asserts `<= 2 x the committed measurement` (tests/conftest.py::parity_bound)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_r02.json")
d = json.load(open(src))
keep = {k: v for k, v in d.items() if k.startswith(("e2e/", "options/"))}
with open(os.path.join(ROOT, "tests", "golden", "parity_bounds.json"), "w") as f:
    json.dump(keep, f, indent=1, sort_keys=True)
print(f"wrote {len(keep)} measured cases")
