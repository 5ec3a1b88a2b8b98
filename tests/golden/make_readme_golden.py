#!/usr/bin/env python
"""Transcribes the reference's only literal known-answer vector — the 20 values README.md:73-94 prints for
`A \\ [exp(1); zeros(n-1)]` on the exp(z) Taylor system (README.md:50) — into tests/golden/readme_expz.json.
Runs only where /root/reference is mounted (the dev container); the JSON is what the tests read."""
import json, re
from pathlib import Path

src = Path("/root/reference/README.md").read_text().splitlines()
start = next(i for i, ln in enumerate(src) if ln.startswith("julia> A \\ [exp(1)"))
vals = []
for ln in src[start + 2:]:
    ln = ln.strip()
    if not re.fullmatch(r"[0-9.e+-]+", ln):
        break
    vals.append(float(ln))
assert len(vals) == 20, len(vals)
out = {
    "source": "/root/reference/README.md:50,73-94",
    "n": 20, "l": 1, "u": 0, "r": 1,
    "matrix": "bands: diag [1;1:n-1], subdiag -1; fill U=[1;0...], V=ones(1,n); b=[e;0...]",
    "x": vals,
}
Path(__file__).with_name("readme_expz.json").write_text(json.dumps(out, indent=1))
print("wrote", len(vals), "values")
