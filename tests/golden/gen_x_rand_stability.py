#!/usr/bin/env python3
"""Extracts the PRNG / x-initial known answers the reference's own regression pins
(test_regress/t/t_x_rand_stability.out, produced by t_x_rand_stability.v with
--x-initial unique and +verilator+rand+reset+2, default seed 1) into
tests/golden/x_rand_stability.json.  Run where /root/reference exists.
Scope/salt strings per SURVEY.md 8c: scope = model instance name ("top" + ".cell" path for
non-inlined modules), salt = prettyName of the variable."""
import json
import os
import re

REF = "/root/reference/test_regress/t/t_x_rand_stability.out"
vals = {}
for line in open(REF):
    m = re.match(r"^(\w+) = 0x([0-9a-f]+)$", line.strip())
    if m and m.group(1) in ("uninitialized", "big"):
        vals[m.group(1)] = m.group(2)
    m = re.match(r"^(top\.t\.the_sub_no_inline_1) no_init 0x([0-9a-f]+)$", line.strip())
    if m:
        vals["no_inline_1"] = m.group(2)
out = [
    {"kind": "I", "bits": 32, "scope": "top", "salt": "t.uninitialized", "hex": vals["uninitialized"]},
    {"kind": "W", "bits": 256, "scope": "top", "salt": "t.big", "hex": vals["big"]},
    {"kind": "Q", "bits": 64, "scope": "top.t.the_sub_no_inline_1", "salt": "no_init",
     "hex": vals["no_inline_1"]},
]
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "x_rand_stability.json")
json.dump({"seed": 1, "rand_reset": 2, "cases": out}, open(path, "w"), indent=1)
print(out)
