#!/usr/bin/env python
"""Write tests/golden/matrix_golden.json: hash matrices drawn by the UNMODIFIED reference library
(oracle/_ref/ref_matrix = oracle/ref_matrix.cc linked with the reference's own objects), for shapes
on both sides of 30 rows -- random_bits() (reference lib/misc.cc:66-72) assembles a value from
31-bit random() draws placed 30 bits apart, so matrices of more than 30 rows (tables of 2^31 slots
and more) exercise the overlap.  Run in the build container after `make -C oracle`.
"""
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "oracle", "_ref", "ref_matrix")
SHAPES = [(10, 10, 0), (21, 42, 0), (27, 42, 0), (30, 42, 0), (31, 42, 0), (32, 42, 0), (34, 42, 0), (34, 42, 2),
          (34, 62, 1), (40, 62, 0), (45, 126, 0), (60, 128, 0), (62, 126, 2), (19, 42, 2)]
out = []
for r, c, skip in SHAPES:
    cols = [int(x) for x in subprocess.check_output([TOOL, str(r), str(c), str(skip)]).split()]
    assert len(cols) == c
    out.append({"r": r, "c": c, "skip": skip, "columns": cols})
with open(os.path.join(ROOT, "tests", "golden", "matrix_golden.json"), "w") as f:
    json.dump(out, f)
print("wrote %d matrices" % len(out))
