#!/usr/bin/env python
"""Regenerate tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref/jellyfish).

Run in the build container (where /root/reference exists and `make -C oracle` has been run).
Each case records the semantic header keys and the md5 / length of the record body that
`jellyfish count` writes for a deterministic input of tests/gen.py.
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen      # noqa: E402
import jfutil   # noqa: E402
from cases import BC_CASES, BF_CASES, BIG_CASES, CASES, DISK_CASES, EDGE_CASES, QUAL_CASES  # noqa: E402

os.environ["SOURCE_DATE_EPOCH"] = "0"
with tempfile.TemporaryDirectory() as d:
  files = gen.make_all(d)
  sets = ((BIG_CASES, "golden_big.json"),) if "--big" in sys.argv else ((CASES, "golden.json"), (QUAL_CASES, "golden_qual.json"), (BF_CASES, "golden_bf.json"), (EDGE_CASES, "golden_edge.json"), (DISK_CASES, "golden_disk.json"))
  for cases, target in sets:
    out = {}
    for name, (args, ins) in sorted(cases.items()):
        db = os.path.join(d, name + ".jf")
        jfutil.run([jfutil.REF_JF, "count", "-t", "8" if "--big" in sys.argv else "1"] + jfutil.subst(args, files) + ["-o", db] + [files[i] for i in ins])
        h, b = jfutil.split_db(db)
        out[name] = {"args": args, "inputs": ins, "header": jfutil.semantic(h), "body_md5": jfutil.md5(b), "body_len": len(b)}
        print(name, out[name]["body_md5"], len(b))
    with open(os.path.join(ROOT, "tests", "golden", target), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)

# `jellyfish bc` + `count --bc`
BC_KEYS = ("format", "key_len", "matrix1", "matrix2", "size", "nb_hashes", "canonical")
if "--big" not in sys.argv:
    out = {}
    with tempfile.TemporaryDirectory() as d:
        files = gen.make_all(d)
        for name, (bargs, bins, cargs, cins) in sorted(BC_CASES.items()):
            bc = os.path.join(d, name + ".bc")
            jfutil.run([jfutil.REF_JF, "bc", "-t", "3"] + bargs + ["-o", bc] + [files[i] for i in bins])
            hb, bb = jfutil.split_db(bc)
            db = os.path.join(d, name + ".jf")
            jfutil.run([jfutil.REF_JF, "count", "-t", "3"] + cargs + ["--bc", bc, "-o", db] + [files[i] for i in cins])
            h, b = jfutil.split_db(db)
            out[name] = {"bc_header": {k: hb.get(k) for k in BC_KEYS}, "bc_md5": jfutil.md5(bb), "bc_len": len(bb),
                         "header": jfutil.semantic(h), "body_md5": jfutil.md5(b), "body_len": len(b)}
            print(name, out[name]["bc_md5"], out[name]["body_md5"], len(b))
    with open(os.path.join(ROOT, "tests", "golden", "golden_bc.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
