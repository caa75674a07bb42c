#!/usr/bin/env python
"""Diagnostic for the corner goldens: header differences and record-multiset equality of our CLI vs the restatement."""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen, jfutil
from cases import EDGE_CASES
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_edge.json")))
with tempfile.TemporaryDirectory() as d:
    inputs = gen.make_all(d)
    for name in sorted(EDGE_CASES):
        args, ins = EDGE_CASES[name]
        ours, ref = os.path.join(d, "o.jf"), os.path.join(d, "r.jf")
        jfutil.run([jfutil.OUR_JF, "count"] + list(args) + ["-o", ours] + [inputs[i] for i in ins])
        jfutil.run([jfutil.ORACLE_C, "count"] + list(args) + ["-o", ref] + [inputs[i] for i in ins])
        h1, b1 = jfutil.split_db(ours); h2, b2 = jfutil.split_db(ref)
        s1, s2 = jfutil.semantic(h1), jfutil.semantic(h2)
        diff = {k: (s1[k], s2[k]) for k in s1 if s1[k] != s2[k] and k not in ("matrix1", "reprobes")}
        if s1["matrix1"] != s2["matrix1"]: diff["matrix1"] = "differs"
        r1, r2 = jfutil.records(h1, b1), jfutil.records(h2, b2)
        print(name, "golden_ok" if jfutil.md5(b2) == golden[name]["body_md5"] else "ORACLE!=GOLDEN", "body_equal" if b1 == b2 else "body_differs",
              "multiset_equal" if sorted(r1) == sorted(r2) else "MULTISET_DIFFERS", "n", len(r1), len(r2), "hdr_diff", diff)
