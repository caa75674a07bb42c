#!/bin/bash
cd "$(dirname "$0")/.."
python - <<'PY'
import sys, os, json, subprocess, tempfile
sys.path.insert(0, "tests")
import gen, jfutil
d = tempfile.mkdtemp()
f = gen.make_all(d)
g = json.load(open("tests/golden/golden.json"))["multi_files"]
out = os.path.join(d, "cm.jf")
r = subprocess.run([sys.executable, "-m", "jellyfish_b200.count_multi", "-m", "17", "-s", "1M", "-C", "-o", out] + [f[i] for i in g["inputs"]], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
print(r.stdout.decode()[-500:])
h, b = jfutil.split_db(out)
print("count_multi world=1:", jfutil.semantic(h) == g["header"], jfutil.md5(b) == g["body_md5"])
PY
