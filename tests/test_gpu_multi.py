"""Multi-GPU parity (needs >= 2 GPUs, e.g. `gpurun --gpus 2`): the rank-ordered concatenation of
the shard dumps equals the single-GPU / reference database byte for byte."""
import json
import os
import subprocess
import sys

import pytest

import jfutil
from cases import CASES

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


# x17_4M: a table large enough for every shard to be filled region by region, i.e. the RECORD form of the exchange
# (k <= 21, 32-bit slots); no reference golden of that size: the yardstick is the single-GPU engine, held to the goldens
# in test_gpu_parity.py.  The other cases go through the KEY form.
X17 = (["-m", "17", "-s", "4M", "-C"], ["plain1m.fa", "multi.fa", "dos.fa"])


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("name,world", [("multi_files", 2), ("k63_multi", 2), ("k21C", 2), ("x17_4M", 2), ("multi_files", 4), ("x17_4M", 4),
                                        ("multi_files", 8), ("x17_4M", 8)])
def test_sharded_count_matches_golden(name, world, built, workdir, inputs):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    from jellyfish_b200.distributed import concat_shards
    args, ins = CASES[name] if name in CASES else X17
    opt = dict(zip(args[0::2], args[1::2])) if "-C" not in args else None
    k = int(args[args.index("-m") + 1])
    v = args[args.index("-s") + 1]
    size = int(v[:-1]) * {"k": 10**3, "M": 10**6, "G": 10**9}[v[-1]] if v[-1] in "kMG" else int(v)
    out = os.path.join(workdir, "multi_%s_%d" % (name, world))
    cfg = {"size": size, "k": k, "canonical": "-C" in args, "files": [inputs[i] for i in ins], "out": out, "batch_bytes": 300000}
    if name not in CASES:
        cfg["engine"] = {"part_min_mb": 1, "pool_bytes": 4 << 30}      # small shards filled region by region
    worker = os.path.join(os.path.dirname(__file__), "multi_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29641", worker, json.dumps(cfg)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600, env=dict(os.environ, SOURCE_DATE_EPOCH="0"))
    assert r.returncode == 0, r.stdout.decode()[-3000:]
    db = concat_shards(out, world, out + ".jf")
    h, b = jfutil.split_db(db)
    if name in GOLDEN:
        g = GOLDEN[name]
        assert jfutil.semantic(h) == g["header"]
        assert jfutil.md5(b) == g["body_md5"]
    else:
        assert b"records" in r.stdout          # the worker says which form of the exchange it used
        from jellyfish_b200 import HashCounter
        with HashCounter(size, 7, k=k, canonical="-C" in args, allow_regrow=False) as one:
            one.add_files([inputs[i] for i in ins])
            one.done()
            assert jfutil.md5(one.dump_records()) == jfutil.md5(b) and len(b) > 0
            hdr = one.header()
            assert {x: hdr[x] for x in jfutil.SEMANTIC_KEYS} == jfutil.semantic(h)
