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


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("name,world", [("multi_files", 2), ("k63_multi", 2), ("k21C", 2), ("multi_files", 4)])
def test_sharded_count_matches_golden(name, world, built, workdir, inputs):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    from jellyfish_b200.distributed import concat_shards
    args, ins = CASES[name]
    opt = dict(zip(args[0::2], args[1::2])) if "-C" not in args else None
    k = int(args[args.index("-m") + 1])
    v = args[args.index("-s") + 1]
    size = int(v[:-1]) * {"k": 10**3, "M": 10**6, "G": 10**9}[v[-1]] if v[-1] in "kMG" else int(v)
    out = os.path.join(workdir, "multi_%s_%d" % (name, world))
    cfg = {"size": size, "k": k, "canonical": "-C" in args, "files": [inputs[i] for i in ins], "out": out, "batch_bytes": 300000}
    worker = os.path.join(os.path.dirname(__file__), "multi_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29641", worker, json.dumps(cfg)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600, env=dict(os.environ, SOURCE_DATE_EPOCH="0"))
    assert r.returncode == 0, r.stdout.decode()[-3000:]
    db = concat_shards(out, world, out + ".jf")
    h, b = jfutil.split_db(db)
    g = GOLDEN[name]
    assert jfutil.semantic(h) == g["header"]
    assert jfutil.md5(b) == g["body_md5"]
