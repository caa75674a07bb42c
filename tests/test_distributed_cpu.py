"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: bucket -> count exchange ->
uneven all-to-all -> owner-side insertion, and the shard concatenation rule.  The device
kernels are replaced by a tiny numpy backend behind the same `RouteBackend` seam."""
import os
import subprocess
import sys
import textwrap

import jfutil

WORKER = textwrap.dedent('''
    import os, sys, json
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %(root)r)
    from jellyfish_b200.distributed import RouteBackend, exchange_and_insert

    class NumpyBackend(RouteBackend):
        """owner = key %% world ; "table" = python dict"""
        key_words = 1
        def __init__(self, world): self.world = world; self.table = {}
        def extract_route(self, keys_in, begin, end, keys, capacity, counts):
            for d in range(self.world):
                mine = keys_in[keys_in %% self.world == d]
                keys[d, :len(mine)] = torch.from_numpy(mine)
                counts[d] = len(mine)
        def insert_keys(self, keys, n):
            for v in keys[:n].tolist(): self.table[v] = self.table.get(v, 0) + 1

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    rng = np.random.default_rng(100 + rank)
    be = NumpyBackend(world)
    cap = 5000
    send = torch.zeros((world, cap), dtype=torch.int64); recv = torch.zeros((world, cap), dtype=torch.int64)
    counts = torch.zeros(world, dtype=torch.int64)
    all_mine = []
    for rnd in range(3):                      # ragged rounds, one of them empty on rank 1
        n = 0 if (rank == 1 and rnd == 1) else int(rng.integers(1, 4000))
        keys_in = rng.integers(0, 500, size=n).astype(np.int64)
        all_mine.append(keys_in)
        counts.zero_()
        be.extract_route(keys_in, True, True, send, cap, counts)
        exchange_and_insert(be, world, send, counts, cap, recv)
    # every rank tells everyone what it generated; check ownership and totals
    gathered = [None] * world
    dist.all_gather_object(gathered, np.concatenate(all_mine).tolist())
    expect = {}
    for lst in gathered:
        for v in lst:
            if v %% world == rank: expect[v] = expect.get(v, 0) + 1
    assert be.table == expect, "rank %%d: table differs" %% rank
    assert all(k %% world == rank for k in be.table)
    print("OK", rank, sum(be.table.values()))
    dist.destroy_process_group()
''')


def test_exchange_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": jfutil.ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    out = r.stdout.decode()
    assert r.returncode == 0, out[-3000:]
    assert out.count("OK") == 2, out[-3000:]


def test_concat_shards(tmp_path):
    from jellyfish_b200.distributed import concat_shards
    from jellyfish_b200.engine import write_header
    hdr = {"alignment": 8, "key_len": 42, "counter_len": 4, "size": 16}
    bodies = [b"A" * 20, b"", b"C" * 30]
    for r, b in enumerate(bodies):
        with open("%s.%d" % (tmp_path / "db", r), "wb") as f:
            write_header(f, dict(hdr, rank=r))
            f.write(b)
    out = concat_shards(str(tmp_path / "db"), 3, str(tmp_path / "all.jf"))
    h, body = jfutil.split_db(out)
    assert h["rank"] == 0 and body == b"".join(bodies)
