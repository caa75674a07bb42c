"""torchrun worker of tests/test_gpu_multi.py: every rank counts its own files into its shard."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jellyfish_b200.distributed import ShardedCounter  # noqa: E402

cfg = json.loads(sys.argv[1])
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_MAX_CTAS", "16")      # K1 leaves 16 SMs to the exchange that runs beside it
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sc = ShardedCounter(cfg["size"], 7, k=cfg["k"], canonical=cfg["canonical"], rank=rank, world=world, device=local,
                    batch_bytes=cfg.get("batch_bytes", 1 << 20), **cfg.get("engine", {}))
files = cfg["files"][rank::world]
rounds = torch.tensor([len(files)], device="cuda")
dist.all_reduce(rounds, op=dist.ReduceOp.MAX)
for i in range(int(rounds.item())):
    if i < len(files):
        data = torch.frombuffer(bytearray(open(files[i], "rb").read() or b"\n"), dtype=torch.uint8)
        real = os.path.getsize(files[i])
        buf = torch.zeros(max(16, data.numel() + 256), dtype=torch.uint8, device="cuda")
        buf[:data.numel()] = data.cuda()
        sc.add_device_text(buf.data_ptr(), real)
    else:
        sc.add_device_text(0, 0)
st = sc.done()
sc.dump_shard(cfg["out"])
tot = torch.tensor([st["kmers"], st["inserted"]], dtype=torch.int64, device="cuda")
dist.all_reduce(tot)
if rank == 0:
    print("TOTAL", tot.tolist(), "exchange:", "records" if sc.records is not None else "keys")
dist.destroy_process_group()
