#!/usr/bin/env python
"""What the GPUs of this box can do to each other: peer access, one-process cross-device copy bandwidth, and (under torchrun)
NCCL send/recv and all-to-all bandwidth.  `python scripts/diag/p2p_probe.py` / `torchrun --nproc-per-node 2 scripts/diag/p2p_probe.py`."""
import os
import time

import torch

world = int(os.environ.get("WORLD_SIZE", "1"))
if world == 1:
    n = torch.cuda.device_count()
    print("devices", n, [torch.cuda.get_device_name(i) for i in range(n)])
    for i in range(n):
        print("peer access from", i, [torch.cuda.can_device_access_peer(i, j) if i != j else None for j in range(n)])
    if n >= 2:
        a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:0")
        b = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:1")
        for _ in range(2):
            b.copy_(a)
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        t0 = time.perf_counter()
        for _ in range(8):
            b.copy_(a, non_blocking=True)
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        dt = time.perf_counter() - t0
        print("one process, cuda:0 -> cuda:1 copy: %.1f GB/s" % (8 * a.numel() / dt / 1e9))
else:
    import torch.distributed as dist
    rank = int(os.environ["RANK"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nb = 1 << 30
    send = torch.empty(world * nb, dtype=torch.uint8, device="cuda")
    recv = torch.empty(world * nb, dtype=torch.uint8, device="cuda")
    for form in ("single", "list", "list_no_self"):
        for it in range(3):
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            if form == "single":
                dist.all_to_all_single(recv, send)
            else:
                outs = [recv[i * nb:(i + 1) * nb] for i in range(world)]
                ins = [send[i * nb:(i + 1) * nb] for i in range(world)]
                if form == "list_no_self":
                    outs[rank] = outs[rank][:0]; ins[rank] = ins[rank][:0]
                dist.all_to_all(outs, ins)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        if rank == 0:
            print("NCCL all_to_all (%s), %d ranks, 1 GiB per peer: %.1f ms -> %.1f GB/s per GPU to the others" % (form, world, dt * 1e3, (world - 1) * nb / dt / 1e9))
    dist.destroy_process_group()
