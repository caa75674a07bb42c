#!/bin/bash
# Round 2, step m (2 GPUs): what limits the exchange -- peer access / NVLink probe, NCCL transport, stage times per round.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2m_topo.txt 2>&1
timeout 120 python scripts/diag/p2p_probe.py > gpurun_out/r2m_p2p_single.txt 2>&1
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P,SHM,NET timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 scripts/diag/p2p_probe.py > gpurun_out/r2m_p2p_nccl.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2m_bench_n2.txt 2>&1
cat gpurun_out/r2m_topo.txt | head -12; cat gpurun_out/r2m_p2p_single.txt; grep -E "via|all_to_all|P2P|SHM" gpurun_out/r2m_p2p_nccl.txt | head -30; tail -c 700 gpurun_out/r2m_bench_n2.txt
