#!/usr/bin/env python
"""Copy the evidence of a round from gpurun_out/ (scratch) into profiles/ (tracked): bench lines, ncu summaries (the
`--page details` text of the full captures), the per-launch list with DRAM traffic, traffic.json, micro-benchmarks, test tails."""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def last_json_line(path):
    try:
        lines = [l for l in open(path) if l.startswith("{")]
        return json.loads(lines[-1]) if lines else None
    except OSError:
        return None


def copy(src, dst):
    if os.path.exists(os.path.join(G, src)):
        shutil.copyfile(os.path.join(G, src), os.path.join(P, dst))
        print("profiles/" + dst)


for src, dst in [("r2p_bench_k21.txt", "bench_k21_n1"), ("r2p_bench_k31.txt", "bench_k31_n1"), ("r2p_bench_k63.txt", "bench_k63_n1"),
                 ("r2p_bench_bf.txt", "bench_bf_n1"), ("r2p_bench_reference.txt", "bench_reference_arm"), ("r2r_bench_n2.txt", "bench_k21_n2")]:
    d = last_json_line(os.path.join(G, src))
    if d:
        json.dump(d, open(os.path.join(P, "%s_%s.json" % (tag, dst)), "w"), indent=1)
        print("profiles/%s_%s.json" % (tag, dst))
for src, dst in [("r2p_k1_details.txt", "k1_extract_kernel_ncu.txt"), ("r2p_insert_details.txt", "k2c_win_insert2_ncu.txt"),
                 ("r2k_scatter_details.txt", "k2b_win_scatter_ncu.txt"), ("r2e_hist_details.txt", "k2a_win_hist_ncu.txt"),
                 ("micro_smem_atomics.txt", "micro_smem_atomics.txt"), ("micro_scatter_store.txt", "micro_scatter_store.txt"),
                 ("r2m_p2p_single.txt", "nvlink_probe.txt"), ("r2p_gpu.txt", "gpu.txt")]:
    copy(src, "%s_%s" % (tag, dst))
for src, dst in [("r2p_pytest.txt", "pytest_gpu_tail.txt"), ("r2r_pytest_multi.txt", "pytest_gpu_multi_n2_tail.txt")]:
    p = os.path.join(G, src)
    if os.path.exists(p):
        open(os.path.join(P, "%s_%s" % (tag, dst)), "w").write("".join(open(p).readlines()[-15:]))
        print("profiles/%s_%s" % (tag, dst))
t = os.path.join(G, "r2p_traffic.csv")
if os.path.exists(t):
    subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "tools", "make_traffic.py"), t, os.path.join(P, "traffic.json"),
                           os.path.join(P, "%s_launches_with_dram_traffic.csv" % tag)])
