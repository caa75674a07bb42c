#!/usr/bin/env python
"""bench.py -- k-mers counted/sec at k=21 on N B200s (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our engine (CUDA, via the C ABI)
  python bench.py --impl reference --gpus N ...            the reference's own CPU path

A "step" = one pass of the hot path (FASTA text -> canonical 21-mers -> GF(2) hash ->
insert/increment) over the whole synthetic input, into a zeroed table.

  value  device-resident input (HBM), timed with CUDA events            [k-mers/s, whole job]
  e2e    the same through the public host API (jfgpu_feed) from pinned HOST memory:
         host->device copies and the device->host read of the result inside the timed region
  roofline      for the fused count kernel, against MEASURED_PEAKS.json's HBM figure
  cpu_baseline  the reference binary (oracle/_ref/jellyfish, all host threads) on a bounded
                sample of the same workload

Workload (N=1, default --config k21): BASELINE configs[1], k=21 canonical, 10 Gbp synthetic FASTA.
configs[1] names a "4 G-entry hash", which cannot hold the ~9.98e9 distinct 21-mers of 10 Gbp iid
sequence: the reference doubles it twice to 2^34 slots.  The bench therefore sizes the table at
its final size, -s 16G (2^34 slots), for both arms; set --size 4G to time the doubling too.
Input is larger than L2 (10 GB text, 68 GB table), so no L2 flush is needed between steps.

The other BASELINE configs are bench lines of their own (`--config`, results under profiles/):
  k31   configs[2]  k=31 canonical, 10 Gbp over 8 GPUs = 1.25 Gbp and 2^31 slots (64-bit) per GPU
  k63   configs[4]  k=63 canonical, 2 Gbp, 2^32 slots of 128 bits
  bf    configs[3]  k=21 with the --bf-size 10G Bloom prefilter in front of the table
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "k-mers counted/sec at k=21"
UNIT = "k-mers/s"

# per-GPU workloads of the BASELINE configs: k, bases, -s (table share of one GPU), Bloom prefilter size
CONFIGS = {
    "k21": {"k": 21, "bases": 10_000_000_000, "size": "16G", "bf": 0},
    "k31": {"k": 31, "bases": 1_250_000_000, "size": "2G", "bf": 0},
    "k63": {"k": 63, "bases": 2_000_000_000, "size": "4G", "bf": 0},
    "bf":  {"k": 21, "bases": 10_000_000_000, "size": "16G", "bf": 10_000_000_000},
}


def workload_config(args, world):
    """The `config` object of the JSON line: identical in both arms (it names the workload, not the engine)."""
    bf = (", --bf-size %d Bloom prefilter" % args.bf_size) if args.bf_size else ""
    return {"workload": "k=%d canonical, %d bp synthetic FASTA per GPU (one record, 70-column lines: the shape generate_sequence writes), "
                        "-s %s per GPU%s" % (args.k, args.bases, args.size, bf),
            "l2": "text and table are far larger than L2 (126 MB): no flush between steps",
            "parallelism": "1 GPU" if world == 1 else "table sharded by the top hash bits over %d GPUs" % world}


def parse_size(s):
    mult = {"k": 10**3, "M": 10**6, "G": 10**9, "T": 10**12}
    return int(s[:-1]) * mult[s[-1]] if s[-1] in mult else int(s)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.index = index
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v == "Active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# the reference's own CPU path on a bounded sample
# ------------------------------------------------------------------------------------------------
def write_sample_fasta(path, n_bases, seed=4242):
    """Same shape as the device generator / generate_sequence: '>read1', 70 bases per line."""
    import numpy as np
    rng = np.random.default_rng(seed)
    with open(path, "wb") as f:
        f.write(b">read1\n")
        left = n_bases
        lut = np.frombuffer(b"ACGT", dtype=np.uint8)
        while left > 0:
            n = min(left, 70 * 400000)
            seq = lut[rng.integers(0, 4, size=n, dtype=np.uint8)]
            full = (n // 70) * 70
            body = np.empty((n // 70, 71), dtype=np.uint8)
            body[:, :70] = seq[:full].reshape(-1, 70)
            body[:, 70] = 10
            f.write(body.tobytes())
            if n > full:
                f.write(seq[full:].tobytes() + b"\n")
            left -= n


def cpu_reference_run(sample_fa, n_bases, k, size, threads, workdir):
    """One timed run of the reference binary (or the C restatement) -> k-mers/s of its Counting phase."""
    import jfutil
    timing = os.path.join(workdir, "timing.txt")
    out = os.path.join(workdir, "ref.jf")
    if os.path.exists(jfutil.REF_JF):
        kind = "reference"
        cmd = [jfutil.REF_JF, "count", "-m", str(k), "-s", str(size), "-t", str(threads), "-C", "--no-write",
               "--timing", timing, "-o", out, sample_fa]
        t0 = time.perf_counter()
        subprocess.check_call(cmd, env=dict(os.environ, SOURCE_DATE_EPOCH="0"))
        wall = time.perf_counter() - t0
        secs = None
        for line in open(timing):
            if line.startswith("Counting"):
                secs = float(line.split()[1])
        secs = secs or wall
    else:
        kind = "port"
        threads = 1
        cmd = [jfutil.ORACLE_C, "count", "-m", str(k), "-s", str(size), "-C", "-o", out, sample_fa]
        t0 = time.perf_counter()
        subprocess.check_call(cmd, stderr=subprocess.DEVNULL)
        secs = time.perf_counter() - t0
    return (n_bases - k + 1) / secs, kind, threads, secs


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import jfutil
    k = 21
    sample = args.cpu_sample_bases
    threads = os.cpu_count() or 1
    have_ref = os.path.exists(jfutil.REF_JF)
    if not have_ref:
        sample = min(sample, 5_000_000)
    size = 1 << max(10, (int(sample / 0.6) - 1).bit_length())   # final load ~0.3-0.6, no doubling
    with tempfile.TemporaryDirectory() as d:
        fa = os.path.join(d, "sample.fa")
        write_sample_fasta(fa, sample)
        with open(fa, "rb") as f:          # warm the page cache
            while f.read(1 << 24):
                pass
        vals = []
        for i in range(args.warmup + args.steps):
            v, kind, cores, secs = cpu_reference_run(fa, sample, k, size, threads, d)
            if i >= args.warmup:
                vals.append((v, secs))
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(s for _, s in vals) / len(vals)
    sample_desc = "%d bp iid FASTA (70-col), jellyfish count -m 21 -s %d -t %d -C, Counting phase of --timing" % (sample, size, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": "k=21 canonical, 10 Gbp synthetic FASTA (bounded CPU sample per step: %s)" % sample_desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bases", type=int, default=10_000_000_000, help="bases of synthetic sequence per GPU")
    ap.add_argument("--size", default="16G", help="-s of the per-GPU table share (global table = N x this)")
    ap.add_argument("--k", type=int, default=21)
    ap.add_argument("--cpu-sample-bases", type=int, default=400_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from jellyfish_b200 import HashCounter, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus:
        if rank == 0:
            sys.stderr.write("bench.py: --gpus %d but WORLD_SIZE=%d; launch with torchrun for N>1\n" % (args.gpus, world))
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()

    k = args.k
    n_bases = args.bases
    nbytes = lib.jfgpu_synth_fasta_bytes(n_bases)
    text = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
    got = C.c_uint64(0)
    rc = lib.jfgpu_synth_fasta_device(local_rank, C.c_void_p(text.data_ptr()), nbytes + 256, n_bases, 1000003 * (rank + 1), C.byref(got), None)
    assert rc == 0, "synthetic FASTA generation failed"
    torch.cuda.synchronize()
    n_text = got.value
    kmers_per_step = n_bases - k + 1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    size = parse_size(args.size)
    if world > 1:
        from jellyfish_b200.distributed import ShardedCounter
        counter = ShardedCounter(size * world, 7, k=k, canonical=True, rank=rank, world=world, device=local_rank)
        hc = counter.hc
    else:
        counter = None
        hc = HashCounter(size, 7, k=k, canonical=True, device=local_rank)
    info = hc.info()

    def one_step_device():
        hc.clear()
        if counter is not None:
            counter.add_device_text(text.data_ptr(), n_text)
        else:
            hc.add_device_text(text.data_ptr(), n_text)
        return hc.done()

    # ---- value: device-resident input, CUDA events on the engine's stream, max over ranks ----
    for _ in range(args.warmup):
        st = one_step_device()
    launches0 = lib.jfgpu_kernel_launches()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    dev_secs, kern_secs, kern_launches, drain_secs = 0.0, 0.0, 0, 0.0
    for _ in range(args.steps):
        st = one_step_device()
        dev_secs += st["seconds_count"]
        kern_secs += st["seconds_count_kernel"]
        kern_launches += st["count_kernel_launches"]
        drain_secs += st["seconds_drain"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = lib.jfgpu_kernel_launches() - launches0
    tot = torch.tensor([st["kmers"], st["inserted"], st["distinct"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot)          # keys are inserted by their owner: only the sums must agree
    assert st["kmers"] == kmers_per_step and tot[0].item() == tot[1].item() == kmers_per_step * world, (st, tot.tolist())
    distinct_total = int(tot[2].item())
    t = torch.tensor([dev_secs, wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_secs, wall = t.tolist()
    # the timed quantity: whole steps (table clear + all kernels), wall clock between synchronised barriers
    ms_per_step = 1e3 * wall / args.steps
    value = kmers_per_step * world * args.steps / wall

    # ---- roofline of the dominant kernel ----
    # direct insertion:      K1 count_kernel does everything; B_alg = 71/70 + 32*p + 32 bytes per k-mer
    # region-by-region mode: K1 (parse/hash/stage records) writes rec bytes per k-mer, K2
    #   (insert_chunks_kernel) reads them and sweeps the table once: rec + 2*table_bytes/kmers
    p_mean = 1.0 + st["reprobes"] / max(1, st["inserted"])
    peak, peak_src = measured_peak()
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "traffic.json")
    tinfo = {}
    if os.path.exists(tfile):
        try:
            tinfo = json.load(open(tfile))
        except Exception:
            tinfo = {}
    kernels = []
    nk = kmers_per_step * args.steps
    if info["part_regions"]:
        rec = info["part_rec_bytes"]
        b1 = 71.0 / 70.0 + rec
        b2 = rec + 2.0 * info["table_bytes"] / kmers_per_step
        kernels.append({"kernel": "count_kernel<1,%d,2,1024> (K1: parse+hash+stage records)" % info["slot_bits"], "seconds": kern_secs / args.steps,
                        "alg_bytes_per_kmer": b1, "achieved": nk * b1 / kern_secs / 1e9 if kern_secs else None, "launches": kern_launches,
                        "traffic": tinfo.get("k1_dram_bytes_per_launch")})
        kernels.append({"kernel": "insert_chunks_kernel<1,%d> (K2: region-by-region insert)" % info["slot_bits"], "seconds": drain_secs / args.steps,
                        "alg_bytes_per_kmer": b2, "achieved": nk * b2 / drain_secs / 1e9 if drain_secs else None,
                        "traffic": tinfo.get("k2_dram_bytes_per_launch")})
    else:
        b_alg = 71.0 / 70.0 + 32.0 * p_mean + 32.0
        kernels.append({"kernel": "count_kernel<1,%d,0,512> (direct insert)" % info["slot_bits"], "seconds": kern_secs / args.steps,
                        "alg_bytes_per_kmer": b_alg, "achieved": nk * b_alg / kern_secs / 1e9 if kern_secs else None, "launches": kern_launches,
                        "traffic": tinfo.get("direct_dram_bytes_per_launch")})
    dom = max(kernels, key=lambda x: x["seconds"] or 0)
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak, "unit": "GB/s",
                "frac": (dom["achieved"] / peak) if dom["achieved"] else None, "traffic": dom.get("traffic"), "peak_source": peak_src,
                "alg_bytes_per_kmer": dom["alg_bytes_per_kmer"], "mean_probes": p_mean, "avg_seconds_per_step": dom["seconds"],
                "note": "HBM traffic of the region-by-region pipeline is near its algorithmic minimum (ncu, profiles/traffic.json); the "
                        "kernels are bound by instruction issue (K1) and by L2 atomic/load throughput (K2), not by HBM", "kernels": kernels}
    # the north star's own yardstick: k-mers/s x (71/70 + 32 p + 32) bytes against the HBM peak, i.e. what a
    # table filled by random HBM accesses would have to move; and the measured random-atomic ceiling of this
    # GPU (scripts/micro/atomics.cu: 20.5 G random 32-bit atomics/s over a 32 GB region)
    b_rand = 71.0 / 70.0 + 32.0 * p_mean + 32.0
    roofline["random_access_model"] = {
        "alg_bytes_per_kmer": b_rand, "achieved": value / world * b_rand / 1e9, "unit": "GB/s", "frac_of_hbm_peak": value / world * b_rand / 1e9 / peak,
        "measured_random_atomic_ceiling_gops": 20.5, "ceiling_kmers_per_s": 20.5e9 / p_mean,
        "vs_random_atomic_ceiling": (value / world) / (20.5e9 / p_mean)}
    if info["part_regions"] and drain_secs:
        # K2 against the measured L2 ceiling for its operation mix (scripts/micro/mix.cu: ~50 G keys/s at load 0.58)
        roofline["l2_op_mix"] = {"achieved_gkeys_per_s": nk / drain_secs / 1e9, "measured_ceiling_gkeys_per_s": 50.0,
                                 "frac": nk / drain_secs / 1e9 / 50.0}

    # ---- e2e: the public host API with pinned HOST buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        hptr = lib.jfgpu_host_alloc(n_text)
        assert hptr, "pinned host allocation failed"
        torch.cuda.synchronize()
        # fill the host buffer once (untimed set-up): device -> pinned host
        host = torch.frombuffer((C.c_uint8 * n_text).from_address(hptr), dtype=torch.uint8)
        host.copy_(text[:n_text])
        torch.cuda.synchronize()

        def one_step_host():
            hc.clear()
            if counter is not None:
                counter.add_host_text(hptr, n_text)
            else:
                hc.add_text((C.c_void_p(hptr), n_text))
            return hc.done()          # reads the statistics block back to the host

        for _ in range(min(args.warmup, 1)):
            one_step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st2 = one_step_host()
        barrier()
        wall2 = time.perf_counter() - t0
        assert st2["kmers"] == kmers_per_step
        t = torch.tensor([wall2], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall2 = t.item()
        e2e = {"value": kmers_per_step * world * args.steps / wall2, "unit": UNIT, "h2d_bytes_per_step": n_text * world,
               "d2h_bytes_per_step": 80 * world, "ms_per_step": 1e3 * wall2 / args.steps}
        lib.jfgpu_host_free(hptr)

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import jfutil
        sample = args.cpu_sample_bases if os.path.exists(jfutil.REF_JF) else 5_000_000
        csize = 1 << max(10, (int(sample / 0.6) - 1).bit_length())
        with tempfile.TemporaryDirectory() as d:
            fa = os.path.join(d, "sample.fa")
            write_sample_fasta(fa, sample)
            v, kind, cores, secs = cpu_reference_run(fa, sample, k, csize, os.cpu_count() or 1, d)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "%d bp iid FASTA, jellyfish count -m 21 -s %d -t %d -C (Counting phase, %.1f s)" % (sample, csize, cores, secs)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": "k=%d canonical, %d bp synthetic FASTA per GPU (device-generated iid ACGT, 70-col lines, "
                                   "shape of generate_sequence), -s %s per GPU -> global 2^%d slots (%d-bit slots, %.1f GB per GPU)"
                                   % (k, n_bases, args.size, info["lsize"], info["slot_bits"], info["table_bytes"] / 1e9),
                       "l2": "inputs (%.1f GB text, %.1f GB table) exceed L2; no flush needed" % (n_text / 1e9, info["table_bytes"] / 1e9),
                       "timed_region": "table clear + all kernels of a step; inputs resident in HBM",
                       "parallelism": "1 GPU" if world == 1 else "table sharded by top hash bits over %d GPUs, NCCL all-to-all" % world},
            "device_seconds_per_step": dev_secs / args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "distinct": distinct_total, "load_factor": distinct_total / float(info["size"]),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
