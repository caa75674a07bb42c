"""One process per GPU: the hash table sharded by the top bits of the hash position.

The reference is a single process; this module is the multi-GPU form of
`hash_counter`: every rank parses its own part of the input, the canonical k-mers are bucketed
by the rank that owns their table position (device kernel, `jfgpu_extract_route`), exchanged
with one NCCL all-to-all per batch, and inserted by their owner (`jfgpu_insert_keys`).  All
ranks draw the same hash matrix (the reference's deterministic random stream), and shard r owns
the positions whose top log2(world) bits equal r, so the rank-ordered concatenation of the
shard dumps is byte-identical to the single-GPU dump.

Two forms of the exchange: 4-byte region RECORDS (RecordExchange, the default where the table geometry allows it: k <= 21,
32-bit slots) and packed KEYS (any geometry).  The exchange logic (bucket capacities, count exchange, uneven all-to-all,
ordering of the shard files) is plain torch.distributed code and is exercised on CPU with the gloo backend in
tests/test_distributed_cpu.py through the `RouteBackend` seam below.
"""
import os

import torch
import torch.distributed as dist


class RouteBackend(object):
    """What the exchange needs from the engine; the CUDA engine implements it with kernels."""

    key_words = 1

    def extract_route(self, text, begin, end, keys, capacity, counts):
        raise NotImplementedError

    def insert_keys(self, keys, n):
        raise NotImplementedError


class EngineBackend(RouteBackend):
    """libjfgpu.so (sm_100a kernels) behind the seam."""

    def __init__(self, hc):
        self.hc = hc
        self.key_words = hc.key_words

    def extract_route(self, text, begin, end, keys, capacity, counts):
        # the engine works on torch's current stream, so its kernels are ordered with the NCCL
        # collectives and the tensor ops around them
        ptr, n = text
        self.hc.extract_route(ptr, n, keys.data_ptr(), capacity, counts.data_ptr(), begin=begin, end=end,
                              stream=torch.cuda.current_stream().cuda_stream)

    def insert_keys(self, keys, n):
        if n:
            self.hc.insert_keys(keys.data_ptr(), n, stream=torch.cuda.current_stream().cuda_stream)


def exchange_and_insert(backend, world, send, counts, capacity, recv, before_payload=None):
    """Uneven all-to-all of the bucketed keys, then insertion on the owner.

    send:   int64 tensor [world, capacity * key_words]; bucket d holds counts[d] keys for rank d
    counts: int64 tensor [world] (device of `send`)
    recv:   int64 tensor [world, capacity * key_words] scratch
    before_payload: optional callable run after the (host-synchronising) count exchange and before
        the payload exchange -- the pipelined caller launches the next batch's extraction there, so
        that it overlaps the NVLink transfer and the insertion of this batch.
    Returns the number of keys this rank received."""
    kw = backend.key_words
    if world == 1:
        n = int(counts[0].item())
        if before_payload:
            before_payload()
        backend.insert_keys(send[0], n)
        return n
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts)
    sc = counts.tolist()
    rc = recv_counts.tolist()
    if max(sc) > capacity or max(rc) > capacity:
        raise RuntimeError("route bucket capacity exceeded (%d > %d)" % (max(max(sc), max(rc)), capacity))
    if before_payload:
        before_payload()
    total = sum(rc)
    flat_out = recv.view(-1)[:total * kw]
    if dist.get_backend() == "nccl":
        # list form: the buckets are sent straight from where the kernel wrote them (no packing copy)
        outs, o = [], 0
        for s in range(world):
            outs.append(flat_out[o:o + rc[s] * kw])
            o += rc[s] * kw
        dist.all_to_all(outs, [send[d, :sc[d] * kw] for d in range(world)])
    else:
        flat_in = torch.cat([send[d, :sc[d] * kw] for d in range(world)])
        dist.all_to_all_single(flat_out, flat_in, output_split_sizes=[c * kw for c in rc], input_split_sizes=[c * kw for c in sc])
    backend.insert_keys(flat_out, total)
    return total


CHUNK = 8192          # bytes of a record chunk (jf_kernels.cuh CHUNK_BYTES)


class RecordExchange(object):
    """The record form of the exchange (include/jfgpu.h, jfgpu_shard_*): every rank's K1 writes 4-byte records of the GLOBAL
    table's regions into a send pool whose chunk arenas belong to the owning shards; the chunks cross NVLink as they are
    (one NCCL all-to-all per round, list form: no packing copy), the owner re-files them under its own regions
    (restage_kernel) and drains them like a single GPU.  Two send banks: the extraction of round r+1 (stream A) runs beside
    the exchange and the restaging of round r (stream B)."""

    def __init__(self, hc, world, rank, dev, send_gb=None):
        self.hc, self.world, self.rank, self.dev = hc, world, rank, dev
        self.ok = False
        self.trace = None
        if not hc.shard_setup(0, 0, 0, 0, 0, 0):       # geometry probe: nothing is allocated for tables the record form does not cover
            return
        free, _ = torch.cuda.mem_get_info(dev)
        # two send banks + one receive pool = 3 x world x arena; a quarter of the free memory, at most 36 GB, for the three
        budget = min(free // 4, 36 << 30) if send_gb is None else int(send_gb * (1 << 30))
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        floor = 2 * n_sm * max(1, 1024 // world) + 64
        self.arena = max(floor, budget // (3 * world * (CHUNK + 8)))
        try:
            self.send = torch.empty(2 * world * self.arena * CHUNK, dtype=torch.uint8, device=dev)
            self.send_dir = torch.empty(2 * world * self.arena * 8, dtype=torch.uint8, device=dev)
            self.recv = torch.empty(world * self.arena * CHUNK, dtype=torch.uint8, device=dev)
            self.recv_dir = torch.empty(world * self.arena * 8, dtype=torch.uint8, device=dev)
        except RuntimeError:
            return
        if not hc.shard_setup(self.send.data_ptr(), self.send_dir.data_ptr(), self.arena, self.recv.data_ptr(), self.recv_dir.data_ptr(), self.arena):
            self.send = self.send_dir = self.recv = self.recv_dir = None
            return
        self.round_bytes = hc.shard_round_bytes()
        if self.round_bytes < (1 << 20):
            return
        # the chunk counts travel on a communicator of their own, so that they never queue behind the chunks of the round before
        self.pg_counts = dist.new_group(backend="nccl") if dist.get_backend() == "nccl" else None
        self.sa = torch.cuda.Stream(device=dev)       # extraction
        self.sb = torch.cuda.Stream(device=dev)       # exchange + restaging
        self.sent = [torch.cuda.Event(), torch.cuda.Event()]     # bank b has been sent (may be overwritten)
        self.ok = True

    def _views(self, pool, bank, counts, unit):
        w, a = self.world, self.arena
        return [pool[((bank * w + d) * a) * unit:((bank * w + d) * a + counts[d]) * unit] for d in range(w)]

    def add_device_text(self, ptr, n, begin, end):
        self._rounds(n, begin, end, lambda off, ln, bank: ptr + off)

    def add_host_text(self, hptr, n, begin, end):
        """Pinned host text: every round's slice is copied to a device staging buffer (one per bank) on the extraction stream."""
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        if getattr(self, "_stage", None) is None:
            self._stage = [torch.empty(self.round_bytes + 256, dtype=torch.uint8, device=self.dev) for _ in range(2)]

        def fetch(off, ln, bank):
            dst = self._stage[bank].data_ptr()
            if ln and lib.jfgpu_memcpy_h2d(C.c_void_p(dst), C.c_void_p(hptr + off), ln, C.c_void_p(self.sa.cuda_stream)):
                raise RuntimeError("host to device copy failed")
            return dst
        self._rounds(n, begin, end, fetch)

    def _rounds(self, n, begin, end, fetch):
        w = self.world
        rounds = (n + self.round_bytes - 1) // self.round_bytes if n else 0
        t = torch.tensor([rounds], dtype=torch.int64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rounds_all = max(int(t.item()), 1)        # every rank takes part in every exchange
        for ev in self.sent:
            ev.record(self.sb)
        marks = []                                  # CUDA events around the stages of every round (self.trace)
        for r in range(rounds_all):
            bank = r & 1
            off = r * self.round_bytes
            ln = max(0, min(self.round_bytes, n - off))
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            marks.append(ev)
            with torch.cuda.stream(self.sa):
                self.sa.wait_event(self.sent[bank])
                ev[0].record(self.sa)
                if ln or (r == 0 and begin) or (r == rounds_all - 1 and end):
                    src = fetch(off, ln, bank)
                    self.hc.shard_extract(src, ln, bank, begin and r == 0, end and off + ln >= n, stream=self.sa.cuda_stream)
                ev[1].record(self.sa)
                counts = self.hc.shard_pack(bank, stream=self.sa.cuda_stream)       # synchronises stream A
                sc = torch.tensor(counts, dtype=torch.int64, device=self.dev)
                rc = torch.empty_like(sc)
                dist.all_to_all_single(rc, sc, group=self.pg_counts)
                rcl = rc.tolist()                  # (waits for this tiny exchange only: stream B keeps working on the round before)
                if max(rcl) > self.arena:
                    raise RuntimeError("route bucket capacity exceeded (%d chunks > %d)" % (max(rcl), self.arena))
            with torch.cuda.stream(self.sb):
                ev[2].record(self.sb)
                # this rank's own chunks do not travel: the restaging reads them where K1 left them
                me = self.rank
                outs = [self.recv[(s * self.arena) * CHUNK:(s * self.arena + (0 if s == me else rcl[s])) * CHUNK] for s in range(w)]
                ins = self._views(self.send, bank, [0 if d == me else c for d, c in enumerate(counts)], CHUNK)
                dist.all_to_all(outs, ins)
                outs_d = [self.recv_dir[(s * self.arena) * 8:(s * self.arena + (0 if s == me else rcl[s])) * 8] for s in range(w)]
                dist.all_to_all(outs_d, self._views(self.send_dir, bank, [0 if d == me else c for d, c in enumerate(counts)], 8))
                ev[3].record(self.sb)
                self.hc.shard_unpack(rcl, self_bank=bank, stream=self.sb.cuda_stream)
                ev[4].record(self.sb)
                self.sent[bank].record(self.sb)      # (the bank is free again once its own chunks have been restaged)
        self.sa.synchronize()
        self.sb.synchronize()
        t = [0.0, 0.0, 0.0]
        for ev in marks:
            t[0] += ev[0].elapsed_time(ev[1]); t[1] += ev[2].elapsed_time(ev[3]); t[2] += ev[3].elapsed_time(ev[4])
        self.trace = {"rounds": rounds_all, "extract_ms": t[0], "exchange_ms": t[1], "restage_ms": t[2]}


class ShardedCounter(object):
    """hash_counter over `world` GPUs.  `size` is the GLOBAL table size (jellyfish count -s)."""

    def __init__(self, size, val_len=7, k=None, canonical=False, rank=0, world=1, device=0, reprobes=126,
                 batch_bytes=256 << 20, slack=1.25, exchange="auto", send_gb=None, **engine_kw):
        from .engine import HashCounter
        self.rank, self.world = rank, world
        self.hc = HashCounter(size, val_len, k=k, canonical=canonical, reprobes=reprobes, device=device,
                              shard_index=rank, n_shards=world, allow_regrow=(world == 1), max_batch_bytes=batch_bytes, **engine_kw)
        self.backend = EngineBackend(self.hc)
        self.batch_bytes = batch_bytes
        self.dev = torch.device("cuda", device)
        self.records = None
        if world > 1 and exchange in ("auto", "records"):
            rx = RecordExchange(self.hc, world, rank, self.dev, send_gb=send_gb)
            if rx.ok:
                self.records = rx
            elif exchange == "records":
                raise RuntimeError("the record exchange does not cover this table geometry")
            del rx
        if world > 1 and self.records is None:
            kw = self.hc.key_words
            # a batch of B bytes yields at most B k-mers, spread evenly over the owners by the hash
            self.capacity = int(batch_bytes / world * slack) + 65536
            self.send = torch.empty((world, self.capacity * kw), dtype=torch.int64, device=self.dev)
            self.recv = torch.empty((world, self.capacity * kw), dtype=torch.int64, device=self.dev)
            self.counts = torch.zeros(world, dtype=torch.int64, device=self.dev)
        self._host_stage = None
        self._send2 = self._counts2 = self._sa = None
        # a dedicated (non-default) stream: its handle is passed to the engine so that kernels, tensor
        # ops and NCCL collectives are ordered on one stream (handle 0 would mean "engine stream")
        self.stream = torch.cuda.Stream(device=self.dev)

    def add_device_text(self, ptr, n, begin=True, end=True):
        if self.world == 1:
            self.hc.add_device_text(ptr, n, begin=begin, end=end)
            return
        torch.cuda.current_stream(self.dev).synchronize()     # the caller's text is complete
        if self.records is not None:
            self.records.add_device_text(ptr, n, begin, end)
            return
        with torch.cuda.stream(self.stream):
            self._add_device_text(ptr, n, begin, end)
        self.stream.synchronize()

    def _add_device_text(self, ptr, n, begin, end):
        """Two-stage software pipeline: the extraction of batch i+1 (stream A) overlaps the NVLink
        exchange and the owner-side insertion of batch i (stream B).  Two send/count buffer sets."""
        rounds = (n + self.batch_bytes - 1) // self.batch_bytes
        t = torch.tensor([rounds], dtype=torch.int64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rounds_all = int(t.item())        # every rank takes part in every exchange
        if self._send2 is None:
            self._send2 = torch.empty_like(self.send)
            self._counts2 = torch.zeros_like(self.counts)
            self._sa = torch.cuda.Stream(device=self.dev)
        sends, cnts = (self.send, self._send2), (self.counts, self._counts2)
        sb = self.stream                    # stream B: exchange + insertion
        sa = self._sa                       # stream A: extraction
        sa.wait_stream(sb)
        done_extract = [torch.cuda.Event(), torch.cuda.Event()]
        done_use = [torch.cuda.Event(), torch.cuda.Event()]
        for ev in done_use:
            ev.record(sb)

        def launch_extract(i):
            off = i * self.batch_bytes
            ln = max(0, min(self.batch_bytes, n - off))
            with torch.cuda.stream(sa):
                sa.wait_event(done_use[i & 1])           # the buffers of batch i-2 have been sent
                cnts[i & 1].zero_()
                if ln:
                    self.backend.extract_route((ptr + off, ln), begin and off == 0, end and off + ln >= n,
                                               sends[i & 1], self.capacity, cnts[i & 1])
                done_extract[i & 1].record(sa)

        if rounds_all:
            launch_extract(0)
        for i in range(rounds_all):
            sb.wait_event(done_extract[i & 1])
            nxt = (lambda j=i + 1: launch_extract(j)) if i + 1 < rounds_all else None
            exchange_and_insert(self.backend, self.world, sends[i & 1], cnts[i & 1], self.capacity, self.recv, before_payload=nxt)
            done_use[i & 1].record(sb)
        sa.synchronize()

    def add_host_text(self, hptr, n, begin=True, end=True):
        """Host memory (pinned) -> staged through a device buffer batch by batch."""
        if self.world == 1:
            import ctypes as C
            self.hc.add_text((C.c_void_p(hptr), n), begin=begin, end=end)
            return
        if self.records is not None:
            self.records.add_host_text(hptr, n, begin, end)
            return
        with torch.cuda.stream(self.stream):
            self._add_host_text(hptr, n, begin, end)
        self.stream.synchronize()

    def _add_host_text(self, hptr, n, begin, end):
        if self._host_stage is None:
            self._host_stage = torch.empty(self.batch_bytes + 256, dtype=torch.uint8, device=self.dev)
        from . import _lib
        lib = _lib.load()
        import ctypes as C
        off = 0
        rounds = (n + self.batch_bytes - 1) // self.batch_bytes
        t = torch.tensor([rounds], dtype=torch.int64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for i in range(int(t.item())):
            ln = max(0, min(self.batch_bytes, n - off))
            self.counts.zero_()
            if ln:
                # host -> device on the working stream (asynchronous for pinned memory)
                if lib.jfgpu_memcpy_h2d(C.c_void_p(self._host_stage.data_ptr()), C.c_void_p(hptr + off), ln, C.c_void_p(self.stream.cuda_stream)):
                    raise RuntimeError("host to device copy failed")
                self.backend.extract_route((self._host_stage.data_ptr(), ln), begin and off == 0, end and off + ln >= n,
                                           self.send, self.capacity, self.counts)
            exchange_and_insert(self.backend, self.world, self.send, self.counts, self.capacity, self.recv)
            off += ln

    def done(self):
        return self.hc.done()

    def dump_shard(self, path, **kw):
        """Every rank writes `path.<rank>`: header (global size/matrix) + its sorted records."""
        return self.hc.dump("%s.%d" % (path, self.rank), **kw)


def concat_shards(path, world, out=None):
    """Rank-ordered concatenation of the shard files = the single-GPU database
    (positions of shard r all precede those of shard r+1)."""
    out = out or path
    with open(out, "wb") as fo:
        for r in range(world):
            with open("%s.%d" % (path, r), "rb") as fi:
                data = fi.read()
            hlen = int(data[:9])
            fo.write(data if r == 0 else data[9 + hlen:])
    return out
