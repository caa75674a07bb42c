"""Host-side mirror of the reference's counting interfaces, over the C ABI.

`HashCounter` follows the reference's `hash_counter` / SWIG `HashCounter`
(include/jellyfish/hash_counter.hpp:50-172, swig/hash_counter.i): construct with a size and a
value length, add k-mers, `done()`, read counts back, dump.  The difference is the unit of
work: k-mers are added a *buffer of FASTA text* at a time (`add_text`, `add_files`) because
parsing, canonicalisation, hashing and insertion run fused on the device.
"""
import ctypes as C
import json
import os
import time

from . import _lib as L

UINT64_MAX = (1 << 64) - 1


class JellyfishError(RuntimeError):
    """Raised for any non-zero status of the engine (reference: std::runtime_error / err::die)."""

    def __init__(self, code, msg):
        RuntimeError.__init__(self, msg)
        self.code = code


def reference_matrix(r, c, skip=0):
    """Columns of the hash matrix the reference draws (host arithmetic, no GPU needed)."""
    lib = L.load()
    cols = (C.c_uint64 * c)()
    rc = lib.jfgpu_reference_matrix(r, c, skip, cols)
    if rc:
        raise JellyfishError(rc, "invalid matrix dimensions")
    return list(cols)


def mer_to_int(s):
    """'ACGT...' -> 2-bit packed integer, first base most significant (mer_dna.hpp:525-542)."""
    v = 0
    for ch in s:
        v = (v << 2) | "ACGT".index(ch.upper())
    return v


def int_to_mer(v, k):
    return "".join("ACGT"[(v >> (2 * (k - 1 - i))) & 3] for i in range(k))


def canonical_int(v, k):
    rc = 0
    x = v
    for _ in range(k):
        rc = (rc << 2) | (3 - (x & 3))
        x >>= 2
    return min(v, rc)


class HashCounter(object):
    def __init__(self, size, val_len=7, k=None, canonical=False, reprobes=126, device=0,
                 shard_index=0, n_shards=1, allow_regrow=True, max_batch_bytes=0, matrix_skip=0,
                 pool_bytes=0, no_partition=False, part_min_mb=0, k2_mode=0, region_mb=0, bf_size=0, bf_fp=0.0,
                 bloom_counter=False, min_qual=0):
        if k is None:
            raise ValueError("k (mer length) is required")
        self._lib = L.load()
        self._h = C.c_void_p()
        p = L.Params()
        p.struct_size = C.sizeof(L.Params)
        p.k, p.size, p.counter_len, p.max_reprobe = k, size, val_len, reprobes
        p.canonical, p.allow_regrow, p.device = int(bool(canonical)), int(bool(allow_regrow)), device
        p.shard_index, p.n_shards, p.max_batch_bytes, p.matrix_skip = shard_index, n_shards, max_batch_bytes, matrix_skip
        p.pool_bytes, p.no_partition, p.part_min_mb = pool_bytes, int(bool(no_partition)), part_min_mb
        p.k2_mode, p.region_mb = k2_mode, region_mb
        p.bf_size, p.bf_fp, p.bloom_counter = bf_size, bf_fp, int(bool(bloom_counter))
        p.min_qual = ord(min_qual) if isinstance(min_qual, str) else int(min_qual)
        rc = self._lib.jfgpu_create(C.byref(p), C.byref(self._h))
        if rc:
            self._h = C.c_void_p()
            raise JellyfishError(rc, self._lib.jfgpu_last_error(None).decode())
        self.k = k
        self.canonical = bool(canonical)
        self.key_words = 2 if k > 32 else 1
        self.n_shards = n_shards

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc):
        if rc:
            raise JellyfishError(rc, self._lib.jfgpu_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.jfgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- hash_counter interface -----------------------------------------------------------------
    def info(self):
        ti = L.TableInfo()
        self._check(self._lib.jfgpu_table_info_get(self._h, C.byref(ti)))
        d = {f: getattr(ti, f) for f, _ in L.TableInfo._fields_ if f not in ("matrix_columns", "reprobes")}
        d["matrix_columns"] = None if not ti.matrix_columns else [ti.matrix_columns[i] for i in range(ti.matrix_c)]
        d["reprobes"] = [ti.reprobes[i] for i in range(ti.max_reprobe + 1)]
        return d

    def size(self):
        return self.info()["size"]

    def val_len(self):
        return self.info()["val_len"]

    def add_text(self, data, begin=True, end=True):
        """Count every k-mer of a buffer of FASTA text held in host memory (bytes or a pointer/size pair)."""
        flags = (L.FILE_BEGIN if begin else 0) | (L.FILE_END if end else 0)
        if isinstance(data, tuple):
            ptr, n = data
        else:
            buf = bytes(data)
            ptr, n = C.cast(C.c_char_p(buf), C.c_void_p), len(buf)
        self._check(self._lib.jfgpu_feed(self._h, ptr, n, flags))

    def add_device_text(self, dev_ptr, n, begin=True, end=True, stream=None):
        """Same with the text already in device memory (e.g. a torch uint8 tensor's data_ptr())."""
        flags = (L.FILE_BEGIN if begin else 0) | (L.FILE_END if end else 0)
        self._check(self._lib.jfgpu_feed_device(self._h, C.c_void_p(dev_ptr), n, flags, C.c_void_p(stream or 0)))

    def add_files(self, paths, chunk=64 << 20):
        """mer_counter_base::start over a list of files (count_main.cc:152-184)."""
        for path in paths:
            with open(path, "rb") as f:
                first = True
                cur = f.read(chunk)
                if not cur:
                    continue
                while True:
                    nxt = f.read(chunk)
                    self.add_text(cur, begin=first, end=not nxt)
                    first = False
                    if not nxt:
                        break
                    cur = nxt

    def extract_route(self, dev_ptr, n, keys_ptr, capacity, counts_ptr, begin=True, end=True, stream=None):
        flags = (L.FILE_BEGIN if begin else 0) | (L.FILE_END if end else 0)
        self._check(self._lib.jfgpu_extract_route(self._h, C.c_void_p(dev_ptr), n, flags, C.c_void_p(keys_ptr),
                                                  capacity, C.c_void_p(counts_ptr), C.c_void_p(stream or 0)))

    def insert_keys(self, keys_ptr, n, stream=None):
        self._check(self._lib.jfgpu_insert_keys(self._h, C.c_void_p(keys_ptr), n, C.c_void_p(stream or 0)))

    # -- sharded counting, record exchange (include/jfgpu.h: jfgpu_shard_*) ----------------------------
    def shard_setup(self, send_pool, send_dir, send_arena_chunks, recv_pool, recv_dir, recv_seg_chunks):
        """Register the exchange buffers (device pointers).  False when the table geometry is not covered by the
        record exchange (the caller then uses extract_route / insert_keys)."""
        b = L.ShardBuffers(send_pool, send_dir, send_arena_chunks, recv_pool, recv_dir, recv_seg_chunks)
        rc = self._lib.jfgpu_shard_setup(self._h, C.byref(b))
        if rc == L.ERR_ARG:
            return False
        self._check(rc)
        return True

    def shard_round_bytes(self):
        return self._lib.jfgpu_shard_round_bytes(self._h)

    def shard_extract(self, dev_ptr, n, bank, begin=True, end=True, stream=None):
        flags = (L.FILE_BEGIN if begin else 0) | (L.FILE_END if end else 0)
        self._check(self._lib.jfgpu_shard_extract(self._h, C.c_void_p(dev_ptr), n, flags, bank, C.c_void_p(stream or 0)))

    def shard_pack(self, bank, stream=None):
        """Close the round: chunks per destination shard (synchronises the stream)."""
        n = self.n_shards
        counts = (C.c_uint64 * n)()
        self._check(self._lib.jfgpu_shard_pack(self._h, bank, counts, C.c_void_p(stream or 0)))
        return list(counts)

    def shard_unpack(self, counts, self_bank=None, stream=None):
        """counts: chunks received from every shard.  self_bank 0/1: this shard's own chunks are read from that send bank
        instead of the receive pool (they were not exchanged)."""
        arr = (C.c_uint64 * len(counts))(*counts)
        self._check(self._lib.jfgpu_shard_unpack(self._h, arr, 0xFFFFFFFF if self_bank is None else self_bank, C.c_void_p(stream or 0)))

    OP_COUNT, OP_PRIME, OP_UPDATE = 0, 1, 2

    def set_op(self, op):
        """COUNT (add), PRIME (insert with count 0) or UPDATE (add only to present keys): the two passes
        of `jellyfish count --if` (sub_commands/count_main.cc:288-295)."""
        self._check(self._lib.jfgpu_set_op(self._h, op))

    def clear(self):
        """Zero the table and statistics (same geometry and hash matrix)."""
        self._check(self._lib.jfgpu_clear(self._h))

    def done(self):
        """hash_counter::done -- drain the device, returns the statistics."""
        st = L.Stats()
        self._check(self._lib.jfgpu_finish(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in L.Stats._fields_}

    def stats(self):
        st = L.Stats()
        self._check(self._lib.jfgpu_get_stats(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in L.Stats._fields_}

    def get_many(self, mers):
        """Counts of a list of k-mers given as strings or packed ints (0 when absent)."""
        n = len(mers)
        kw = self.key_words
        keys = (C.c_uint64 * (n * kw))()
        for i, m in enumerate(mers):
            v = mer_to_int(m) if isinstance(m, str) else int(m)
            if self.canonical:
                v = canonical_int(v, self.k)
            keys[i * kw] = v & UINT64_MAX
            if kw == 2:
                keys[i * kw + 1] = v >> 64
        vals = (C.c_uint64 * n)()
        self._check(self._lib.jfgpu_lookup(self._h, keys, n, vals))
        return list(vals)

    def get(self, mer):
        v = self.get_many([mer])[0]
        return v if v else None

    __getitem__ = get

    def histogram(self, n_bins=10002):
        hist = (C.c_uint64 * n_bins)()
        self._check(self._lib.jfgpu_histogram(self._h, hist, n_bins))
        return list(hist)

    # -- Bloom structures (count --bf-size / --bc, `jellyfish bc`) ----------------------------------
    def bloom_info(self):
        bi = L.BloomInfo()
        self._check(self._lib.jfgpu_bloom_info_get(self._h, C.byref(bi)))
        d = {"mode": bi.mode, "nb_hashes": bi.nb_hashes, "m": bi.m, "nb_bytes": bi.nb_bytes}
        if bi.mode:
            d["matrix1"] = [bi.matrix1[i] for i in range(bi.matrix_c)]
            d["matrix2"] = [bi.matrix2[i] for i in range(bi.matrix_c)]
        return d

    def load_bloom_counter(self, path):
        """count --bc FILE (count_main.cc:191-206): filter by a Bloom counter written by `jellyfish bc`."""
        with open(path, "rb") as f:
            data = f.read()
        hlen = int(data[:9])
        hdr = json.loads(data[9:9 + hlen].rstrip(b"\0").decode())
        if hdr.get("format") != "bloomcounter":
            raise JellyfishError(L.ERR_FORMAT, "Invalid format '%s'. Expected 'bloomcounter'" % hdr.get("format"))
        if hdr["key_len"] != 2 * self.k:
            raise JellyfishError(L.ERR_ARG, "Invalid mer length in bloom filter")
        body = data[9 + hlen:]
        c = 2 * self.k
        m1 = (C.c_uint64 * c)(*hdr["matrix1"]["columns"])
        m2 = (C.c_uint64 * c)(*hdr["matrix2"]["columns"])
        self._check(self._lib.jfgpu_bloom_load(self._h, hdr["size"], hdr["nb_hashes"], m1, m2, body, len(body)))

    # -- dumper ------------------------------------------------------------------------------
    def dump_records(self, lower=0, upper=UINT64_MAX, out_counter_len=4, sink=None):
        """Sorted (position, key) record stream of this shard; returns bytes when no sink is given."""
        chunks = []

        def _sink(ctx, ptr, n):
            if sink == "discard":        # timing the device side of a dump: the bytes stay in the engine's pinned buffer
                return 0
            data = C.string_at(ptr, n)
            if sink is not None:
                sink(data)
            else:
                chunks.append(data)
            return 0

        cb = L.SINK_FN(_sink)
        nrec = C.c_uint64(0)
        self._check(self._lib.jfgpu_dump(self._h, lower, upper, out_counter_len, cb, None, C.byref(nrec)))
        return b"".join(chunks) if sink is None else nrec.value

    def header(self, out_counter_len=4, cmdline=()):
        """The file_header dictionary the reference writes (file_header.hpp:26-108)."""
        ti = self.info()
        m = {"r": ti["matrix_r"], "c": ti["matrix_c"], "identity": bool(ti["matrix_identity"])}
        if not ti["matrix_identity"]:
            m["columns"] = ti["matrix_columns"]
        sde = os.environ.get("SOURCE_DATE_EPOCH")
        return {
            "alignment": 8, "canonical": self.canonical, "cmdline": list(cmdline), "counter_len": out_counter_len,
            "exe_path": os.path.realpath(L.LIB_PATH), "format": "binary/sorted",
            "hostname": "hostname" if sde else os.uname().nodename, "key_len": ti["key_len"], "matrix1": m,
            "max_reprobe": ti["max_reprobe"], "pwd": "." if sde else os.getcwd(), "reprobes": ti["reprobes"],
            "size": ti["size"], "time": time.asctime(time.gmtime(int(sde))) if sde else time.asctime(),
            "val_len": ti["val_len"],
        }

    def dump(self, path, lower=0, upper=UINT64_MAX, out_counter_len=4, cmdline=()):
        """binary_dumper::dump -- header + sorted records (binary_dumper.hpp:62-69)."""
        with open(path, "wb") as f:
            write_header(f, self.header(out_counter_len, cmdline))
            return self.dump_records(lower, upper, out_counter_len, sink=f.write)


class BloomCounter(object):
    """`jellyfish bc` (sub_commands/bc_main.cc): a Bloom counter of the k-mers of some text, built on the device."""

    def __init__(self, size, fpr=0.001, k=None, canonical=False, device=0, max_batch_bytes=0):
        self.hc = HashCounter(1, 7, k=k, canonical=canonical, device=device, bf_size=size, bf_fp=fpr, bloom_counter=True,
                              max_batch_bytes=max_batch_bytes)
        self.k, self.canonical = k, bool(canonical)

    def close(self):
        self.hc.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def add_files(self, paths):
        self.hc.add_files(paths)

    def add_text(self, data, begin=True, end=True):
        self.hc.add_text(data, begin=begin, end=end)

    def info(self):
        return self.hc.bloom_info()

    def header(self, cmdline=()):
        bi = self.info()
        sde = os.environ.get("SOURCE_DATE_EPOCH")

        def mat(cols):
            return {"r": 64, "c": 2 * self.k, "identity": False, "columns": cols}
        return {
            "alignment": 8, "canonical": self.canonical, "cmdline": list(cmdline), "exe_path": os.path.realpath(L.LIB_PATH),
            "format": "bloomcounter", "hostname": "hostname" if sde else os.uname().nodename, "key_len": 2 * self.k,
            "matrix1": mat(bi["matrix1"]), "matrix2": mat(bi["matrix2"]), "nb_hashes": bi["nb_hashes"],
            "pwd": "." if sde else os.getcwd(), "size": bi["m"],
            "time": time.asctime(time.gmtime(int(sde))) if sde else time.asctime(),
        }

    def dump(self, path, cmdline=()):
        """header + filter.write_bits (bc_main.cc:113,139)."""
        with open(path, "wb") as f:
            write_header(f, self.header(cmdline))

            def _sink(ctx, ptr, n):
                f.write(C.string_at(ptr, n))
                return 0
            cb = L.SINK_FN(_sink)
            self.hc._check(self.hc._lib.jfgpu_bloom_dump(self.hc._h, cb, None))


def write_header(f, header):
    """generic_file_header::write (generic_file_header.hpp:88-111)."""
    h = json.dumps(header, sort_keys=True, separators=(",", ":"), ensure_ascii=False).encode()
    hlen = len(h)
    pad = (9 + hlen) % 8
    if pad:
        hlen += 8 - pad
    f.write(b"%09d" % hlen)
    f.write(h)
    if pad:
        f.write(b"\0" * (8 - pad))


class ReadMerFile(object):
    """Iterate the (mer, count) records of a binary/sorted database (swig/mer_file.i ReadMerFile)."""

    def __init__(self, path):
        with open(path, "rb") as f:
            data = f.read()
        hlen = int(data[:9])
        self.header = json.loads(data[9:9 + hlen].rstrip(b"\0").decode())
        self.body = data[9 + hlen:]
        self.k = self.header["key_len"] // 2
        self.key_bytes = (self.header["key_len"] + 7) // 8
        self.counter_len = self.header["counter_len"]

    def __iter__(self):
        rec = self.key_bytes + self.counter_len
        b = self.body
        for i in range(0, len(b) - rec + 1, rec):
            key = int.from_bytes(b[i:i + self.key_bytes], "little")
            yield int_to_mer(key, self.k), int.from_bytes(b[i + self.key_bytes:i + rec], "little")
