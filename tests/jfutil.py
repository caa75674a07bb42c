"""Helpers shared by the tests: locate binaries, run them, split jellyfish databases."""
import json
import os
import subprocess
import hashlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_JF = os.path.join(REF_DIR, "jellyfish")            # the unmodified reference, built by oracle/Makefile
REF_GEN = os.path.join(REF_DIR, "generate_sequence")
ORACLE_C = os.path.join(REF_DIR, "jf_oracle")          # the independent C restatement
OUR_JF = os.path.join(ROOT, "jellyfish_b200", "lib", "jellyfish-b200")
LIB = os.path.join(ROOT, "jellyfish_b200", "lib", "libjfgpu.so")

SEMANTIC_KEYS = ("size", "key_len", "val_len", "max_reprobe", "reprobes", "counter_len", "format",
                 "canonical", "matrix1", "alignment")


def run(cmd, **kw):
    env = dict(os.environ, SOURCE_DATE_EPOCH="0")
    env.update(kw.pop("env", {}))
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, **kw)
    if r.returncode != 0:
        raise RuntimeError("command failed (%d): %s\nstdout: %s\nstderr: %s" % (
            r.returncode, " ".join(map(str, cmd)), r.stdout.decode(errors="replace")[-2000:],
            r.stderr.decode(errors="replace")[-2000:]))
    return r


def split_db(path):
    """-> (header dict, body bytes) of a jellyfish database."""
    with open(path, "rb") as f:
        data = f.read()
    hlen = int(data[:9])
    raw = data[9:9 + hlen].rstrip(b"\0")
    return json.loads(raw.decode()), data[9 + hlen:]


def semantic(header):
    return {k: header.get(k) for k in SEMANTIC_KEYS}


def md5(b):
    return hashlib.md5(b).hexdigest()


def records(header, body):
    """-> list of (key int, count int)"""
    kb = (header["key_len"] + 7) // 8
    cl = header["counter_len"]
    rec = kb + cl
    out = []
    for i in range(0, len(body) - rec + 1, rec):
        out.append((int.from_bytes(body[i:i + kb], "little"), int.from_bytes(body[i + kb:i + rec], "little")))
    return out


def generate(prefix, seed, *lengths, extra=()):
    run([REF_GEN, "-o", prefix, "-s", str(seed)] + list(extra) + [str(x) for x in lengths])


def subst(args, files):
    """'@name' in a case's switches stands for the path of that generated input (e.g. --if @multi2.fa)."""
    return [files[a[1:]] if a.startswith("@") else a for a in args]


def hash_pos(header, key):
    """Original position of a key: RectangularBinaryMatrix::times (bit i of the key selects columns[c-1-i],
    rectangular_binary_matrix.hpp:223-261) modulo the table size."""
    m = header["matrix1"]
    size = header["size"]
    if m.get("identity"):
        return key & (size - 1)
    cols, c = m["columns"], m["c"]
    h, i = 0, 0
    while key:
        if key & 1:
            h ^= cols[c - 1 - i]
        key >>= 1
        i += 1
    return h & (size - 1)
