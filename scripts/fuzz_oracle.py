#!/usr/bin/env python
"""Differential fuzz of the C restatement (oracle/_ref/jf_oracle) against the UNMODIFIED reference
(oracle/_ref/jellyfish): random switches and small random inputs, header keys and record bodies
compared byte for byte. Build-container tool (needs the reference binary); failures print the
command line so that the case can be added to tests/cases.py.
    python scripts/fuzz_oracle.py [N] [SEED]
"""
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import jfutil  # noqa: E402

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)


def rand_seq(n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def rand_fasta(path):
    out = []
    for _ in range(rng.randrange(1, 6)):
        n = rng.choice([0, 1, 5, 40, 200, 3000, 20000])
        alpha = rng.choice(["ACGT", "ACGT", "ACGTacgt", "ACGTN", "AC", "A", "ACGTRYn-"])
        s = rand_seq(n, alpha)
        if rng.random() < 0.3 and n > 100:      # repeats -> big counts
            s = s[:50] * (n // 50)
        w = rng.choice([1, 7, 60, 70, 100000])
        eol = rng.choice(["\n", "\n", "\r\n"])
        out.append(">h%d x" % rng.randrange(1000) + eol + "".join(s[i:i + w] + eol * rng.choice([1, 1, 1, 2]) for i in range(0, len(s), w)))
    data = "".join(out)
    if rng.random() < 0.3:
        data = data.rstrip("\r\n")
    open(path, "w", newline="").write(data)


def rand_fastq(path):
    out = []
    wrap = rng.choice([0, 0, 0, 50])        # multi-line records now and then
    for i in range(rng.randrange(1, 40)):
        n = rng.choice([1, 30, 76, 150, 400])
        s = rand_seq(n, rng.choice(["ACGT", "ACGTN", "ACGTacgt"]))
        q = "".join(chr(rng.randrange(33, 75)) if rng.random() < 0.1 else rng.choice("FGHIJ") for _ in range(n))
        if wrap:
            s = "\n".join(s[j:j + wrap] for j in range(0, n, wrap))
            q = "\n".join(q[j:j + wrap] for j in range(0, n, wrap))
        out.append("@r%d\n%s\n+%s\n%s\n" % (i, s, rng.choice(["", "r%d" % i]), q))
    open(path, "w").write("".join(out))


bad = 0
with tempfile.TemporaryDirectory() as d:
    for it in range(n_iter):
        files = []
        kinds = sorted(rng.random() < 0.3 for _ in range(rng.randrange(1, 4)))   # FASTA files first (see -Q note in tests/cases.py)
        for j, fq in enumerate(kinds):
            p = os.path.join(d, "in%d_%d" % (it, j))
            (rand_fastq if fq else rand_fasta)(p)
            files.append(p)
        k = rng.choice([1, 2, 3, 4, 5, 8, 11, 12, 15, 16, 17, 21, 25, 31, 32, 33, 40, 48, 63, 64, rng.randrange(1, 65)])
        size = rng.choice(["1", "2", "10", "100", "1k", "5k", "64k", "100k", "1M", str(rng.randrange(1, 300000))])
        args = ["-m", str(k), "-s", size]
        if rng.random() < 0.6:
            args.append("-C")
        if rng.random() < 0.3:
            args += ["-c", str(rng.choice([1, 2, 3, 5, 7, 10, 16]))]
        if rng.random() < 0.3:
            args += ["-p", str(rng.choice([1, 2, 5, 10, 30, 62, 126, 200]))]
        if rng.random() < 0.2:
            args += ["--out-counter-len", str(rng.choice([1, 2, 3, 7]))]
        if rng.random() < 0.2:
            args += ["-L", str(rng.choice([1, 2, 5]))]
        if rng.random() < 0.15:
            args += ["-U", str(rng.choice([1, 3, 100]))]
        if rng.random() < 0.1:
            args.append("--text")
        if rng.random() < 0.15:
            args += ["--if", rng.choice(files)]
        if rng.random() < 0.15:
            args += rng.choice([["-Q", rng.choice("#5AF")], ["--min-quality", str(rng.randrange(0, 9)), "--quality-start", "33"]])
        u = rng.random()
        if u < 0.1:
            args += ["--bf-size", rng.choice(["100", "5k", "100k"]), "--bf-fp", rng.choice(["0.01", "0.2", "0.001"])]
        elif u < 0.2:
            bc = os.path.join(d, "f.bc")
            bargs = ["-m", str(k), "-s", rng.choice(["100", "5k", "100k"]), "-f", rng.choice(["0.001", "0.05", "0.3"])] + (["-C"] if "-C" in args else [])
            r1 = subprocess.run([jfutil.REF_JF, "bc", "-t", "2"] + bargs + ["-o", bc] + files[:2], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            r2 = subprocess.run([jfutil.ORACLE_C, "bc"] + bargs + ["-o", bc + ".o"] + files[:2], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            if r1.returncode == 0 and r2.returncode == 0:
                if jfutil.split_db(bc)[1] != jfutil.split_db(bc + ".o")[1]:
                    bad += 1
                    print("MISMATCH #%d: bc files differ: bc %s %s" % (it, " ".join(bargs), " ".join(files[:2])))
                args += ["--bc", bc]
        r_db, o_db = os.path.join(d, "r.jf"), os.path.join(d, "o.jf")
        for f in (r_db, o_db):
            if os.path.exists(f):
                os.remove(f)
        rr = subprocess.run([jfutil.REF_JF, "count", "-t", "1"] + args + ["-o", r_db] + files, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        ro = subprocess.run([jfutil.ORACLE_C, "count"] + args + ["-o", o_db] + files, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        ok = True
        why = ""
        if (rr.returncode == 0) != (ro.returncode == 0):
            ok, why = False, "exit status ref %d oracle %d: %s | %s" % (rr.returncode, ro.returncode, rr.stderr[-200:], ro.stderr[-200:])
        elif rr.returncode == 0:
            h1, b1 = jfutil.split_db(r_db)
            h2, b2 = jfutil.split_db(o_db)
            if jfutil.semantic(h1) != jfutil.semantic(h2):
                diff = [kk for kk in jfutil.SEMANTIC_KEYS if h1.get(kk) != h2.get(kk)]
                ok, why = False, "header keys differ: %s (ref %s oracle %s)" % (diff, [str(h1.get(kk))[:60] for kk in diff], [str(h2.get(kk))[:60] for kk in diff])
            elif b1 != b2:
                ok, why = False, "bodies differ (%d vs %d bytes)" % (len(b1), len(b2))
        if not ok:
            bad += 1
            keep = os.path.join("/tmp", "fuzz_fail_%d" % it)
            os.makedirs(keep, exist_ok=True)
            kept = []
            for f in files:
                subprocess.run(["cp", f, keep])
                kept.append(os.path.join(keep, os.path.basename(f)))
            print("MISMATCH #%d: %s\n   count %s %s" % (it, why, " ".join(args), " ".join(kept)))
print("%d iterations, %d mismatches" % (n_iter, bad))
sys.exit(1 if bad else 0)
