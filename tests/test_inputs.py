"""Host side of the inputs of `count` / `bc` (jellyfish_b200/csrc/host/jf_inputs.hpp): sequence files, then the standard output of
the generator commands of -g / -G / -S (reference: lib/generator_manager.cc, tests/multi_file.sh:16-33).  `jellyfish-b200 inputs`
writes the byte stream the engine would be fed and, with --marks, the chunks and their FILE_BEGIN / FILE_END flags: no device needed."""
import hashlib
import os
import shutil
import subprocess

import pytest

import jfutil


def _inputs(args, cwd, timeout=60):
    return subprocess.run([jfutil.OUR_JF, "inputs"] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)


@pytest.fixture()
def seqdir(tmp_path, built):
    (tmp_path / "a.fa").write_bytes(b">a\nACGTACGTAC\nGGT\n")
    (tmp_path / "b.fa").write_bytes(b">b\nTTTT\n>b2\nGGGGCC\n")
    (tmp_path / "r.fq").write_bytes(b"@r\nACGT\n+\nIIII\n")
    (tmp_path / "empty.fa").write_bytes(b"")
    return tmp_path


def test_files_then_command_outputs_in_order(seqdir):
    (seqdir / "cmds").write_text("\n   \n# comments and blank lines are skipped (generator_manager.cc:222-232)\ncat b.fa\n   \t cat r.fq\n  # another\nprintf '>c\\nCC\\n'\n")
    r = _inputs(["--marks", "-g", "cmds", "-G", "2", "-S", "/bin/sh", "a.fa", "empty.fa"], seqdir)
    assert r.returncode == 0, r.stderr
    want = [(seqdir / n).read_bytes() for n in ("a.fa", "empty.fa", "b.fa", "r.fq")] + [b">c\nCC\n"]
    assert r.stdout == b"".join(want)
    # every file and every command output is one input of its own: begin and end flags, nothing spans two of them
    assert r.stderr.decode().splitlines() == ["chunk %d begin end" % len(w) for w in want]


def test_chunks_of_one_input_carry_begin_first_and_end_last(seqdir):
    (seqdir / "cmds").write_text("cat a.fa b.fa\n")
    r = _inputs(["--marks", "--chunk", "7", "-g", "cmds", "r.fq"], seqdir)
    assert r.returncode == 0, r.stderr
    a = (seqdir / "r.fq").read_bytes()
    b = (seqdir / "a.fa").read_bytes() + (seqdir / "b.fa").read_bytes()
    assert r.stdout == a + b
    marks = r.stderr.decode().splitlines()

    def expect(n):
        sizes = [7] * (n // 7) + ([n % 7] if n % 7 else [])
        return ["chunk %d%s%s" % (s, " begin" if i == 0 else "", " end" if i == len(sizes) - 1 else "") for i, s in enumerate(sizes)]
    assert marks == expect(len(a)) + expect(len(b))


def test_commands_run_side_by_side(seqdir):
    """-G 2: two commands that each wait for the other one to have started can only finish when both run at the same time."""
    wait = "touch m%d; while [ ! -e m%d ]; do sleep 0.05; done; printf '>s%d\\nACGT\\n'"
    (seqdir / "cmds").write_text(wait % (1, 2, 1) + "\n" + wait % (2, 1, 2) + "\nprintf '>s3\\nAC\\n'\n")
    r = _inputs(["-g", "cmds", "-G", "2"], seqdir, timeout=30)
    assert r.returncode == 0, r.stderr
    assert r.stdout == b">s1\nACGT\n>s2\nACGT\n>s3\nAC\n"          # outputs are taken in the order of the command file


def test_large_outputs_pass_the_pump_unchanged(seqdir):
    """outputs larger than the pump's blocks and the pipe, ragged against the chunk size; a slow consumer must not lose bytes"""
    n1, n2 = 9_000_001, 5_000_003
    (seqdir / "cmds").write_text("head -c %d /dev/zero | tr '\\0' A\nhead -c %d /dev/zero | tr '\\0' C\n" % (n1, n2))
    r = _inputs(["--marks", "--chunk", "1M", "-g", "cmds", "-G", "2"], seqdir)
    assert r.returncode == 0, r.stderr
    assert len(r.stdout) == n1 + n2 and r.stdout[:n1] == b"A" * n1 and r.stdout[n1:] == b"C" * n2
    marks = r.stderr.decode().splitlines()
    assert sum(1 for m in marks if m.endswith(" begin")) == 2 and sum(1 for m in marks if m.endswith(" end")) == 2


@pytest.mark.skipif(not (shutil.which("gzip") and shutil.which("gunzip")), reason="no gzip here")
def test_gunzip_generators(seqdir):
    """the reference's own use of generators: `gunzip -c` lines (tests/multi_file.sh:16-23)"""
    payload = b">z\n" + b"ACGTTGCA" * 50000 + b"\n"
    (seqdir / "z.fa").write_bytes(payload)
    subprocess.run(["gzip", "-k", "z.fa"], cwd=seqdir, check=True)
    (seqdir / "cmds").write_text("gunzip -c z.fa.gz\ngunzip -c z.fa.gz\n")
    r = _inputs(["-g", "cmds", "-G", "2", "a.fa"], seqdir)
    assert r.returncode == 0, r.stderr
    assert hashlib.md5(r.stdout).hexdigest() == hashlib.md5((seqdir / "a.fa").read_bytes() + payload * 2).hexdigest()


def test_failing_command_fails_the_run(seqdir):
    """tests/multi_file.sh:25-33: a generator command that fails makes the run fail, with the reference's messages"""
    (seqdir / "cmds").write_text("cat a.fa\nfalse\ncat b.fa\n")
    r = _inputs(["-g", "cmds", "-G", "2"], seqdir)
    assert r.returncode != 0
    assert b"Command 'false' exited with error status 1" in r.stderr and b"Some generator commands failed" in r.stderr
    (seqdir / "cmds").write_text("kill -9 $$\n")
    r = _inputs(["-g", "cmds"], seqdir)
    assert r.returncode != 0 and b"killed by signal 9" in r.stderr


def test_command_file_and_shell_errors(seqdir):
    r = _inputs(["-g", "no_such_cmds"], seqdir)
    assert r.returncode != 0 and b"Failed to open cmds file 'no_such_cmds'" in r.stderr
    (seqdir / "cmds").write_text("cat a.fa\n")
    r = _inputs(["-g", "cmds", "-S", "/no/such/shell"], seqdir)
    assert r.returncode != 0 and b"not run" in r.stderr
    r = _inputs(["a.fa", "missing.fa"], seqdir)
    assert r.returncode != 0 and b"Can't open file 'missing.fa'" in r.stderr


@pytest.mark.skipif(not os.path.exists("/bin/bash"), reason="no bash here")
def test_shell_switch_selects_the_shell(seqdir):
    (seqdir / "cmds").write_text("[[ -n $BASH_VERSION ]] && printf '>bash\\nAC\\n'\n")
    r = _inputs(["-g", "cmds", "-S", "/bin/bash"], seqdir)
    assert r.returncode == 0 and r.stdout == b">bash\nAC\n", r.stderr


def test_failed_feed_stops_an_endless_generator(seqdir):
    """when the consumer fails (here: the standard output cannot be written; in `count`: the engine reports an error) the reader
    stops and the commands are ended -- an endless generator must not keep the run alive"""
    (seqdir / "cmds").write_text("yes ACGT\n")
    with open("/dev/full", "wb") as sink:
        r = subprocess.run([jfutil.OUR_JF, "inputs", "-g", "cmds"], cwd=seqdir, stdout=sink, stderr=subprocess.PIPE, timeout=60)
    assert r.returncode != 0 and b"Error writing the standard output" in r.stderr


def _running(cmdline):
    n = 0
    for pid in os.listdir("/proc"):
        if pid.isdigit():
            try:
                with open("/proc/%s/cmdline" % pid, "rb") as f:
                    n += f.read().replace(b"\0", b" ").strip() == cmdline
            except OSError:
                pass
    return n


def test_sigterm_to_the_driver_ends_the_commands(seqdir):
    """generator_manager.cc:120-160: a killed run takes its generator commands with it"""
    import signal
    import time
    secs = "9%05d.25" % (os.getpid() % 100000)
    (seqdir / "cmds").write_text("cat a.fa\nexec sleep %s\n" % secs)
    p = subprocess.Popen([jfutil.OUR_JF, "inputs", "-g", "cmds", "-G", "2"], cwd=seqdir, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    try:
        deadline = time.time() + 20
        while not _running(b"sleep " + secs.encode()) and time.time() < deadline:
            time.sleep(0.05)
        assert _running(b"sleep " + secs.encode()) == 1
        p.send_signal(signal.SIGTERM)
        p.wait(timeout=20)
        assert p.returncode == -signal.SIGTERM
        deadline = time.time() + 10
        while _running(b"sleep " + secs.encode()) and time.time() < deadline:
            time.sleep(0.05)
        assert _running(b"sleep " + secs.encode()) == 0
    finally:
        if p.poll() is None:
            p.kill()
