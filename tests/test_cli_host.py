"""CPU tests of the `basal` CLI's host logic that runs before any GPU work: option parsing and the
reference's error contract (main.cpp:272-364, 616-633)."""
import os
import subprocess

import pytest

import helpers

pytestmark = pytest.mark.skipif(not os.path.exists(helpers.GPU_BIN), reason="basal CLI not built")


def run(*args):
    return subprocess.run([helpers.GPU_BIN] + list(args), capture_output=True, timeout=60)


def test_unknown_option_exits_with_its_argv_index():
    r = run("-a", "x.fq", "-Q")
    assert r.returncode == 3 and b"unknown option: -Q" in r.stderr
    r = run("stray")
    assert r.returncode == 1 and b"unknown option: stray" in r.stderr
    r = run("-a=x.fq", "-d=y.fa", "-ax")                     # malformed short form
    assert r.returncode == 3


def test_missing_and_invalid_rule():
    r = run("-a", "x.fq", "-d", "y.fa")
    assert r.returncode == 1 and b"-M option is required" in r.stderr
    r = run("-a", "x.fq", "-d", "y.fa", "-M", "CT")
    assert r.returncode == 1 and b"should be assigned first before :" in r.stderr
    r = run("-a", "x.fq", "-d", "y.fa", "-M", "N:T")
    assert r.returncode == 1 and b"not in A/C/G/T" in r.stderr
    r = run("-a", "x.fq", "-d", "y.fa", "-M", "C:TX")
    assert r.returncode == 1 and b"not in A/C/G/T/-" in r.stderr


def test_range_checks():
    assert b"seed size must be between 10 and 16" in run("-s", "9", "-M", "C:T").stderr
    assert b"index interval exceeds max value:16" in run("-I", "17", "-M", "C:T").stderr
    assert b"number of multi-hits exceeds max value:1000" in run("-w", "1001", "-M", "C:T").stderr
    assert b"invalid -r value" in run("-r", "3", "-M", "C:T").stderr
    r = run("-a", "x.fq", "-d", "y.fa", "-M", "C:T")         # default -S 0 is refused (irreproducible in the reference)
    assert r.returncode == 1 and b"-S" in r.stderr
    assert b"convert-from base: C" in r.stdout and b"convert-to base(s):T" in r.stdout   # SetAlign prints to stdout


def test_usage():
    r = run()
    assert r.returncode == 1 and b"Usage:" in r.stderr


# ------------------------------------------------------------------ the read loader alone ($BASAL_PARSE_ONLY, no GPU)

SAM2BAM = os.path.join(helpers.ROOT, "basal_b200", "bin", "sam2bam")
FQ2UBAM = os.path.join(helpers.ROOT, "tools", "fq2ubam.py")


def parsed(tmp, *args):
    import sys
    env = dict(os.environ, BASAL_PARSE_ONLY="1")
    r = subprocess.run([helpers.GPU_BIN] + list(args) + ["-M", "C:T", "-S", "7"], capture_output=True, timeout=60, env=env, cwd=tmp)
    assert r.returncode == 0, r.stderr.decode()
    body = r.stdout.decode().split("\n")
    start = next(i for i, l in enumerate(body) if l.startswith("@"))          # SetAlign's two lines come first on stdout
    return "\n".join(body[start:]), r.stderr.decode()


def make_ubam(tmp, out, *fqs):
    import sys
    sam = os.path.join(tmp, out + ".sam")
    with open(sam, "w") as fh:
        subprocess.check_call([sys.executable, FQ2UBAM] + list(fqs), stdout=fh)
    subprocess.check_call([SAM2BAM, sam, os.path.join(tmp, out)])


@pytest.mark.skipif(not os.path.exists(SAM2BAM), reason="sam2bam not built")
def test_loader_reads_fastq_gz_fasta_and_bam_alike(tmp_path):
    import gzip
    import shutil
    tmp = str(tmp_path)
    g = os.path.join(helpers.GOLDEN, "ct_se")
    shutil.copy(os.path.join(g, "ref.fa"), tmp); shutil.copy(os.path.join(g, "reads.fq"), tmp)
    with open(os.path.join(tmp, "reads.fq"), "rb") as a, gzip.open(os.path.join(tmp, "reads.fq.gz"), "wb") as b:
        b.write(a.read())
    make_ubam(tmp, "reads.bam", os.path.join(tmp, "reads.fq"))
    fq, e1 = parsed(tmp, "-a", "reads.fq", "-d", "ref.fa")
    gz, e2 = parsed(tmp, "-a", "reads.fq.gz", "-d", "ref.fa")
    bam, e3 = parsed(tmp, "-a", "reads.bam", "-d", "ref.fa")
    assert "format: FASTQ" in e1 and "format: FASTQ" in e2 and "format: BAM" in e3
    src = open(os.path.join(tmp, "reads.fq")).read().split("\n")
    src = "\n".join(l.split()[0] if i % 4 == 0 and l else l for i, l in enumerate(src))      # the name is the first token (reads.cpp:56)
    assert fq == src and fq == gz == bam
    cut, _ = parsed(tmp, "-a", "reads.bam", "-d", "ref.fa", "-L", "60")          # -L truncates sequence and qualities (reads.cpp:93)
    recs = cut.strip().split("\n")
    assert all(len(recs[i + 1]) == 60 and len(recs[i + 3]) == 60 for i in range(0, len(recs), 4))
    fa_path = os.path.join(tmp, "reads.fa")
    with open(fa_path, "w") as fh:
        lines = fq.strip().split("\n")
        for i in range(0, len(lines), 4):
            fh.write(">" + lines[i][1:] + "\n" + lines[i + 1] + "\n")
    fa, e4 = parsed(tmp, "-a", "reads.fa", "-d", "ref.fa")
    assert "format: FASTA" in e4
    assert [l for i, l in enumerate(fa.strip().split("\n")) if i % 4 < 2] == [l for i, l in enumerate(fq.strip().split("\n")) if i % 4 < 2]


@pytest.mark.skipif(not os.path.exists(SAM2BAM), reason="sam2bam not built")
def test_loader_paired_bam_is_one_interleaved_file(tmp_path):
    """reads.cpp:88,107: with BAM input both -a and -b name the same interleaved file; file #1 takes the even records."""
    import shutil
    tmp = str(tmp_path)
    g = os.path.join(helpers.GOLDEN, "ag_pe")
    for f in ("ref.fa", "reads_1.fq", "reads_2.fq"):
        shutil.copy(os.path.join(g, f), tmp)
    make_ubam(tmp, "pairs.bam", os.path.join(tmp, "reads_1.fq"), os.path.join(tmp, "reads_2.fq"))
    fq, _ = parsed(tmp, "-a", "reads_1.fq", "-b", "reads_2.fq", "-d", "ref.fa")
    bam, err = parsed(tmp, "-a", "pairs.bam", "-b", "pairs.bam", "-d", "ref.fa")
    assert "format: BAM" in err and fq == bam and fq.count("\n") > 2000


def test_loader_block_boundaries_and_read_window(tmp_path):
    """The splitter cuts the input into blocks of whole records; any block size, -B / -E windows, blank lines, CRLF line
    ends and a missing final newline must give the same reads (reads.cpp:13-40, 42-84)."""
    import shutil
    tmp = str(tmp_path)
    g = os.path.join(helpers.GOLDEN, "ag_pe")
    for f in ("ref.fa", "reads_1.fq", "reads_2.fq"):
        shutil.copy(os.path.join(g, f), tmp)

    def run_parse(args, batch):
        env = dict(os.environ, BASAL_PARSE_ONLY="1", BASAL_BATCH=str(batch))
        r = subprocess.run([helpers.GPU_BIN] + args + ["-M", "C:T", "-S", "7"], capture_output=True, timeout=60, env=env, cwd=tmp)
        assert r.returncode == 0, r.stderr.decode()
        out = r.stdout.decode().split("\n")
        return "\n".join(out[next(i for i, l in enumerate(out) if l.startswith("@")):])
    base = ["-a", "reads_1.fq", "-b", "reads_2.fq", "-d", "ref.fa"]
    whole = run_parse(base, 1 << 20)
    assert whole.count("\n") == 4 * 600
    for batch in (1, 7, 64, 299, 300):
        assert run_parse(base, batch) == whole
    recs = whole.strip().split("\n")
    win = run_parse(base + ["-B", "11", "-E", "25"], 4)                      # pairs 11..25
    assert win.strip().split("\n") == recs[8 * 10: 8 * 25]
    # blank lines before headers, CRLF line ends, no newline at the end of the file
    src = open(os.path.join(tmp, "reads_1.fq")).read().rstrip("\n").split("\n")
    messy = []
    for i in range(0, len(src), 4):
        messy += ([""] if (i // 4) % 3 == 0 else []) + [src[i] + "\r", src[i + 1] + "\r", src[i + 2], src[i + 3] + "\r"]
    open(os.path.join(tmp, "messy.fq"), "w").write("\n".join(messy))
    clean = run_parse(["-a", "reads_1.fq", "-d", "ref.fa"], 50)
    assert run_parse(["-a", "messy.fq", "-d", "ref.fa"], 13) == clean
    import gzip
    with gzip.open(os.path.join(tmp, "messy.fq.gz"), "wb") as fh:
        fh.write("\n".join(messy).encode())
    assert run_parse(["-a", "messy.fq.gz", "-d", "ref.fa"], 17) == clean
