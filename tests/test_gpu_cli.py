"""GPU parity tests at the process boundary: our `basal` CLI against the UNMODIFIED reference binary.

Both programs get the same files and flags (fixed -S); SAM output must be byte-identical except the @PG CL line.
The reference binary is oracle/_ref/basal (built by oracle/Makefile.ref; it travels to the GPU box).
"""
import os

import pytest

import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def data(tmp_path_factory):
    root = tmp_path_factory.mktemp("synth")
    out = {}
    for cid, scale in ((1, 0.01), (2, 0.002), (3, 0.002), (4, 0.002), (5, 0.0003)):
        cfg = helpers.synth.baseline_config(cid, scale)
        d = str(root / f"c{cid}")
        paths = helpers.synth.materialise(cfg, d, limit=20000)
        out[cid] = (cfg, d, paths)
    return out


def _inputs(cfg, paths):
    a = ["-a", os.path.basename(paths["a"])]
    if paths["b"]:
        a += ["-b", os.path.basename(paths["b"])]
    return a + ["-d", "ref.fa", "-M", cfg.rule] + list(cfg.flags)


CASES = [
    (1, ["-S", "7"]),
    (1, ["-S", "7", "-u", "-R"]),
    (1, ["-S", "4321", "-n", "1", "-g", "2", "-u"]),
    (1, ["-S", "7", "-H", "-B", "101", "-E", "5000"]),
    (1, ["-S", "7", "-r", "0", "-u", "-w", "2"]),
    (1, ["-S", "7", "-r", "2", "-w", "5"]),
    (1, ["-S=9", "-s=12", "-I=2", "-v=4", "-f", "0", "-u"]),
    (1, ["-S", "7", "-L", "60", "-q", "20", "-A", "AGATCGGAAGAGC", "-u"]),
    (2, ["-S", "7"]),
    (2, ["-S", "7", "-u", "-R"]),
    (2, ["-S", "7", "-m", "250", "-x", "320", "-u"]),
    (2, ["-S", "7", "-n", "1", "-g", "1", "-u", "-w", "3"]),
    (2, ["-S", "7", "-r", "0", "-u"]),
    (3, ["-S", "7"]),
    (3, ["-S", "7", "-n", "1", "-R", "-u"]),
    (4, ["-S", "7"]),
    (4, ["-S", "7", "-n", "1", "-R", "-u"]),
    (5, ["-S", "7", "-u"]),
    (1, ["-S", "7", "-z", "64", "-q", "3", "-u"]),          # -z: quality base other than '!' (qualities are rewritten, align.cpp:56-60)
    (4, ["-S", "7", "-g", "0", "-u"]),
]


@pytest.mark.parametrize("cid,extra", CASES)
def test_cli_matches_reference_binary(data, cid, extra):
    assert helpers.have_ref(), "oracle/_ref/basal missing: run make -f oracle/Makefile.ref"
    cfg, d, paths = data[cid]
    args = _inputs(cfg, paths) + extra
    want = helpers.run_cli(helpers.REF_BIN, args + ["-p", "1"], d, "ref.sam")
    got = helpers.run_cli(helpers.GPU_BIN, args, d, "gpu.sam")
    if got != want:
        g, w = got.splitlines(), want.splitlines()
        assert len(g) == len(w), f"{len(g)} lines vs {len(w)} expected"
        for i, (x, y) in enumerate(zip(g, w)):
            assert x == y, f"line {i}:\n got {x}\nwant {y}"


def test_cli_fasta_reads_and_stdout(data, tmp_path):
    """FASTA read input (qualities synthesised as 'I') and SAM on stdout."""
    import subprocess
    cfg, d, paths = data[1]
    fa = os.path.join(d, "reads.fa")
    with open(paths["a"]) as fq, open(fa, "w") as out:
        for i, line in enumerate(fq):
            if i >= 4000:
                break
            if i % 4 == 0:
                out.write(">" + line[1:])
            elif i % 4 == 1:
                out.write(line)
    args = ["-a", "reads.fa", "-d", "ref.fa", "-M", "C:T", "-S", "7", "-u"]
    want = helpers.run_cli(helpers.REF_BIN, args + ["-p", "1"], d, "ref_fa.sam")
    p = subprocess.run([helpers.GPU_BIN] + args, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    lines = p.stdout.decode().splitlines(keepends=True)
    assert lines[0].startswith("[BASAL @") and "convert-from base: C" in lines[0]      # SetAlign prints to stdout (param.cpp:175,200)
    assert "convert-to base(s):T" in lines[1]
    got = "".join(l for l in lines[2:] if not l.startswith("@PG"))
    assert got == want


def test_cli_bam_output(data):
    """`-o x.bam` is written natively (bam_writer.hpp): decoded, it holds the header and records of the SAM run."""
    from test_bam_writer import decode_bam
    cfg, d, paths = data[2]
    args = _inputs(cfg, paths) + ["-S", "7", "-u", "-R"]
    sam = helpers.run_cli(helpers.GPU_BIN, args, d, "bam_cmp.sam")            # without the @PG line
    import subprocess
    subprocess.run([helpers.GPU_BIN] + args + ["-o", "bam_cmp.bam"], cwd=d, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    text, refs, recs = decode_bam(os.path.join(d, "bam_cmp.bam"))
    hdr = [l for l in sam.splitlines() if l.startswith("@")]
    body = [l for l in sam.splitlines() if l and not l.startswith("@")]
    assert [l for l in text.splitlines() if not l.startswith("@PG")] == hdr
    assert len(recs) == len(body) > 1000
    for r, l in zip(recs, body):
        f = l.split("\t")
        assert [r["name"], r["flag"], r["rname"], r["pos"], r["mapq"], r["cigar"], r["pnext"], r["tlen"], r["seq"], r["qual"], r["tags"]] == \
               [f[0], int(f[1]), f[2], int(f[3]), int(f[4]), f[5], int(f[7]), int(f[8]), f[9].upper(), f[10], f[11:]], l


def _golden_cases():
    import json
    spec = json.load(open(os.path.join(helpers.GOLDEN, "cases.json")))
    return sorted(spec)


@pytest.mark.parametrize("case", _golden_cases())
def test_cli_matches_golden(case, tmp_path):
    """The CUDA `basal` on the committed golden inputs: SAM identical to what the reference binary printed
    (tests/golden/make_golden.py), including the seed-geometry extremes (-s 10 -I 1, -I 16, -k 0.5)."""
    import json
    import shutil
    g = os.path.join(helpers.GOLDEN, case); tmp = str(tmp_path)
    spec = json.load(open(os.path.join(helpers.GOLDEN, "cases.json")))[case]
    for f in os.listdir(g):
        if f != "expected.sam":
            shutil.copy(os.path.join(g, f), tmp)
    got = helpers.run_cli(helpers.GPU_BIN, spec["args"], tmp, "out.sam")
    assert got == open(os.path.join(g, "expected.sam")).read()


@pytest.mark.parametrize("case", ["ct_se", "ag_pe"])
def test_cli_bam_read_input(case, tmp_path):
    """Unaligned BAM as read input (reads.cpp:85-108; a paired run names the same interleaved file twice). The expected
    SAM is the committed golden one: the reference binary gives the same output for BAM and FASTQ input (checked in the
    build container with oracle/_ref/basal on BAM files written by sam2bam)."""
    import json
    import shutil
    import subprocess
    import sys
    g = os.path.join(helpers.GOLDEN, case); tmp = str(tmp_path)
    spec = json.load(open(os.path.join(helpers.GOLDEN, "cases.json")))[case]
    shutil.copy(os.path.join(g, "ref.fa"), tmp)
    fqs = [os.path.join(g, a) for a in spec["args"] if a.endswith(".fq")]
    with open(os.path.join(tmp, "u.sam"), "w") as fh:
        subprocess.check_call([sys.executable, os.path.join(helpers.ROOT, "tools", "fq2ubam.py")] + fqs, stdout=fh)
    subprocess.check_call([os.path.join(helpers.ROOT, "basal_b200", "bin", "sam2bam"), os.path.join(tmp, "u.sam"), os.path.join(tmp, "reads.bam")])
    args = ["reads.bam" if a.endswith(".fq") else a for a in spec["args"]]
    got = helpers.run_cli(helpers.GPU_BIN, args, tmp, "out.sam")
    want = "".join(l for l in open(os.path.join(g, "expected.sam")) if not l.startswith("@PG"))
    assert got == want


def test_cli_error_contract(data):
    import subprocess
    cfg, d, paths = data[1]
    r = subprocess.run([helpers.GPU_BIN, "-a", "reads.fq", "-d", "ref.fa", "-Q"], cwd=d, capture_output=True)
    assert r.returncode == 5 and b"unknown option: -Q" in r.stderr                       # exit(argv index), main.cpp:621-624
    r = subprocess.run([helpers.GPU_BIN, "-a", "reads.fq", "-d", "ref.fa"], cwd=d, capture_output=True)
    assert r.returncode == 1 and b"-M option is required" in r.stderr
    r = subprocess.run([helpers.GPU_BIN, "-a", "reads.fq", "-d", "ref.fa", "-M", "C:C"], cwd=d, capture_output=True)
    assert r.returncode == 1 and b"should not be equal to ref base" in r.stderr
    r = subprocess.run([helpers.GPU_BIN, "-a", "nope.fq", "-d", "ref.fa", "-M", "C:T", "-S", "3"], cwd=d, capture_output=True)
    assert r.returncode == 1 and b"failed to open read file" in r.stderr


@pytest.mark.parametrize("case,batch,procs", [("ct_se_mixed", 37, 4), ("ag_pe_mixed", 23, 3), ("ct_se_mixed_s12", 50, 2), ("ag_pe_r2", 41, 4), ("ct_se_w2_r2", 29, 5), ("ct_pe_gap", 64, 3)])
def test_cli_many_batches_match_golden(case, batch, procs, tmp_path):
    """Small batches ($BASAL_BATCH) and several workers: the ticketed in-order output, lanes running concurrently on one
    context, and — for the mixed-length cases — the context reads that carry the aligner state from batch to batch
    (SURVEY trap 3) must not change a byte of the SAM the reference printed with -p 1."""
    import json
    import shutil
    import subprocess
    g = os.path.join(helpers.GOLDEN, case); tmp = str(tmp_path)
    spec = json.load(open(os.path.join(helpers.GOLDEN, "cases.json")))[case]
    for f in os.listdir(g):
        if f != "expected.sam":
            shutil.copy(os.path.join(g, f), tmp)
    env = dict(os.environ, BASAL_BATCH=str(batch))
    p = subprocess.run([helpers.GPU_BIN] + spec["args"] + ["-p", str(procs), "-o", "out.sam"], cwd=tmp, env=env, capture_output=True, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    got = "".join(l for l in open(os.path.join(tmp, "out.sam")) if not l.startswith("@PG"))
    assert got == open(os.path.join(g, "expected.sam")).read()


def test_cli_two_gpus_match_one(data):
    """$BASAL_GPUS=2 inside one process (batches shard over both devices, every device builds its own replica of the
    index): byte-identical SAM. Skipped on a one-GPU box."""
    import subprocess
    try:
        n = len(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.strip().splitlines())
    except Exception:
        n = 0
    if n < 2:
        pytest.skip("needs two GPUs")
    cfg, d, paths = data[2]
    args = _inputs(cfg, paths) + ["-S", "7", "-u"]
    one = helpers.run_cli(helpers.GPU_BIN, args, d, "g1.sam")
    env = dict(os.environ, BASAL_GPUS="2", BASAL_BATCH="1500")
    p = subprocess.run([helpers.GPU_BIN] + args + ["-p", "6", "-o", "g2.sam"], cwd=d, env=env, capture_output=True, timeout=900)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    two = "".join(l for l in open(os.path.join(d, "g2.sam")) if not l.startswith("@PG"))
    assert two == one
