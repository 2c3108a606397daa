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
