"""CPU tests (-m "not gpu"): the oracle against the golden fixtures made from the reference binary,
and — when the reference binary is present — against the binary itself on fresh seeded cases."""
import json
import os
import subprocess

import pytest

import helpers

CASES = json.load(open(os.path.join(helpers.GOLDEN, "cases.json")))


def _oracle_sam(args, cwd, out):
    if not os.path.exists(helpers.ORACLE_BIN):
        subprocess.check_call(["make", "-s", "-C", os.path.join(helpers.ROOT, "oracle")])
    return helpers.run_cli(helpers.ORACLE_BIN, args, cwd, out)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(name, tmp_path):
    d = os.path.join(helpers.GOLDEN, name)
    got = _oracle_sam(CASES[name]["args"], d, str(tmp_path / "orc.sam"))
    want = open(os.path.join(d, "expected.sam")).read()
    assert got == want


@pytest.mark.skipif(not helpers.have_ref(), reason="reference binary not built here (oracle/Makefile.ref)")
@pytest.mark.parametrize("cid,scale,extra", [
    (1, 0.004, []), (1, 0.004, ["-g", "1", "-n", "1", "-u"]), (2, 0.0006, ["-u", "-R"]),
    (3, 0.0006, ["-w", "4", "-r", "2"]), (4, 0.0006, ["-u"]), (5, 0.0001, ["-r", "0", "-u"]),
])
def test_oracle_matches_reference_binary(cid, scale, extra, tmp_path):
    cfg = helpers.synth.baseline_config(cid, scale)
    paths = helpers.synth.materialise(cfg, str(tmp_path), limit=3000)
    args = ["-a", os.path.basename(paths["a"])] + (["-b", os.path.basename(paths["b"])] if paths["b"] else [])
    args += ["-d", "ref.fa", "-M", cfg.rule] + list(cfg.flags) + ["-S", "7"] + extra
    want = helpers.run_cli(helpers.REF_BIN, args + ["-p", "1"], str(tmp_path), "ref.sam")
    got = _oracle_sam(args, str(tmp_path), "orc.sam")
    assert got == want
