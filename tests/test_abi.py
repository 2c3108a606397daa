"""CPU tests: the C-ABI library loads, exports every symbol of include/basal_gpu.h, validates parameters
like the reference, and refuses to run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers
from basal_b200 import capi


def _declared():
    text = open(os.path.join(helpers.ROOT, "include", "basal_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bsl_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    api = capi.load_library()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(api.lib, n), f"{n} declared in include/basal_gpu.h but not exported"
    assert api.abi_version() == 1


def test_struct_layouts_match_header(tmp_path):
    """Compile the header with gcc (plain C) and compare sizeof/offsetof with the ctypes / numpy mirrors."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "basal_gpu.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(bsl_params),sizeof(bsl_batch),sizeof(bsl_hit),sizeof(bsl_pair),sizeof(bsl_index_info),sizeof(bsl_stats),'
                   'offsetof(bsl_hit,gap_pos),offsetof(bsl_hit,all_first),offsetof(bsl_params,max_kmer_ratio));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(helpers.ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got == [C.sizeof(capi.Params), C.sizeof(capi.Batch), capi.HIT_DTYPE.itemsize, capi.PAIR_DTYPE.itemsize, C.sizeof(capi.IndexInfo),
                   C.sizeof(capi.Stats), capi.HIT_DTYPE.fields["gap_pos"][1], capi.HIT_DTYPE.fields["all_first"][1], capi.Params.max_kmer_ratio.offset]


def test_parameter_validation_mirrors_reference():
    api = capi.load_library()
    h = C.c_void_p()
    bad = [dict(rule="C:C"), dict(rule="X:T"), dict(rule="C:Z"), dict(s=9), dict(s=17), dict(I=17), dict(w=1001), dict(S=0)]
    for kw in bad:
        p = capi.make_params(**kw)
        rc = api.ctx_create(C.byref(h), 0, C.byref(p))
        assert rc == -1, kw                                         # BSL_EINVAL before any device work
        assert api.last_error(None)
    msg_rule = capi.make_params(rule="C:C")
    api.ctx_create(C.byref(h), 0, C.byref(msg_rule))
    assert b"should not be equal to ref base" in api.last_error(None)


def test_no_cpu_fallback_without_gpu():
    import shutil
    if shutil.which("nvidia-smi") and os.system("nvidia-smi -L >/dev/null 2>&1") == 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.BasalError) as e:
        capi.Context(capi.make_params())
    assert "no CPU fallback" in str(e.value) or "(-2)" in str(e.value)


def test_budget_and_rand_match_oracle_and_known_answers():
    api, orc = capi.load_library(), helpers.oracle_api()
    # FilterReads (align.cpp:548-563): L=100 -v 0.1 -> 10 ; with -g 3 -> 14 ; L=150 -> 15
    assert api.read_budget(C.byref(capi.make_params(v="0.1")), 100, 100) == 10
    assert api.read_budget(C.byref(capi.make_params(v="0.1", g=3)), 100, 100) == 14
    assert api.read_budget(C.byref(capi.make_params(v="0.1")), 150, 150) == 15
    for v in ("0.1", "0.05", "3", "15", "0"):
        for g in (0, 1, 3):
            p = capi.make_params(v=v, g=g)
            for raw in (16, 36, 75, 100, 101, 150, 251, 480):
                for ln in {raw, max(raw - 7, 1), max(raw // 2, 1)}:
                    assert api.read_budget(C.byref(p), raw, ln) == orc.read_budget(C.byref(p), raw, ln)
    # myrand (utilities.cpp:38-48) incl. the 32-bit wrap of S*1000000
    def ref(i, S):
        m = (1 << 64) - 1
        v = ((i + ((S * 1000000) & 0xffffffff)) * 3935559000370003845 + 2691343689449507681) & m
        v ^= v >> 21; v ^= (v << 37) & m; v ^= v >> 4
        v = (v * 4768777513237032717) & m
        v ^= (v << 20) & m; v ^= v >> 41; v ^= (v << 5) & m
        return v & 0xffffffff
    for S in (7, 4321, 99999):
        for i in (0, 1, 2, 49999, 50000, 123456789, 2 ** 31 - 1):
            assert api.myrand(i, S) == ref(i, S) == orc.myrand(i, S)


def test_oracle_rule_tables_known_codes():
    """SetAlign codes: C:T -> A0 C1 G2 T3 ; A:G -> A1 C0 G3 T2 ; A:CGT -> A1 C0 G2 T3 ; T:- -> A0 C2 G3 T1."""
    orc = helpers.oracle_api()
    fn = orc.lib.orc_rule_tables
    fn.argtypes = [C.c_void_p] * 6
    want = {"C:T": (0, 1, 2, 3, 1), "A:G": (1, 0, 3, 2, 1), "A:CGT": (1, 0, 2, 3, 0), "T:-": (0, 2, 3, 1, 0), "G:ACT-": (0, 2, 1, 3, 0)}
    for rule, (a, c, g, t, single) in want.items():
        ctx = helpers.oracle_context(capi.make_params(rule=rule))
        code = np.zeros(256, np.uint8); rcode = np.zeros(256, np.uint8); conv = np.zeros(256, np.uint8); rconv = np.zeros(256, np.uint8)
        letter = C.create_string_buffer(4)
        s = fn(ctx._h, code.ctypes.data, rcode.ctypes.data, conv.ctypes.data, rconv.ctypes.data, C.addressof(letter))
        assert (code[ord("A")], code[ord("C")], code[ord("G")], code[ord("T")]) == (a, c, g, t), rule
        assert (code[ord("a")], code[ord("t")], code[ord("N")]) == (a, t, 0)
        assert (rcode[ord("A")], rcode[ord("C")], rcode[ord("G")], rcode[ord("T")]) == (t, g, c, a)
        assert s == single
        for to in rule[2:].replace("-", ""):
            assert conv[ord(to)] == 1 and rconv[ord("TGCA"["ACGT".index(to)])] == 1
        ctx.close()
