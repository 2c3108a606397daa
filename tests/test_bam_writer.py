"""CPU test of the native BAM writer (basal_b200/csrc/host/bam_writer.hpp, the `-o x.bam` path of the CLI):
every golden SAM of tests/golden is converted with basal_b200/bin/sam2bam, decoded by an independent BGZF/BAM reader
written here from the SAM/BAM specification, and compared field by field with the SAM text."""
import glob
import os
import struct
import subprocess
import zlib

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SAM2BAM = os.path.join(ROOT, "basal_b200", "bin", "sam2bam")
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*", "expected.sam")))


def bgzf_blocks(data: bytes):
    """Yield the decompressed payload of every BGZF block, checking the header fields, CRC32 and ISIZE."""
    pos = 0
    while pos < len(data):
        magic, cm, flg, _mtime, _xfl, _os, xlen = struct.unpack_from("<HBBIBBH", data, pos)
        assert magic == 0x8B1F and cm == 8 and flg == 4 and xlen == 6
        si1, si2, slen, bsize = struct.unpack_from("<BBHH", data, pos + 12)
        assert (si1, si2, slen) == (66, 67, 2)
        cdata = data[pos + 18: pos + bsize + 1 - 8]
        crc, isize = struct.unpack_from("<II", data, pos + bsize + 1 - 8)
        raw = zlib.decompress(cdata, -15)
        assert len(raw) == isize and (zlib.crc32(raw) & 0xFFFFFFFF) == crc and isize <= 0x10000
        yield raw
        pos += bsize + 1


def reg2bin(beg, end):
    end -= 1
    for shift, off in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return off + (beg >> shift)
    return 0


def decode_bam(path):
    data = open(path, "rb").read()
    blocks = list(bgzf_blocks(data))
    assert blocks[-1] == b"" and data.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    raw = b"".join(blocks)
    assert raw[:4] == b"BAM\x01"
    (l_text,) = struct.unpack_from("<i", raw, 4)
    text = raw[8: 8 + l_text].decode()
    p = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", raw, p); p += 4
    refs = []
    for _ in range(n_ref):
        (ln,) = struct.unpack_from("<i", raw, p); p += 4
        name = raw[p: p + ln - 1].decode(); p += ln
        (lr,) = struct.unpack_from("<i", raw, p); p += 4
        refs.append((name, lr))
    recs = []
    while p < len(raw):
        (bs,) = struct.unpack_from("<i", raw, p); p += 4
        end = p + bs
        refid, pos, bmn, fnc, l_seq, nref, npos, tlen = struct.unpack_from("<iiIIiiii", raw, p); p += 32
        l_name, mapq, bin_ = bmn & 0xFF, (bmn >> 8) & 0xFF, bmn >> 16
        flag, n_cig = fnc >> 16, fnc & 0xFFFF
        name = raw[p: p + l_name - 1].decode(); p += l_name
        cig = ""
        reflen = 0
        for _ in range(n_cig):
            (c,) = struct.unpack_from("<I", raw, p); p += 4
            cig += f"{c >> 4}{'MIDNSHP=X'[c & 15]}"
            if (c & 15) in (0, 2, 3, 7, 8):
                reflen += c >> 4
        seq = "".join("=ACMGRSVTWYHKDBN"[(raw[p + i // 2] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq)); p += (l_seq + 1) // 2
        qual = "".join(chr(q + 33) for q in raw[p: p + l_seq]); p += l_seq
        tags = []
        while p < end:
            tag = raw[p: p + 2].decode(); ty = chr(raw[p + 2]); p += 3
            if ty in "cCsSiI":
                fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}[ty]
                (v,) = struct.unpack_from(fmt, raw, p); p += struct.calcsize(fmt)
                # smallest type that holds the value, like samtools
                assert ty == ("C" if 0 <= v <= 255 else "S" if 0 <= v <= 65535 else "I" if v >= 0 else "c" if v >= -127 else "s" if v >= -32767 else "i")
                tags.append(f"{tag}:i:{v}")
            elif ty == "A":
                tags.append(f"{tag}:A:{chr(raw[p])}"); p += 1
            else:
                z = raw.index(b"\0", p); tags.append(f"{tag}:{ty}:{raw[p:z].decode()}"); p = z + 1
        assert p == end
        assert bin_ == reg2bin(pos, pos + reflen)
        recs.append(dict(name=name, flag=flag, rname=refs[refid][0] if refid >= 0 else "*", pos=pos + 1, mapq=mapq, cigar=cig or "*",
                         rnext=(refs[nref][0] if nref >= 0 else "*"), pnext=npos + 1, tlen=tlen,
                         seq=seq or "*", qual=qual if l_seq else "*", tags=tags))
    return text, refs, recs


@pytest.mark.skipif(not os.path.exists(SAM2BAM), reason="basal_b200/bin/sam2bam not built (python __graft_entry__.py)")
@pytest.mark.parametrize("sam", GOLDEN, ids=[os.path.basename(os.path.dirname(g)) for g in GOLDEN])
def test_bam_matches_sam(sam, tmp_path):
    out = str(tmp_path / "x.bam")
    subprocess.check_call([SAM2BAM, sam, out])
    text, refs, recs = decode_bam(out)
    lines = open(sam).read().splitlines()
    hdr = [l for l in lines if l.startswith("@")]
    body = [l for l in lines if l and not l.startswith("@")]
    assert text == "".join(l + "\n" for l in hdr)
    assert refs == [(f[1][3:], int(f[2][3:])) for f in (l.split("\t") for l in hdr) if f[0] == "@SQ"]
    assert len(recs) == len(body) and len(body) > 0
    for r, l in zip(recs, body):
        f = l.split("\t")
        assert [r["name"], r["flag"], r["rname"], r["pos"], r["mapq"], r["cigar"], r["rnext"], r["pnext"], r["tlen"], r["seq"], r["qual"]] == \
               [f[0], int(f[1]), f[2], int(f[3]), int(f[4]), f[5], f[2] if f[6] == "=" else f[6], int(f[7]), int(f[8]), f[9].upper(), f[10]], l   # BAM stores the mate's reference id; "=" is a SAM abbreviation
        assert r["tags"] == f[11:], l


@pytest.mark.skipif(not os.path.exists(SAM2BAM), reason="basal_b200/bin/sam2bam not built (python __graft_entry__.py)")
@pytest.mark.parametrize("sam", GOLDEN, ids=[os.path.basename(os.path.dirname(g)) for g in GOLDEN])
def test_bam_stream_equals_reference_samtools(sam, tmp_path):
    """The decompressed BAM stream must be byte-identical to what the reference's vendored samtools 0.1.18 writes with
    `samtools view -bS` (hashes generated by tests/golden/make_bam_golden.py in the build container)."""
    import gzip
    import hashlib
    import json
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "bam_sha256.json")))[os.path.basename(os.path.dirname(sam))]
    out = str(tmp_path / "x.bam")
    subprocess.check_call([SAM2BAM, sam, out])
    raw = gzip.open(out, "rb").read()
    assert len(raw) == want["bytes"] and hashlib.sha256(raw).hexdigest() == want["sha256"]


def test_unmapped_record_and_large_tags(tmp_path):
    sam = tmp_path / "u.sam"
    sam.write_text("@HD\tVN:1.0\n@SQ\tSN:c1\tLN:100000000\n"
                   "u1\t4\t*\t0\t0\t*\t*\t0\t0\tACGTN\tIIIII\tNM:i:0\n"
                   "m1\t16\tc1\t70000000\t255\t3M2D2M1I4M\t=\t1\t-5\tACGTACGTAC\t*\tNM:i:300\tXX:i:-5\tXY:i:70000\tZS:Z:-+\tXA:A:c\n")
    out = str(tmp_path / "u.bam")
    subprocess.check_call([SAM2BAM, str(sam), out])
    _, _, recs = decode_bam(out)
    assert recs[0]["rname"] == "*" and recs[0]["pos"] == 0 and recs[0]["cigar"] == "*"
    assert recs[1]["cigar"] == "3M2D2M1I4M" and recs[1]["rname"] == "c1" and recs[1]["pos"] == 70000000 and recs[1]["tlen"] == -5
    assert recs[1]["qual"] == chr(255 + 33) * 10                      # missing qualities are 0xff bytes
    assert recs[1]["tags"] == ["NM:i:300", "XX:i:-5", "XY:i:70000", "ZS:Z:-+", "XA:A:c"]
