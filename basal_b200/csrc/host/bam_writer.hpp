// bam_writer.hpp — SAM text -> BAM (BGZF) on the host, so that `-o x.bam` does not need an external samtools.
//
// The reference pipes its SAM text through `samtools view -bS - > x.bam` (main.cpp:505,575; SURVEY.md §8f rank 1: at
// GPU mapping speed that single-threaded pipe is the wall-clock limiter). Here every worker thread converts the SAM
// text of its own batch into COMPLETE BGZF blocks (a BGZF block is a self-contained gzip member, so the concatenation
// of per-batch block sequences in input order is a valid BAM), and the ordered writer only appends bytes.
//
// Record encoding follows the SAM/BAM specification v1 as samtools 0.1.18 (the version vendored by the reference)
// writes it: bin from reg2bin(pos, end), integer tags in the smallest type that holds the value, unmapped records with
// refID/pos = -1 (bin 4680).  The compressed bytes differ from samtools' (block boundaries, zlib level); the decoded
// records do not.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace bam {

struct Refs {
    std::vector<std::string> names; std::vector<uint32_t> lens;
    std::unordered_map<std::string, int32_t> id;
    void add(const std::string &n, uint32_t len) { id.emplace(n, (int32_t)names.size()); names.push_back(n); lens.push_back(len); }
    int32_t find(const char *s, size_t n) const { auto it = id.find(std::string(s, n)); return it == id.end() ? -1 : it->second; }
};

inline void put32(std::string &o, uint32_t v) { char b[4] = {(char)v, (char)(v >> 8), (char)(v >> 16), (char)(v >> 24)}; o.append(b, 4); }
inline void put16(std::string &o, uint32_t v) { char b[2] = {(char)v, (char)(v >> 8)}; o.append(b, 2); }

// ---- BGZF: raw bytes -> gzip members of at most 0xff00 input bytes each, with the BC extra field
inline bool bgzf_append(const char *data, size_t n, std::string &out, int level = Z_DEFAULT_COMPRESSION) {
    const size_t kIn = 0xff00;
    std::vector<unsigned char> buf(0x10000);
    for (size_t off = 0; off < n || (n == 0 && off == 0); off += kIn) {
        const size_t len = n - off < kIn ? n - off : kIn;
        z_stream zs; memset(&zs, 0, sizeof zs);
        if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
        zs.next_in = (Bytef *)(data + off); zs.avail_in = (uInt)len;
        zs.next_out = buf.data(); zs.avail_out = (uInt)(buf.size() - 26);
        const int rc = deflate(&zs, Z_FINISH);
        const size_t clen = zs.total_out; deflateEnd(&zs);
        if (rc != Z_STREAM_END) return false;                       // cannot happen: 0xff00 bytes always fit (stored blocks)
        const uint32_t bsize = (uint32_t)(clen + 25);               // total block size - 1
        static const unsigned char hdr[12] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0};
        out.append((const char *)hdr, 12); out.append("BC", 2); put16(out, 2); put16(out, bsize);
        out.append((const char *)buf.data(), clen);
        put32(out, (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef *)(data + off), (uInt)len)); put32(out, (uint32_t)len);
        if (n == 0) break;
    }
    return true;
}
inline void bgzf_eof(std::string &out) {                              // the 28-byte empty block that ends a BAM file
    static const unsigned char e[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 0x42, 0x43, 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    out.append((const char *)e, 28);
}

// ---- header
inline std::string header_bytes(const std::string &sam_header_text, const Refs &R) {
    std::string o = "BAM\1";
    put32(o, (uint32_t)sam_header_text.size()); o += sam_header_text;
    put32(o, (uint32_t)R.names.size());
    for (size_t i = 0; i < R.names.size(); i++) { put32(o, (uint32_t)R.names[i].size() + 1); o += R.names[i]; o.push_back('\0'); put32(o, R.lens[i]); }
    return o;
}

// UCSC binning scheme of the BAM index (SAM specification section 5.3): the smallest bin of the 5-level hierarchy
// (16 kb, 128 kb, 1 Mb, 8 Mb, 64 Mb windows; 512 Mb root) that contains [beg, end). Arithmetic shifts on purpose: an
// unmapped record (pos -1, no CIGAR) lands in bin 4680 like in every samtools-written file.
inline int reg2bin(int32_t beg, int32_t end) {
    static const struct { int shift, first; } level[5] = {{14, 4681}, {17, 585}, {20, 73}, {23, 9}, {26, 1}};
    const int32_t last = end - 1;
    for (const auto &lv : level)
        if ((beg >> lv.shift) == (last >> lv.shift)) return lv.first + (beg >> lv.shift);
    return 0;
}

// ---- one SAM record line (no trailing newline) -> BAM record appended to `o`. Returns false on a malformed line.
inline bool record(const char *p, const char *end, const Refs &R, std::string &o) {
    const char *f[11]; size_t fl[11]; int nf = 0; const char *q = p;
    while (nf < 11) { const char *t = (const char *)memchr(q, '\t', (size_t)(end - q)); f[nf] = q; fl[nf] = (size_t)((t ? t : end) - q); nf++; if (!t) { q = end; break; } q = t + 1; }
    if (nf < 11) return false;
    const char *tags = q;                                             // after the 11th field (or end)
    auto num = [](const char *s, size_t n) { long v = 0; bool neg = n && s[0] == '-'; for (size_t i = neg; i < n; i++) v = v * 10 + (s[i] - '0'); return neg ? -v : v; };
    const uint32_t flag = (uint32_t)num(f[1], fl[1]);
    const int32_t refid = (fl[2] == 1 && f[2][0] == '*') ? -1 : R.find(f[2], fl[2]);
    const int32_t pos = (int32_t)num(f[3], fl[3]) - 1;
    const uint32_t mapq = (uint32_t)num(f[4], fl[4]);
    // CIGAR
    std::vector<uint32_t> cig; int32_t reflen = 0;
    if (!(fl[5] == 1 && f[5][0] == '*')) {
        uint32_t len = 0;
        for (size_t i = 0; i < fl[5]; i++) {
            const char ch = f[5][i];
            if (ch >= '0' && ch <= '9') { len = len * 10 + (uint32_t)(ch - '0'); continue; }
            const char *ops = "MIDNSHP=X"; const char *w = strchr(ops, ch); if (!w) return false;
            const uint32_t op = (uint32_t)(w - ops);
            cig.push_back(len << 4 | op);
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) reflen += (int32_t)len;
            len = 0;
        }
    }
    const int32_t next_ref = (fl[6] == 1 && f[6][0] == '=') ? refid : (fl[6] == 1 && f[6][0] == '*') ? -1 : R.find(f[6], fl[6]);
    const int32_t next_pos = (int32_t)num(f[7], fl[7]) - 1;
    const int32_t tlen = (int32_t)num(f[8], fl[8]);
    const bool no_seq = fl[9] == 1 && f[9][0] == '*';
    const uint32_t l_seq = no_seq ? 0u : (uint32_t)fl[9];
    const int bin = reg2bin(pos, pos + reflen);                      // bam_calend: pos + reference length of the CIGAR (pos for none)
    const size_t at = o.size();
    put32(o, 0);                                                      // block_size, patched below
    put32(o, (uint32_t)refid); put32(o, (uint32_t)pos);
    put32(o, (uint32_t)bin << 16 | mapq << 8 | (uint32_t)(fl[0] + 1));
    put32(o, flag << 16 | (uint32_t)cig.size());
    put32(o, l_seq); put32(o, (uint32_t)next_ref); put32(o, (uint32_t)next_pos); put32(o, (uint32_t)tlen);
    o.append(f[0], fl[0]); o.push_back('\0');
    for (uint32_t c : cig) put32(o, c);
    {
        static const char *k16 = "=ACMGRSVTWYHKDBN";
        unsigned char cur = 0;
        for (uint32_t i = 0; i < l_seq; i++) {
            char ch = f[9][i]; if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 32);
            const char *w = strchr(k16, ch); const unsigned char code = w && ch ? (unsigned char)(w - k16) : 15;
            if (i & 1) { cur |= code; o.push_back((char)cur); } else cur = (unsigned char)(code << 4);
        }
        if (l_seq & 1) o.push_back((char)cur);
    }
    if (fl[10] == 1 && f[10][0] == '*') o.append(l_seq, (char)0xff);
    else { if (fl[10] != l_seq) return false; for (uint32_t i = 0; i < l_seq; i++) o.push_back((char)(f[10][i] - 33)); }
    // optional fields TAG:TYPE:VALUE
    for (const char *t = tags; t < end;) {
        const char *e = (const char *)memchr(t, '\t', (size_t)(end - t)); if (!e) e = end;
        if (e - t >= 5 && t[2] == ':' && t[4] == ':') {
            const char ty = t[3]; const char *v = t + 5; const size_t vn = (size_t)(e - v);
            o.append(t, 2);
            if (ty == 'i') {
                const long x = num(v, vn);
                if (x < 0) { if (x >= -127) { o.push_back('c'); o.push_back((char)x); } else if (x >= -32767) { o.push_back('s'); put16(o, (uint32_t)x); } else { o.push_back('i'); put32(o, (uint32_t)x); } }
                else { if (x <= 255) { o.push_back('C'); o.push_back((char)x); } else if (x <= 65535) { o.push_back('S'); put16(o, (uint32_t)x); } else { o.push_back('I'); put32(o, (uint32_t)x); } }
            } else if (ty == 'A') { o.push_back('A'); o.push_back(vn ? v[0] : ' '); }
            else if (ty == 'f') { o.push_back('f'); const float fv = (float)atof(std::string(v, vn).c_str()); uint32_t u; memcpy(&u, &fv, 4); put32(o, u); }
            else { o.push_back(ty == 'H' ? 'H' : 'Z'); o.append(v, vn); o.push_back('\0'); }
        }
        t = e + 1;
    }
    const uint32_t bs = (uint32_t)(o.size() - at - 4);
    o[at] = (char)bs; o[at + 1] = (char)(bs >> 8); o[at + 2] = (char)(bs >> 16); o[at + 3] = (char)(bs >> 24);
    return true;
}

// SAM text (record lines only, '\n'-terminated) -> complete BGZF blocks appended to `out`
inline bool text_to_blocks(const std::string &sam, const Refs &R, std::string &out) {
    std::string raw; raw.reserve(sam.size());
    const char *p = sam.data(), *end = p + sam.size();
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p)); if (!nl) nl = end;
        if (nl > p && p[0] != '@' && !record(p, nl, R, raw)) return false;
        p = nl + 1;
    }
    return raw.empty() ? true : bgzf_append(raw.data(), raw.size(), out);
}

} // namespace bam
