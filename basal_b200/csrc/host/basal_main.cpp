// basal_main.cpp — the `basal` command line on top of the C-ABI (include/basal_gpu.h).
//
// Drop-in for the reference binary's process contract (main.cpp:272-655): same flags in both
// `-x v` and `-x=v` forms, same SAM header / records / stderr summary; `.bam` output is written by
// bam_writer.hpp (BGZF blocks compressed by the worker threads; $BASAL_SAMTOOLS=1 pipes through an
// external `samtools view -bS -` like the reference, main.cpp:505).  The mapping itself happens on the GPU(s):
// this file only parses text, trims reads, batches them and prints result records.
//
//   reader thread  ->  batches of reads  ->  worker threads (bsl_align_se/pe on a GPU, then SAM text)
//                                        ->  writer (input order)
//
// Reads shard across GPUs by batch; every GPU holds a replica of the index; results are merged
// in input order, so the output does not depend on the number of GPUs.
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/basal_gpu.h"
#include "bam_writer.hpp"

typedef uint8_t u8; typedef uint16_t u16; typedef uint32_t u32; typedef uint64_t u64;
static const char *kVersion = "1.8.1";

// ------------------------------------------------------------------------------------------------ options
struct Options {
    bsl_params P;
    std::string a, b, d, o, rule, cmdline;
    std::vector<std::string> adapters;
    int verbose = 1, procs = 1;
    bool header = true, unmap = false, outref = false, to_stdout = true;
    u32 read_start = 1, read_end = ~0u, max_readlen = 480;
    int qual_threshold = 0; u8 zero_qual = '!'; int default_qual = 40;
    float kmer_ratio_shown = 5e-7f;
    int out_sam = 1;
};

static time_t g_t0;
static const char *now_str() { static thread_local char buf[64]; time_t t = time(nullptr); char *s = ctime_r(&t, buf); s[strlen(s) - 1] = 0; return s; }
static long secs_passed() { return (long)(time(nullptr) - g_t0); }

static void usage() {
    fprintf(stderr,
        "   ___    __    __    __    _    \n  | |_)  / /\\  ( (`  / /\\  | |   \n  |_|_) /_/--\\ _)_) /_/--\\ |_|__ \n"
        "\nWelcome to use BASAL [Version %s] (B200 GPU build)\n"
        "\nUsage:\tbasal [options]\n  Options for input/output files:\n"
        "       -a  <str>    input reads in FASTA/FASTQ format [Required option]\n"
        "       -b  <str>    input reads which is paired with -a, (default: none, single-end)\n"
        "       -d  <str>    reference sequences in FASTA format [Required option]\n"
        "       -o  <str>    output alignment in SAM/BAM format, if omitted, the output will be written to STDOUT in SAM format.\n"
        "\n  Options for base-conversion:\n"
        "       -M  <str>    the convert-from and convert-to base(s) seperated by ':' [Required option], e.g. C:T, A:G, A:CGT, T:-, G:ACT-\n"
        "\n  Options for alignment:\n"
        "       -v  <float>  maximum percentage/number of mismatch bases in each read. (default: 0.1)\n"
        "       -g  <int>    maximum size of gap (deletion/insertion), <=3 bp. default: 0\n"
        "       -w  <int>    maximum number of equal best hits to count, <=1000\n"
        "       -B  <int>    start from the Nth read or read pair, default: 1\n"
        "       -E  <int>    end at the Nth read or read pair, default: 4,294,967,295\n"
        "       -I  <int>    index interval (1~16), default: 4\n"
        "       -k  <float>  the cut-off ratio for over-represented kmers, default: 5e-07\n"
        "       -s  <int>    seed size (10~16), default: 16.\n"
        "       -S  <int>    seed for random number generation used in selecting multiple hits (non-zero)\n"
        "       -p  <int>    number of host worker threads, default: 1 (GPUs: all visible, or $BASAL_GPUS)\n"
        "\n  Options for pair-end alignment:\n       -m  <int>    minimal insert size allowed, default: 28\n       -x  <int>    maximal insert size allowed, default: 1000\n"
        "\n  Options for reads trimming:\n       -q  <int>    quality threshold in trimming, 0-40, default: 0\n       -z  <int>    base quality, default: 33\n"
        "       -f  <int>    reads containing more than this number of Ns will be skipped, default=5\n       -A  <str>    3' end adapter sequence to be trimmed\n"
        "       -L  <int>    map the first N bases of the read, the max is 480 (default).\n"
        "\n  Options for mapping strand:\n       -n  [0,1,2]  0: directional, 1: non-directional, 2: PBAT. default: 0\n"
        "\n  Options for reporting:\n       -r  [0,1,2]  how to report repeat hits, 0=none; 1=random one; 2=all, default:1.\n"
        "       -R           print corresponding reference sequences in SAM output\n       -u           report unmapped reads\n"
        "       -H           do not print header information in SAM format output\n       -V  [0,1,2]  verbose level\n       -h           help\n\n", kVersion);
    exit(1);
}

// mGetOptions (main.cpp:272-364). Returns 0 or the index of the offending argument.
static int parse_options(int argc, char **argv, Options &O) {
    bsl_params_default(&O.P);
    O.cmdline = argv[0]; for (int i = 1; i < argc; i++) O.cmdline += std::string(" ") + argv[i];
    for (int i = 1; i < argc; i++) {
        const char *f = argv[i];
        if (f[0] != '-') return i;
        const char *val = nullptr;
        auto need = [&]() -> bool { if (f[2] == 0) { if (i + 1 >= argc) return false; val = argv[++i]; return true; } if (f[2] == '=') { val = f + 3; return true; } return false; };
        auto flag = [&]() -> bool { return f[2] == 0; };
        switch (f[1]) {
        case 'a': if (!need()) return i; O.a = val; break;
        case 'b': if (!need()) return i; O.b = val; break;
        case 'd': if (!need()) return i; O.d = val; break;
        case 'o': if (!need()) return i; O.o = val; O.to_stdout = false; break;
        case 'M': if (!need()) return i; O.rule = val; break;
        case 's': { if (!need()) return i; int n = atoi(val);
            if (n > 16 || n < 10) { fprintf(stderr, "seed size must be between 10 and 16\n"); exit(1); }                 // param.cpp:109
            O.P.seed_size = n; O.P.min_read_size = n + O.P.index_interval - 1; break; }                                  // param.cpp:112 (uses the -I seen so far)
        case 'm': if (!need()) return i; O.P.min_insert = atoi(val); break;
        case 'x': if (!need()) return i; O.P.max_insert = atoi(val); break;
        case 'n': if (!need()) return i; O.P.chains = atoi(val); break;
        case 'g': if (!need()) return i; O.P.gap = atoi(val);
            if (O.P.gap > 3) { fprintf(stderr, "warning: gap length exceeds max value:3\n"); O.P.gap = 3; } break;
        case 'r': if (!need()) return i; O.P.report_repeat_hits = atoi(val);
            if (O.P.report_repeat_hits > 2) { fprintf(stderr, "invalid -r value: %u, must be 0, 1, or 2.\n", O.P.report_repeat_hits); exit(1); } break;
        case 'V': if (!need()) return i; O.verbose = atoi(val);
            if (O.verbose > 2 || O.verbose < 0) { fprintf(stderr, "invalid -V value: %d, must be 0, 1, or 2.\n", O.verbose); exit(1); } break;
        case 'I': if (!need()) return i; O.P.index_interval = atoi(val);
            if (O.P.index_interval > 16) { fprintf(stderr, "index interval exceeds max value:16\n"); exit(1); }
            if (O.P.index_interval < 1) { fprintf(stderr, "index interval must be at least 1\n"); exit(1); } break;
        case 'k': if (!need()) return i; O.P.max_kmer_ratio = (float)atof(val); break;
        case 'v': { if (!need()) return i; double t = atof(val);                                                           // main.cpp:324-338
            if (t < 1.0) { O.P.max_snp_num = (int)(t * 100 + 0.5) + 100; if (O.P.max_snp_num == 100) O.P.max_snp_num = 0; }
            else { O.P.max_snp_num = (int)(t + 0.5); if (O.P.max_snp_num > 15) { fprintf(stderr, "warning: number of mismatches exceeds max value:15\n"); O.P.max_snp_num = 15; } }
            break; }
        case 'w': if (!need()) return i; O.P.max_num_hits = atoi(val);
            if (O.P.max_num_hits > 1000) { fprintf(stderr, "number of multi-hits exceeds max value:1000\n"); exit(1); } break;
        case 'q': if (!need()) return i; O.qual_threshold = atoi(val); break;
        case 'f': if (!need()) return i; O.P.max_ns = atoi(val); break;
        case 'z': if (!need()) return i; O.zero_qual = (u8)atoi(val); break;
        case 'p': if (!need()) return i; O.procs = atoi(val); break;
        case 'A': if (!need()) return i; if (O.adapters.size() < 10) O.adapters.push_back(val); break;
        case 'R': if (!flag()) return i; O.outref = true; break;
        case 'H': if (!flag()) return i; O.header = false; break;
        case 'u': if (!flag()) return i; O.unmap = true; break;
        case 'B': if (!need()) return i; O.read_start = std::max(atoi(val), 1); break;
        case 'E': if (!need()) return i; O.read_end = (u32)atoi(val); break;
        case 'L': if (!need()) return i; O.max_readlen = (u32)atoi(val); break;
        case 'S': if (!need()) return i; O.P.randseed = (u32)atoi(val); break;
        case '3': case 'N': case 'D':
            fprintf(stderr, "option %s (hidden RRBS / 3-letter / N-as-mismatch mode of the reference) is not supported by the GPU build\n", f); exit(1);
        case 'h': usage();
        default: return i;
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ rule tables for text output (param.cpp:163-263)
struct TextRule { u8 code[256]; char letter[8]; char readnts[5]; char from; };

static void make_text_rule(const std::string &rule, TextRule &T, bsl_params &P) {
    static const char NT[5] = {'A', 'C', 'G', 'T', '-'};
    if (rule.size() < 2 || rule[1] != ':') { fprintf(stderr, "invalid -M, ref base(one letter in A/C/G/T) should be assigned first before :\n"); exit(1); }
    char from = (char)toupper(rule[0]);
    if (!strchr("ACGT", from) || !from) { fprintf(stderr, "invalid -M, ref base %c not in A/C/G/T\n", rule[0]); exit(1); }
    printf("[BASAL @%s] convert-from base: %c\n", now_str(), from);                                                    // stdout, like param.cpp:175
    memset(T.readnts, ' ', 5); int cnt = 0;
    for (size_t i = 2; i < rule.size(); i++) {
        char t = (char)toupper(rule[i]);
        bool valid = memchr(NT, t, 5) != nullptr && t != 0, used = memchr(T.readnts, t, 5) != nullptr;
        if (t == from) { fprintf(stderr, "invalid -M, read base %c should not be equal to ref base %c\n", rule[i], from); exit(1); }
        if (!valid) { fprintf(stderr, "invalid -M, read base %c not in A/C/G/T/-\n", rule[i]); exit(1); }
        if (!used && cnt < 5) T.readnts[cnt++] = t;
    }
    printf("[BASAL @%s] convert-to base(s):%c%c%c%c%c\n", now_str(), T.readnts[0], T.readnts[1], T.readnts[2], T.readnts[3], T.readnts[4]);   // param.cpp:200
    fflush(stdout);
    T.from = from; P.from_base = from; memset(P.to_bases, 0, sizeof P.to_bases);
    for (int i = 0, k = 0; i < cnt && k < 6; i++) P.to_bases[k++] = T.readnts[i];
    int code[4] = {-1, -1, -1, -1}; const char *acgt = "ACGT";
    code[strchr(acgt, from) - acgt] = 1;
    if (cnt == 1 && T.readnts[0] != '-') code[strchr(acgt, T.readnts[0]) - acgt] = 3;
    const int spare[3] = {0, 2, 3};
    for (int i = 0, j = 0; i < 4; i++) if (code[i] < 0) code[i] = spare[j++];
    memset(T.code, 0, 256);
    for (int i = 0; i < 4; i++) { T.code[(u8)acgt[i]] = T.code[(u8)tolower(acgt[i])] = (u8)code[i]; T.letter[code[i]] = acgt[i]; T.letter[code[i] + 4] = (char)tolower(acgt[i]); }
}

// ------------------------------------------------------------------------------------------------ input
struct GzReader {      // zlib reads plain and gzip files alike (the reference sniffs 0x1f8b, main.cpp:375-384)
    gzFile f = nullptr; std::vector<char> buf; size_t pos = 0, len = 0; bool eof = false;
    bool open(const std::string &p) { f = gzopen(p.c_str(), "rb"); if (!f) return false; gzbuffer(f, 1 << 20); buf.resize(8 << 20); return true; }
    void close() { if (f) gzclose(f); f = nullptr; }
    bool fill() { if (eof) return false; if (pos < len) { memmove(buf.data(), buf.data() + pos, len - pos); } len -= pos; pos = 0;
        int n = gzread(f, buf.data() + len, (unsigned)(buf.size() - len)); if (n <= 0) { eof = true; return len > 0; } len += (size_t)n; return true; }
    // exactly n bytes (binary input); false when the stream ends first. dst may be null to skip
    bool bytes(void *dst, size_t n) {
        char *d = (char *)dst;
        while (n) {
            if (pos == len && !fill()) return false;
            if (pos == len) return false;
            const size_t k = std::min(n, len - pos);
            if (d) { memcpy(d, buf.data() + pos, k); d += k; }
            pos += k; n -= k;
        }
        return true;
    }
    // next line without the terminator; false at EOF
    bool line(const char *&s, size_t &n) {
        for (;;) {
            char *nl = (char *)memchr(buf.data() + pos, '\n', len - pos);
            if (nl) { s = buf.data() + pos; n = (size_t)(nl - s); pos += n + 1; if (n && s[n - 1] == '\r') n--; return true; }
            if (eof) { if (pos < len) { s = buf.data() + pos; n = len - pos; pos = len; return true; } return false; }
            if (len - pos == buf.size()) buf.resize(buf.size() * 2);
            if (!fill() && pos >= len) return false;
        }
    }
};

struct Reference { std::vector<std::string> names; std::vector<u8> cat; std::vector<u64> off; std::vector<u32> len; u64 total = 0; };

// RefSeq::LoadNextSeq (refbase.cpp:17-61): name = first token after '>', sequence tokens concatenated
static bool load_reference(const std::string &path, Reference &R) {
    GzReader in; if (!in.open(path)) return false;
    const char *s; size_t n;
    while (in.line(s, n)) {
        size_t i = 0; while (i < n && isspace((unsigned char)s[i])) i++;
        if (i >= n) continue;
        if (s[i] == '>') { i++; while (i < n && isspace((unsigned char)s[i])) i++; size_t j = i; while (j < n && !isspace((unsigned char)s[j])) j++;
            R.names.emplace_back(s + i, j - i); R.off.push_back(R.cat.size()); R.len.push_back(0); continue; }
        if (R.names.empty()) continue;
        for (; i < n; i++) if (!isspace((unsigned char)s[i])) { R.cat.push_back((u8)s[i]); R.len.back()++; }
    }
    in.close();
    // sequences with no bases end the reference's loading loop (refbase.cpp:192: `while(LoadNextSeq(...))`)
    for (size_t c = 0; c < R.len.size(); c++) if (R.len[c] == 0) { R.names.resize(c); R.off.resize(c); R.len.resize(c); break; }
    for (u32 l : R.len) R.total += l;
    return true;
}

struct ReadRec { std::string name, seq, qual; u32 raw_len = 0; };

struct ReadFile {
    GzReader in; int format = -1;   // 0 fasta, 1 fastq, 3 bam (unaligned reads in a BAM file, reads.cpp:85-108)
    int mate = 0;                   // BAM input of a paired run: both -a and -b name the same interleaved file; file #1 takes
                                    // records 0, 2, 4, ... and file #2 records 1, 3, 5, ... (reads.cpp:88,107)
    std::vector<char> rec;
    bool open(const std::string &p) {
        if (!in.open(p)) return false;
        in.fill(); size_t i = 0;
        if (in.len >= 4 && memcmp(in.buf.data(), "BAM\1", 4) == 0) {                 // zlib has already undone the BGZF layer
            format = -1; int32_t l_text = 0, n_ref = 0;
            if (!in.bytes(nullptr, 4) || !in.bytes(&l_text, 4) || l_text < 0 || !in.bytes(nullptr, (size_t)l_text) || !in.bytes(&n_ref, 4) || n_ref < 0) return true;
            for (int32_t k = 0; k < n_ref; k++) { int32_t l_name = 0; if (!in.bytes(&l_name, 4) || l_name < 0 || !in.bytes(nullptr, (size_t)l_name + 4)) return true; }
            format = 3; return true;
        }
        while (i < in.len && isspace((unsigned char)in.buf[i])) i++;
        if (i < in.len && in.buf[i] == '>') format = 0; else if (i < in.len && in.buf[i] == '@') format = 1; else format = -1;
        return true;
    }
    const char *format_name() const { return format == 3 ? "BAM" : format == 1 ? "FASTQ" : "FASTA"; }
    bool bam_record(ReadRec *r, const Options &O) {                                   // one alignment record; r == null skips it
        int32_t bs = 0; if (!in.bytes(&bs, 4) || bs < 32) return false;
        rec.resize((size_t)bs); if (!in.bytes(rec.data(), (size_t)bs)) return false;
        if (!r) return true;
        const unsigned char *b = (const unsigned char *)rec.data();
        const u32 l_name = b[8], n_cig = (u32)b[12] | ((u32)b[13] << 8); u32 l_seq; memcpy(&l_seq, b + 16, 4);
        const size_t o_name = 32, o_seq = o_name + l_name + 4ull * n_cig, o_qual = o_seq + (l_seq + 1) / 2;
        if (o_qual + l_seq > (size_t)bs || l_name == 0) return false;
        r->name.assign((const char *)b + o_name, strnlen((const char *)b + o_name, l_name));
        const u32 n = std::min<u32>(l_seq, (u32)O.max_readlen);                           // reads.cpp:93
        r->seq.resize(n); r->qual.resize(n);
        for (u32 i = 0; i < n; i++) {
            r->seq[i] = "=ACMGRSVTWYHKDBN"[(b[o_seq + i / 2] >> ((i & 1) ? 0 : 4)) & 15];  // bam_nt16_rev_table (reads.cpp:103)
            r->qual[i] = (char)(b[o_qual + i] + 33);
        }
        return true;
    }
    // ReadClass::LoadBatchReads (reads.cpp:42-84), line oriented
    bool next(ReadRec &r, const Options &O) {
        if (format == 3) {
            if (mate == 2 && !bam_record(nullptr, O)) return false;
            if (!bam_record(&r, O)) return false;
            if (mate == 1 && !bam_record(nullptr, O)) return false;
            return true;
        }
        const char *s; size_t n;
        do { if (!in.line(s, n)) return false; } while (n == 0);
        size_t i = 1; while (i < n && isspace((unsigned char)s[i])) i++; size_t j = i; while (j < n && !isspace((unsigned char)s[j])) j++;
        r.name.assign(s + i, j - i);
        if (!in.line(s, n)) return false;
        while (n && isspace((unsigned char)s[n - 1])) n--;
        r.seq.assign(s, n);
        if (format == 1) { if (!in.line(s, n)) return false; if (!in.line(s, n)) return false; while (n && isspace((unsigned char)s[n - 1])) n--; r.qual.assign(s, n); }
        else r.qual.assign(r.seq.size(), (char)(O.zero_qual + O.default_qual));
        if (r.seq.size() > O.max_readlen) { r.seq.erase(O.max_readlen); if (r.qual.size() > O.max_readlen) r.qual.erase(O.max_readlen); }
        return true;
    }
};

// ------------------------------------------------------------------------------------------------ trimming (align.cpp:51-76, 418-435)
static void trim_read(ReadRec &r, const Options &O, bool &too_short) {
    too_short = false;
    r.raw_len = (u32)r.seq.size();
    bool cut = false;
    for (size_t a = 0; a < O.adapters.size() && !cut; a++) {                                          // TrimAdapter
        const std::string &ad = O.adapters[a];
        for (u32 pos = O.P.seed_size + O.P.index_interval - 1; (size_t)pos + 4 < r.seq.size() && r.seq.size() >= 4; pos++) {
            u32 m0 = 0, k = 0;
            for (; k < ad.size() && k < 15 && pos + k < r.seq.size(); k++) if ((m0 += (ad[k] != r.seq[pos + k])) > 4) break;
            if (k >= m0 * 5 && k > 3) { r.seq.erase(pos); if (r.qual.size() > pos) r.qual.erase(pos); cut = true; break; }
        }
    }
    // TrimLowQual
    if (r.seq.size() != r.qual.size()) r.qual.assign(r.seq.size(), (char)(O.zero_qual + O.default_qual));
    u8 qual_thres = (u8)(O.zero_qual + O.qual_threshold);
    if (O.zero_qual != '!') { for (char &c : r.qual) c = (char)(c - (O.zero_qual - '!')); qual_thres = (u8)(qual_thres - (O.zero_qual - '!')); }
    if (O.qual_threshold == 0) return;
    size_t i = r.qual.size();
    while (i > 0 && !((u8)r.qual[i - 1] > qual_thres)) i--;
    if (i < O.P.seed_size + O.P.index_interval - 1) { too_short = true; return; }
    r.qual.erase(i); r.seq.erase(i);
}

// ------------------------------------------------------------------------------------------------ SAM text
static const char kRev[256] = {0};
static inline char rev_char(char c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
                 case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a'; default: return 'N'; }
}
static void append_seq(std::string &os, const std::string &s, bool rev) { if (!rev) { os += s; return; } size_t n = os.size(); os.resize(n + s.size()); for (size_t i = 0; i < s.size(); i++) os[n + i] = rev_char(s[s.size() - 1 - i]); }
static void append_qual(std::string &os, const std::string &s, bool rev) { if (!rev) { os += s; return; } os.append(s.rbegin(), s.rend()); }
static void append_u(std::string &os, u64 v) { char b[24]; int n = 0; do { b[n++] = (char)('0' + v % 10); v /= 10; } while (v); while (n) os.push_back(b[--n]); }
static void append_i(std::string &os, long long v) { if (v < 0) { os.push_back('-'); append_u(os, (u64)(-v)); } else append_u(os, (u64)v); }

static void append_cigar(std::string &os, const bsl_hit &h) {                                            // align.cpp:641-643
    int L = h.read_len;
    if (h.gap_size == 0) { append_u(os, (u32)L); os.push_back('M'); }
    else if (h.gap_size > 0) { append_i(os, h.gap_pos); os.push_back('M'); append_i(os, h.gap_size); os.push_back('D'); append_i(os, L - (int)h.gap_pos); os.push_back('M'); }
    else { append_i(os, h.gap_pos); os.push_back('M'); append_i(os, -h.gap_size); os.push_back('I'); append_i(os, L - (int)h.gap_pos + h.gap_size); os.push_back('M'); }
}

struct Formatter {
    const Options &O; const Reference &R; const TextRule &T;
    Formatter(const Options &o, const Reference &r, const TextRule &t) : O(o), R(r), T(t) {}

    void xr(std::string &os, const bsl_hit &h) const {                                                   // align.cpp:646-658
        u32 c = h.chr >> 1; const u8 *q = R.cat.data() + R.off[c]; u32 n = R.len[c];
        auto letter = [&](u32 p) -> char { u8 code = p < n ? T.code[q[p]] : 0; return T.letter[code]; };
        std::string m;
        for (u32 ii = 2; ii > 0; ii--) { if (h.loc < ii) continue; m.push_back((char)(letter(h.loc - ii) + 32)); }
        for (u32 ii = 0; ii < (u32)h.read_len + 2; ii++) m.push_back(letter(h.loc + ii));
        m[m.size() - 1] += 32; m[m.size() - 2] += 32;
        os += "\tXR:Z:"; os += m;
    }
    void unaligned(std::string &os, const ReadRec &r, int flag) const {
        os += r.name; os.push_back('\t'); append_i(os, flag); os += "\t*\t0\t0\t*\t*\t0\t0\t"; os += r.seq; os.push_back('\t'); os += r.qual; os.push_back('\n');
    }
    // s_OutHit (align.cpp:616-669): n<0 filtered, n==0 unmapped, else hit count
    void single(std::string &os, const ReadRec &r, u32 readset, const bsl_hit &h, int n) const {
        int flag = 0x40 * (int)readset;
        if (n <= 0) { if (!O.unmap) return; unaligned(os, r, flag | (n < 0 ? 0x204 : 0x4)); return; }
        bool rev = (h.read_chain ^ (h.chr & 1)) != 0;
        if (n > 1) flag |= 0x100;
        if (rev) flag |= 0x10;
        os += r.name; os.push_back('\t'); append_i(os, flag); os.push_back('\t'); os += R.names[h.chr >> 1]; os.push_back('\t'); append_u(os, h.loc + 1);
        os += "\t255\t"; append_cigar(os, h); os += "\t*\t0\t0\t"; append_seq(os, r.seq, rev); os.push_back('\t'); append_qual(os, r.qual, rev);
        os += "\tNM:i:"; append_i(os, h.nm);
        if (O.outref) xr(os, h);
        os += "\tZS:Z:"; os.push_back("+-"[h.chr & 1]); os.push_back("+-"[h.read_chain]); os.push_back('\n');
    }
    // s_OutHitPair (pairs.cpp:307-416)
    void pair(std::string &os, const ReadRec &ra, const ReadRec &rb, const bsl_hit &a, const bsl_hit &b, u32 chain, u32 insert, int n) const {
        for (int side = 0; side < 2; side++) {
            const bsl_hit &me = side ? b : a, &mate = side ? a : b; const ReadRec &r = side ? rb : ra;
            u32 ch = side ? !chain : chain; bool rev = (ch ^ (me.chr & 1)) != 0;
            int flag = 0x3; if (n > 1) flag |= 0x100; long long ins;
            if (rev) { flag |= 0x10; ins = -(long long)(int)insert; } else { flag |= 0x20; ins = (int)insert; }
            flag |= 0x40 * (side + 1);
            os += r.name; os.push_back('\t'); append_i(os, flag); os.push_back('\t'); os += R.names[me.chr >> 1]; os.push_back('\t'); append_u(os, me.loc + 1);
            os += "\t255\t"; append_cigar(os, me); os += "\t=\t"; append_u(os, mate.loc + 1); os.push_back('\t'); append_i(os, ins); os.push_back('\t');
            append_seq(os, r.seq, rev); os.push_back('\t'); append_qual(os, r.qual, rev); os += "\tNM:i:"; append_i(os, me.nm);
            if (O.outref) xr(os, me);
            os += "\tZS:Z:"; os.push_back("+-"[me.chr & 1]); os.push_back("+-"[ch]); os.push_back('\n');
        }
    }
    // s_OutHitUnpair (pairs.cpp:418-485)
    void unpair(std::string &os, const ReadRec &r, int side, u32 chain_a, u32 chain_b, int ma, u32 na, const bsl_hit &ha, int mb, const bsl_hit &hb) const {
        int flag = 1 | (0x40 * (side + 1)); bool rev = (chain_a ^ (ha.chr & 1)) != 0;
        if (ma <= 0) {
            if (ma < 0) flag |= 0x204; if (ma == 0) flag |= 0x4;
            if (mb <= 0) { unaligned(os, r, flag | 0x8); return; }
            if (chain_b ^ (hb.chr & 1)) flag |= 0x20;
            os += r.name; os.push_back('\t'); append_i(os, flag); os += "\t*\t0\t0\t*\t"; os += R.names[hb.chr >> 1]; os.push_back('\t'); append_u(os, hb.loc + 1);
            os += "\t0\t"; os += r.seq; os.push_back('\t'); os += r.qual; os.push_back('\n'); return;
        }
        if (ma > 1) flag |= 0x100; if (rev) flag |= 0x10;
        if (mb <= 0) flag |= 0x8; else if (chain_b ^ (hb.chr & 1)) flag |= 0x20;
        os += r.name; os.push_back('\t'); append_i(os, flag); os.push_back('\t'); os += R.names[ha.chr >> 1]; os.push_back('\t'); append_u(os, ha.loc + 1);
        os += "\t255\t"; append_cigar(os, ha);
        if (mb <= 0) os += "\t*\t0\t0\t"; else { os.push_back('\t'); os += R.names[hb.chr >> 1]; os.push_back('\t'); append_u(os, hb.loc + 1); os += "\t0\t"; }
        append_seq(os, r.seq, rev); os.push_back('\t'); append_qual(os, r.qual, rev); os += "\tNM:i:"; append_i(os, na);
        if (O.outref) xr(os, ha);
        os += "\tZS:Z:"; os.push_back("+-"[ha.chr & 1]); os.push_back("+-"[chain_a]); os.push_back('\n');
    }
};

// FixPairReadName (pairs.cpp:487-507)
static void fix_pair_names(std::string &a, std::string &b) {
    if (a == b) return;
    int d = -1; size_t i, n = std::min(a.size(), b.size());
    for (i = 0; i < n; i++) { if (a[i] != b[i]) break; else if (isdigit((unsigned char)a[i])) d = (int)i; }
    if (i > 0) { if (d < 0) d = (int)i - 1; a.erase(d + 1); b.erase(d + 1); }
    else { fprintf(stderr, "Error: Paired reads name not match:\n%s\n%s\n", a.c_str(), b.c_str()); exit(1); }
}

// ------------------------------------------------------------------------------------------------ pipeline
struct Batch {
    u64 ticket = 0; u32 first_index = 0;
    std::vector<ReadRec> a, b;
    std::string text;
};

struct Counters { u64 al = 0, un = 0, mu = 0, pal = 0, pun = 0, pmu = 0, aal = 0, aun = 0, amu = 0, bal = 0, bun = 0, bmu = 0; };

struct Pipeline {
    const Options &O; const Reference &R; const TextRule &T; bool pe;
    std::vector<bsl_ctx *> ctx;
    ReadFile fa, fb;
    std::mutex in_mu, out_mu; std::condition_variable out_cv;
    u64 next_ticket = 0, next_write = 0; u32 next_index; bool input_done = false;
    std::map<u64, std::string> done;
    FILE *out = nullptr;
    bool bam_native = false; bam::Refs refs;          // -o x.bam without an external samtools
    Counters total; u64 reads_seen = 0;
    std::atomic<int> failed{0};
    size_t batch_reads;

    Pipeline(const Options &o, const Reference &r, const TextRule &t) : O(o), R(r), T(t), pe(!o.b.empty()), next_index(o.read_start - 1) {
        const char *e = getenv("BASAL_BATCH"); batch_reads = e ? (size_t)atol(e) : (size_t)(pe ? 262144 : 524288);
    }

    bool load(Batch &B) {
        std::lock_guard<std::mutex> g(in_mu);
        if (input_done) return false;
        B.ticket = next_ticket; B.first_index = next_index; B.a.clear(); B.b.clear();
        ReadRec ra, rb;
        while (B.a.size() < batch_reads && next_index < O.read_end) {
            if (!fa.next(ra, O)) { input_done = true; break; }
            if (pe) { if (!fb.next(rb, O)) { input_done = true; break; } B.b.push_back(rb); }
            B.a.push_back(ra); next_index++;
        }
        if (next_index >= O.read_end) input_done = true;
        if (B.a.empty()) return false;
        next_ticket++; reads_seen += B.a.size();
        return true;
    }

    static int mate_count(const bsl_hit &h) { return h.status == BSL_ST_FILTERED ? -1 : (h.status == BSL_ST_UNMAPPED ? 0 : (int)h.n_hits); }

    void process(Batch &B, bsl_ctx *c, Counters &cn) {
        const size_t n = B.a.size(); Formatter F(O, R, T);
        std::vector<u8> short_a(n, 0), short_b(pe ? n : 0, 0);
        auto pack = [&](std::vector<ReadRec> &v, std::vector<u8> &too_short, std::vector<u8> &bases, std::vector<u64> &off, std::vector<u16> &raw) {
            off.resize(n + 1); raw.resize(n); size_t tot = 0;
            for (size_t i = 0; i < n; i++) { bool ts; trim_read(v[i], O, ts); too_short[i] = ts; tot += v[i].seq.size(); }
            bases.resize(tot + 1); size_t p = 0;
            for (size_t i = 0; i < n; i++) { off[i] = p; const std::string &s = too_short[i] ? std::string() : v[i].seq; memcpy(bases.data() + p, s.data(), s.size()); p += s.size(); raw[i] = (u16)v[i].raw_len; }
            off[n] = p;
        };
        std::vector<u8> ba, bb; std::vector<u64> oa, ob; std::vector<u16> rwa, rwb;
        pack(B.a, short_a, ba, oa, rwa);
        if (pe) { pack(B.b, short_b, bb, ob, rwb); for (size_t i = 0; i < n; i++) fix_pair_names(B.a[i].name, B.b[i].name); }
        bsl_batch qa; memset(&qa, 0, sizeof qa); qa.n = (u32)n; qa.readset = pe ? 1 : 0; qa.bases = ba.data(); qa.offsets = oa.data(); qa.first_index = B.first_index; qa.raw_len = rwa.data();
        std::vector<bsl_hit> ha(n), hb(pe ? n : 0); std::vector<bsl_pair> hp(pe ? n : 0);
        const bool all = O.P.report_repeat_hits == 2;
        std::vector<bsl_hit> alla, allb; u64 n_all = 0; u64 all_cap = all ? std::max<u64>(n * 8, 1u << 20) : 0;
        int rc;
        for (;;) {
            if (all) { alla.resize(all_cap); if (pe) allb.resize(all_cap); }
            if (!pe) rc = bsl_align_se(c, &qa, ha.data(), all ? alla.data() : nullptr, all_cap, &n_all);
            else { bsl_batch qb = qa; qb.readset = 2; qb.bases = bb.data(); qb.offsets = ob.data(); qb.raw_len = rwb.data();
                rc = bsl_align_pe(c, &qa, &qb, ha.data(), hb.data(), hp.data(), all ? alla.data() : nullptr, all ? allb.data() : nullptr, all_cap, &n_all); }
            if (rc == 0 && all && n_all > all_cap) { all_cap = n_all + 16; continue; }      // grow the all-hits buffer and redo the batch
            break;
        }
        if (rc != 0) { fprintf(stderr, "GPU alignment failed (%d): %s\n", rc, bsl_last_error(c)); failed = 1; return; }
        std::string &os = B.text; os.clear(); os.reserve(n * (pe ? 900 : 400));
        for (size_t i = 0; i < n; i++) {
            if (!pe) {                                                                        // StringAlign (align.cpp:583-612)
                const bsl_hit &h = ha[i];
                if (h.status == BSL_ST_FILTERED) F.single(os, B.a[i], 0, h, -1);
                else if (h.status == BSL_ST_UNMAPPED) F.single(os, B.a[i], 0, h, 0);
                else if (h.status == BSL_ST_UNIQUE) { cn.al++; cn.un++; F.single(os, B.a[i], 0, h, 1); }
                else { cn.mu++;
                    if (O.P.report_repeat_hits == 1) { cn.al++; F.single(os, B.a[i], 0, h, (int)h.n_hits); }
                    else if (O.P.report_repeat_hits == 2) { cn.al++; for (u32 k = 0; k < h.n_hits; k++) F.single(os, B.a[i], 0, alla[h.all_first + k], (int)h.n_hits); }
                    else F.single(os, B.a[i], 0, h, 0); }
                continue;
            }
            const bsl_hit &a = ha[i], &b = hb[i]; const bsl_pair &p = hp[i];
            if (a.status == BSL_ST_PAIRED) {                                                  // StringAlignPair (pairs.cpp:204-230)
                cn.pal++; if (p.n_pairs == 1) cn.pun++; else cn.pmu++;
                if (p.n_pairs > 1 && O.P.report_repeat_hits == 2)
                    for (u32 k = 0; k < p.n_pairs; k++) { const bsl_hit &xa = alla[p.all_first + k], &xb = allb[p.all_first + k]; F.pair(os, B.a[i], B.b[i], xa, xb, xa.read_chain, xa.all_first, (int)p.n_pairs); }
                else F.pair(os, B.a[i], B.b[i], a, b, p.chain, p.insert, (int)p.n_pairs);
                continue;
            }
            if (p.n_pairs > 1) cn.pmu++;                                                      // multiple pairs suppressed by -r 0
            // StringAlignUnpair (pairs.cpp:232-305)
            int ma = mate_count(a), mb = mate_count(b);
            int ma1 = (ma > 1 && O.P.report_repeat_hits == 0) ? 0 : ma, mb1 = (mb > 1 && O.P.report_repeat_hits == 0) ? 0 : mb;
            u32 ca = a.read_chain, cb = b.read_chain;
            if (ma <= 0) { if (O.unmap) F.unpair(os, B.a[i], 0, 0, cb, ma, 0, a, mb1, b); }
            else if (ma == 1) { cn.aal++; cn.aun++; F.unpair(os, B.a[i], 0, ca, cb, 1, a.nm, a, mb1, b); }
            else { cn.amu++;
                if (O.P.report_repeat_hits >= 1) { cn.aal++; F.unpair(os, B.a[i], 0, ca, cb, ma, a.nm, a, mb1, b); }   // -r 2 lists every hit in the reference; the GPU build reports the pick
                else if (O.unmap) F.unpair(os, B.a[i], 0, 0, cb, 0, 0, a, mb1, b); }
            if (mb <= 0) { if (O.unmap) F.unpair(os, B.b[i], 1, 0, ca, mb, 0, b, ma1, a); }
            else if (mb == 1) { cn.bal++; cn.bun++; F.unpair(os, B.b[i], 1, cb, ca, 1, b.nm, b, ma1, a); }
            else { cn.bmu++;
                if (O.P.report_repeat_hits >= 1) { cn.bal++; F.unpair(os, B.b[i], 1, cb, ca, mb, b.nm, b, ma1, a); }
                else if (O.unmap) F.unpair(os, B.b[i], 1, 0, ca, 0, 0, b, ma1, a); }
        }
    }

    void emit(Batch &B) {
        std::unique_lock<std::mutex> g(out_mu);
        done[B.ticket] = std::move(B.text);
        while (!done.empty() && done.begin()->first == next_write) {
            const std::string &s = done.begin()->second;
            if (!s.empty()) fwrite(s.data(), 1, s.size(), out);
            done.erase(done.begin()); next_write++;
            if (O.verbose >= 2) fprintf(stderr, "[BASAL @%s] batch %llu finished. %ld secs passed\n", now_str(), (unsigned long long)next_write, secs_passed());
        }
    }

    void worker(int wid) {
        bsl_ctx *c = ctx[wid % ctx.size()]; Counters cn; Batch B;
        while (!failed && load(B)) {
            process(B, c, cn);
            if (bam_native) { std::string blk; if (!bam::text_to_blocks(B.text, refs, blk)) { fprintf(stderr, "\ninternal error: malformed SAM record in BAM conversion\n"); failed = 1; } B.text.swap(blk); }
            emit(B);
        }
        std::lock_guard<std::mutex> g(out_mu);
        total.al += cn.al; total.un += cn.un; total.mu += cn.mu; total.pal += cn.pal; total.pun += cn.pun; total.pmu += cn.pmu;
        total.aal += cn.aal; total.aun += cn.aun; total.amu += cn.amu; total.bal += cn.bal; total.bun += cn.bun; total.bmu += cn.bmu;
    }
};

static void check_input(const std::string &path, const char *msg) {
    FILE *f = fopen(path.c_str(), "rb"); if (!f) { fprintf(stderr, "\n%s%s\n", msg, path.c_str()); exit(1); } fclose(f);
}

int main(int argc, char **argv) {
    if (argc == 1) usage();
    g_t0 = time(nullptr);
    Options O;
    int bad = parse_options(argc, argv, O);
    if (bad) { fprintf(stderr, "unknown option: %s\n", argv[bad]); exit(bad); }
    if (O.rule.empty()) { fprintf(stderr, "\n-M option is required\n"); exit(1); }
    TextRule T; make_text_rule(O.rule, T, O.P);
    if (O.P.randseed == 0) {
        fprintf(stderr, "error: -S 0 (seed from the system clock) makes the reference irreproducible; the GPU build requires a non-zero -S\n"); exit(1);
    }
    if (O.verbose >= 2) fprintf(stderr, "\nBASAL v%s\n", kVersion);
    if (O.verbose >= 1) fprintf(stderr, "[BASAL @%s] loading reference file: %s", now_str(), O.d.c_str());
    check_input(O.d, "failed to open reference file (check -d option): ");
    Reference R;
    if (!load_reference(O.d, R)) { fprintf(stderr, "\nfailed to open reference file (check -d option): %s\n", O.d.c_str()); exit(1); }
    if (R.names.empty()) { fprintf(stderr, "\t(format: unknown)\nreference must be in FASTA format.\n"); exit(1); }
    if (O.verbose >= 1) fprintf(stderr, " \t(format: FASTA)\n[BASAL @%s] %zu reference seqs loaded, total size %llu bp. %ld secs passed\n", now_str(), R.names.size(), (unsigned long long)R.total, secs_passed());

    // ---- $BASAL_PARSE_ONLY: run the read loader alone (no GPU) and print what it parsed as FASTQ, mates interleaved:
    //      the CPU test hook of the host loader (FASTA / FASTQ / gz / BAM input, -L)
    if (getenv("BASAL_PARSE_ONLY")) {
        ReadFile fa, fb; const bool pe2 = !O.b.empty();
        if (!fa.open(O.a) || fa.format < 0 || (pe2 && (!fb.open(O.b) || fb.format != fa.format))) { fprintf(stderr, "\t(format: unknown)\nUnknown input format.\n"); return 1; }
        if (pe2) { fa.mate = 1; fb.mate = 2; }
        fprintf(stderr, "format: %s\n", fa.format_name());
        const bool count_only = strcmp(getenv("BASAL_PARSE_ONLY"), "count") == 0;      // loader throughput without the printing
        ReadRec ra, rb; std::string o; u64 n_rec = 0, n_base = 0;
        while (fa.next(ra, O)) {
            n_rec++; n_base += ra.seq.size();
            if (!count_only) o += "@" + ra.name + "\n" + ra.seq + "\n+\n" + ra.qual + "\n";
            if (pe2) { if (!fb.next(rb, O)) break; n_rec++; n_base += rb.seq.size(); if (!count_only) o += "@" + rb.name + "\n" + rb.seq + "\n+\n" + rb.qual + "\n"; }
            if (o.size() > (1u << 20)) { fwrite(o.data(), 1, o.size(), stdout); o.clear(); }
        }
        fwrite(o.data(), 1, o.size(), stdout);
        fprintf(stderr, "parsed %llu reads, %llu bases\n", (unsigned long long)n_rec, (unsigned long long)n_base);
        return 0;
    }

    // ---- GPUs: every visible device holds a replica of the index
    int ngpu = 1; { const char *e = getenv("BASAL_GPUS"); if (e) ngpu = std::max(1, atoi(e)); else { const char *v = getenv("BASAL_ALL_GPUS"); if (v && atoi(v)) ngpu = 64; } }
    Pipeline P(O, R, T);
    {
        std::vector<std::thread> th; std::vector<int> rcs; std::vector<bsl_ctx *> cs;
        for (int g = 0; g < ngpu; g++) {
            bsl_ctx *c = nullptr; int rc = bsl_ctx_create(&c, g, &O.P);
            if (rc != 0) { if (g == 0) { fprintf(stderr, "cannot open GPU 0 (%d): %s\n", rc, bsl_last_error(nullptr)); exit(1); } break; }
            cs.push_back(c);
        }
        rcs.assign(cs.size(), 0);
        for (size_t g = 0; g < cs.size(); g++) th.emplace_back([&, g]() { rcs[g] = bsl_index_build(cs[g], R.cat.data(), R.off.data(), R.len.data(), (u32)R.len.size()); });
        for (auto &t : th) t.join();
        for (size_t g = 0; g < cs.size(); g++) if (rcs[g] != 0) { fprintf(stderr, "index build failed on GPU %zu (%d): %s\n", g, rcs[g], bsl_last_error(cs[g])); exit(1); }
        P.ctx = cs;
    }
    if (O.verbose >= 1) fprintf(stderr, "[BASAL @%s] create seed table. %ld secs passed\n", now_str(), secs_passed());

    // ---- RunProcess (main.cpp:409-614)
    if (O.o.size() > 4) { if (O.o.compare(O.o.size() - 4, 4, ".sam") == 0) O.out_sam = 1; else if (O.o.compare(O.o.size() - 4, 4, ".bam") == 0) O.out_sam = 2; }
    if (O.verbose >= 2) {
        if (O.P.max_snp_num < 100) fprintf(stderr, "\tmax number of mismatches: %u", O.P.max_snp_num); else fprintf(stderr, "\tmax number of mismatches: read_length * %u%% ", O.P.max_snp_num - 100);
        fprintf(stderr, " \tmax gap size: %u \tkmer cut-off ratio: %g \tmax multi-hits: %u\n", O.P.gap, (double)O.P.max_kmer_ratio, O.P.max_num_hits);
        fprintf(stderr, "\tquality cutoff: %d \tbase quality char: '%c' \tmax Ns: %u\n", O.qual_threshold, O.zero_qual, O.P.max_ns);
        fprintf(stderr, "\twildcard mapping approach \tseed size: %u \tindex interval: %u\n", O.P.seed_size, O.P.index_interval);
    }
    const bool pe = !O.b.empty();
    if (O.verbose >= 1) fprintf(stderr, "[BASAL @%s] %s alignment(%zu GPU(s), %d host threads),\n", now_str(), pe ? "Pair-end" : "Single-end", P.ctx.size(), std::max(O.procs, 1));
    check_input(O.a, pe ? "failed to open read file #1 (check -a option): " : "failed to open read file (check -a option): ");
    if (!P.fa.open(O.a) || P.fa.format < 0) { fprintf(stderr, "\t(format: unknown)\nUnknown input format.\n"); exit(1); }
    if (pe) P.fa.mate = 1;
    if (O.verbose >= 1) fprintf(stderr, "\tInput read file%s: %s \t(format: %s)\n", pe ? " #1" : "", O.a.c_str(), P.fa.format_name());
    if (pe) {
        check_input(O.b, "failed to open read file #2 (check -b option): ");
        if (!P.fb.open(O.b) || P.fb.format < 0) { fprintf(stderr, "\t(format: unknown)\nUnknown input format.\n"); exit(1); }
        if (P.fb.format != P.fa.format) { fprintf(stderr, "Input read file #1 and #2 should be in same format.\n"); exit(1); }
        P.fb.mate = 2;
        if (O.verbose >= 1) fprintf(stderr, "\tInput read file #2: %s \t(format: %s)\n", O.b.c_str(), P.fb.format_name());
    }
    { ReadRec skip; for (u32 i = 1; i < O.read_start; i++) { P.fa.next(skip, O); if (pe) P.fb.next(skip, O); } }     // InitIndex (reads.cpp:13-40)
    bool piped = false;
    if (O.to_stdout) { P.out = stdout; if (O.verbose >= 1) fprintf(stderr, "\tOutput: STDOUT\t (format: SAM)\n"); }
    else {
        if (O.verbose >= 1 || pe) fprintf(stderr, "\tOutput file: %s\t (format: SAM%s)\n", O.o.c_str(), O.out_sam == 2 ? ", automatically convert to BAM" : "");
        FILE *probe = fopen(O.o.c_str(), "wb"); if (!probe) { fprintf(stderr, "\nfailed to open output file (check -o option): %s\n", O.o.c_str()); exit(1); } fclose(probe);
        if (O.out_sam == 2 && getenv("BASAL_SAMTOOLS")) { std::string cmd = "samtools view -bS - >" + O.o; P.out = popen(cmd.c_str(), "w"); piped = P.out != nullptr; }   // main.cpp:505
        if (!P.out) P.out = fopen(O.o.c_str(), "wb");
        if (O.out_sam == 2 && !piped) { P.bam_native = true; for (size_t i = 0; i < R.names.size(); i++) P.refs.add(R.names[i], R.len[i]); }
    }
    static char obuf[8 << 20]; setvbuf(P.out, obuf, _IOFBF, sizeof obuf);
    {
        std::string h;
        if (O.header) {                                                                                  // main.cpp:516-526
            h = "@HD\tVN:1.0\n";
            for (size_t i = 0; i < R.names.size(); i++) { h += "@SQ\tSN:" + R.names[i] + "\tLN:"; append_u(h, R.len[i]); h.push_back('\n'); }
            h += std::string("@PG\tID:BASAL\tVN:") + kVersion + "\tCL:\"" + O.cmdline + "\"\n";
        }
        if (P.bam_native) { const std::string raw = bam::header_bytes(h, P.refs); std::string blk; bam::bgzf_append(raw.data(), raw.size(), blk); fwrite(blk.data(), 1, blk.size(), P.out); }
        else if (!h.empty()) fwrite(h.data(), 1, h.size(), P.out);
    }
    {
        int nw = std::max<int>(O.procs, (int)P.ctx.size() * 3);                 // three lanes per GPU context
        std::vector<std::thread> th; for (int w = 0; w < nw; w++) th.emplace_back([&, w]() { P.worker(w); });
        for (auto &t : th) t.join();
    }
    if (P.bam_native) { std::string e; bam::bgzf_eof(e); fwrite(e.data(), 1, e.size(), P.out); }
    if (piped) pclose(P.out); else if (P.out != stdout) fclose(P.out); else fflush(stdout);
    for (bsl_ctx *c : P.ctx) bsl_ctx_destroy(c);
    if (P.failed) return 2;
    if (O.verbose >= 1) {                                                                                // main.cpp:536-552, 606-612
        const double tot = (double)(P.next_index - (O.read_start - 1)); const Counters &c = P.total; const char *sup = O.P.report_repeat_hits == 0 ? "suppressed " : "";
        if (pe) {
            fprintf(stderr, "[BASAL @%s] total read pairs: %.0f \ttotal time consumed:  %ld secs\n", now_str(), tot, secs_passed());
            fprintf(stderr, "\taligned pairs: %llu (%.1f%%), unique pairs: %llu (%.1f%%), %snon-unique pairs: %llu (%.1f%%)\n", (unsigned long long)c.pal, 100.0 * c.pal / tot, (unsigned long long)c.pun, 100.0 * c.pun / tot, sup, (unsigned long long)c.pmu, 100.0 * c.pmu / tot);
            fprintf(stderr, "\tunpaired read #1: %llu (%.1f%%), unique reads: %llu (%.1f%%), %snon-unique reads: %llu (%.1f%%)\n", (unsigned long long)c.aal, 100.0 * c.aal / tot, (unsigned long long)c.aun, 100.0 * c.aun / tot, sup, (unsigned long long)c.amu, 100.0 * c.amu / tot);
            fprintf(stderr, "\tunpaired read #2: %llu (%.1f%%), unique reads: %llu (%.1f%%), %snon-unique reads: %llu (%.1f%%)\n", (unsigned long long)c.bal, 100.0 * c.bal / tot, (unsigned long long)c.bun, 100.0 * c.bun / tot, sup, (unsigned long long)c.bmu, 100.0 * c.bmu / tot);
        } else {
            fprintf(stderr, "[BASAL @%s] total reads: %.0f \ttotal time:  %ld secs\n", now_str(), tot, secs_passed());
            fprintf(stderr, "\taligned reads: %llu (%.1f%%), unique reads: %llu (%.1f%%), %snon-unique reads: %llu (%.1f%%)\n", (unsigned long long)c.al, 100.0 * c.al / tot, (unsigned long long)c.un, 100.0 * c.un / tot, sup, (unsigned long long)c.mu, 100.0 * c.mu / tot);
        }
    }
    return 0;
}
