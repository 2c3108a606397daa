// basal_main.cpp — the `basal` command line on top of the C-ABI (include/basal_gpu.h).
//
// Drop-in for the reference binary's process contract (main.cpp:272-655): same flags in both
// `-x v` and `-x=v` forms, same SAM header / records / stderr summary; `.bam` output is written by
// bam_writer.hpp (BGZF blocks compressed by the worker threads; $BASAL_SAMTOOLS=1 pipes through an
// external `samtools view -bS -` like the reference, main.cpp:505).  The mapping itself happens on the GPU(s):
// this file only parses text, trims reads, batches them and prints result records.
//
//   splitter thread per input file -> text blocks of whole records -> worker threads (tokenise, pack into pinned memory,
//   bsl_align_se/pe on a GPU, SAM text) -> output in input order (parallel pwrite into reserved byte ranges)
//
// Reads shard across GPUs by batch; every GPU holds a replica of the index; results are merged
// in input order, so the output does not depend on the number of GPUs.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cerrno>
#include <cstring>
#include <ctime>
#include <deque>
#include <map>
#include <memory>
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
        const size_t before = R.cat.size(); R.cat.resize(before + (n - i)); u8 *d = R.cat.data() + before; size_t k = 0;
        for (; i < n; i++) if (!isspace((unsigned char)s[i])) d[k++] = (u8)s[i];
        R.cat.resize(before + k); R.len.back() += (u32)k;
    }
    in.close();
    // sequences with no bases end the reference's loading loop (refbase.cpp:192: `while(LoadNextSeq(...))`)
    for (size_t c = 0; c < R.len.size(); c++) if (R.len[c] == 0) { R.names.resize(c); R.off.resize(c); R.len.resize(c); break; }
    for (u32 l : R.len) R.total += l;
    return true;
}

// ---- read input: every source (plain / gzip'd FASTA or FASTQ, unaligned BAM) is cut by ONE splitter thread per file into
//      text blocks of whole records; the worker threads tokenise the blocks (ReadClass::LoadBatchReads, reads.cpp:42-111).
//      A plain file is memory-mapped and its blocks are views of the mapping; compressed input is inflated by the
//      splitter thread into owned buffers (BAM records are rewritten as FASTQ text there).
struct TextBlock { const char *p = nullptr; size_t n = 0; u32 records = 0; std::shared_ptr<std::vector<char>> keep; };

struct ReadSource {
    int format = -1;                // 0 fasta, 1 fastq, 3 bam (unaligned reads in a BAM file, reads.cpp:85-108)
    int mate = 0;                   // BAM input of a paired run: both -a and -b name the same interleaved file; file #1 takes
                                    // records 0, 2, 4, ... and file #2 records 1, 3, 5, ... (reads.cpp:88,107)
    bool fastq() const { return format != 0; }          // BAM records are handed on as FASTQ text
    const char *format_name() const { return format == 3 ? "BAM" : format == 1 ? "FASTQ" : "FASTA"; }
    // plain file
    const char *map = nullptr; size_t map_len = 0, map_pos = 0;
    // streamed input
    GzReader in; bool streamed = false; std::vector<char> rec; std::vector<char> pending;     // pending: bytes of the block being built
    u32 max_readlen = 480;

    bool open(const std::string &path, u32 max_len) {
        max_readlen = max_len;
        int fd = ::open(path.c_str(), O_RDONLY); if (fd < 0) return false;
        struct stat st; if (fstat(fd, &st) != 0) { ::close(fd); return false; }
        unsigned char magic[4] = {0, 0, 0, 0}; const ssize_t got = pread(fd, magic, 4, 0);
        const bool gz = got >= 2 && magic[0] == 0x1f && magic[1] == 0x8b;
        if (!gz && S_ISREG(st.st_mode) && st.st_size > 0) {
            void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) { madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL); map = (const char *)m; map_len = (size_t)st.st_size; }
        }
        ::close(fd);
        if (map) {
            size_t i = 0; while (i < map_len && isspace((unsigned char)map[i])) i++;
            format = (i < map_len && map[i] == '>') ? 0 : (i < map_len && map[i] == '@') ? 1 : -1;
            return true;
        }
        streamed = true;
        if (!in.open(path)) return false;
        in.fill();
        if (in.len >= 4 && memcmp(in.buf.data(), "BAM\1", 4) == 0) {                 // zlib has already undone the BGZF layer
            format = -1; int32_t l_text = 0, n_ref = 0;
            if (!in.bytes(nullptr, 4) || !in.bytes(&l_text, 4) || l_text < 0 || !in.bytes(nullptr, (size_t)l_text) || !in.bytes(&n_ref, 4) || n_ref < 0) return true;
            for (int32_t k = 0; k < n_ref; k++) { int32_t l_name = 0; if (!in.bytes(&l_name, 4) || l_name < 0 || !in.bytes(nullptr, (size_t)l_name + 4)) return true; }
            format = 3; return true;
        }
        size_t i = 0; while (i < in.len && isspace((unsigned char)in.buf[i])) i++;
        format = (i < in.len && in.buf[i] == '>') ? 0 : (i < in.len && in.buf[i] == '@') ? 1 : -1;
        return true;
    }
    // one BAM alignment record appended to `out` as FASTQ text; out == null skips it
    bool bam_record(std::vector<char> *out) {
        int32_t bs = 0; if (!in.bytes(&bs, 4) || bs < 32) return false;
        rec.resize((size_t)bs); if (!in.bytes(rec.data(), (size_t)bs)) return false;
        if (!out) return true;
        const unsigned char *b = (const unsigned char *)rec.data();
        const u32 l_name = b[8], n_cig = (u32)b[12] | ((u32)b[13] << 8); u32 l_seq; memcpy(&l_seq, b + 16, 4);
        const size_t o_name = 32, o_seq = o_name + l_name + 4ull * n_cig, o_qual = o_seq + (l_seq + 1) / 2;
        if (o_qual + l_seq > (size_t)bs || l_name == 0) return false;
        const size_t nl = strnlen((const char *)b + o_name, l_name);
        const u32 n = std::min<u32>(l_seq, max_readlen);                                  // reads.cpp:93
        const size_t at = out->size(); out->resize(at + nl + 2 * (size_t)n + 6); char *d = out->data() + at;
        *d++ = '@'; memcpy(d, b + o_name, nl); d += nl; *d++ = '\n';
        for (u32 i = 0; i < n; i++) *d++ = "=ACMGRSVTWYHKDBN"[(b[o_seq + i / 2] >> ((i & 1) ? 0 : 4)) & 15];   // bam_nt16_rev_table (reads.cpp:103)
        *d++ = '\n'; *d++ = '+'; *d++ = '\n';
        for (u32 i = 0; i < n; i++) *d++ = (char)(b[o_qual + i] + 33);
        *d++ = '\n';
        return true;
    }
    // The lines of one record: the header is the next non-empty line, then 1 (FASTA) or 3 (FASTQ) more lines, whatever they
    // hold. Returns the end of the record inside [p, e), or null when the input ends first.
    const char *skip_record(const char *p, const char *e, bool at_eof) const {
        const int more = format == 0 ? 1 : 3;
        for (;;) {                                                      // header: skip empty lines (also "\r\n" ones)
            if (p >= e) return nullptr;
            const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
            const char *le = nl ? nl : e; if (!nl && !at_eof) return nullptr;
            size_t n = (size_t)(le - p); if (n && p[n - 1] == '\r') n--;
            p = nl ? nl + 1 : e;
            if (n) break;
        }
        for (int k = 0; k < more; k++) {
            if (p >= e) return nullptr;
            const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
            if (!nl) { if (!at_eof) return nullptr; p = e; continue; }
            p = nl + 1;
        }
        return p;
    }
    // the next `want` records (fewer at the end of the input); false when nothing is left
    bool next(TextBlock &b, u32 want) {
        b = TextBlock();
        if (map) {
            const char *p0 = map + map_pos, *e = map + map_len, *p = p0; u32 k = 0;
            while (k < want) { const char *q = skip_record(p, e, true); if (!q) break; p = q; k++; }
            if (!k) return false;
            b.p = p0; b.n = (size_t)(p - p0); b.records = k; map_pos += b.n;
            return true;
        }
        auto buf = std::make_shared<std::vector<char>>(); buf->swap(pending); u32 k = 0; size_t scanned = 0;
        if (format == 3) {
            buf->reserve((size_t)want * 360);
            while (k < want) {
                if (mate == 2 && !bam_record(nullptr)) break;
                if (!bam_record(buf.get())) break;
                k++;
                if (mate == 1 && !bam_record(nullptr)) break;
            }
            scanned = buf->size();
        } else {
            for (;;) {
                const char *base = buf->data();
                while (k < want) { const char *q = skip_record(base + scanned, base + buf->size(), in.eof && in.pos >= in.len); if (!q) break; scanned = (size_t)(q - base); k++; }
                if (k >= want || (in.eof && in.pos >= in.len)) break;
                if (in.pos >= in.len && !in.fill()) continue;                // eof is set now: one more round takes an unterminated last line
                buf->insert(buf->end(), in.buf.data() + in.pos, in.buf.data() + in.len); in.pos = in.len;
            }
            pending.assign(buf->begin() + (long)scanned, buf->end()); buf->resize(scanned);
        }
        if (!k) return false;
        b.keep = buf; b.p = buf->data(); b.n = scanned; b.records = k;
        return true;
    }
};

// a read as the workers see it: views into a TextBlock (or into the arena of its batch)
struct ReadView { const char *name = nullptr, *seq = nullptr, *qual = nullptr; u32 name_len = 0, len = 0, raw_len = 0; bool too_short = false; };

// tokenise one block (reads.cpp:53-84): name = first token after the '@' / '>' character, the rest of the line is dropped
static void parse_block(const TextBlock &b, bool fastq, u32 max_readlen, const char *fasta_qual, std::vector<ReadView> &out) {
    out.clear(); out.reserve(b.records);
    const char *p = b.p, *e = b.p + b.n;
    auto line = [&](const char *&s, size_t &n) -> bool {
        if (p >= e) return false;
        const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p)); const char *le = nl ? nl : e;
        s = p; n = (size_t)(le - p); if (n && s[n - 1] == '\r') n--; p = nl ? nl + 1 : e; return true;
    };
    for (u32 k = 0; k < b.records; k++) {
        const char *s; size_t n; ReadView r;
        do { if (!line(s, n)) return; } while (n == 0);
        size_t i = 1; while (i < n && isspace((unsigned char)s[i])) i++; size_t j = i; while (j < n && !isspace((unsigned char)s[j])) j++;
        r.name = s + i; r.name_len = (u32)(j - i);
        if (!line(s, n)) return;
        while (n && isspace((unsigned char)s[n - 1])) n--;
        r.seq = s; size_t ls = n, lq;
        if (fastq) { if (!line(s, n) || !line(s, n)) return; while (n && isspace((unsigned char)s[n - 1])) n--; r.qual = s; lq = n; }
        else { r.qual = fasta_qual; lq = std::min<size_t>(ls, 480); }
        if (ls > max_readlen) { ls = max_readlen; if (lq > max_readlen) lq = max_readlen; }
        // a quality string of another length than the sequence is replaced by the default one in TrimLowQual (align.cpp:53); until
        // then only its length matters
        r.len = (u32)ls; r.raw_len = (u32)lq;                       // raw_len carries the quality length until trim_read
        out.push_back(r);
    }
}

// ------------------------------------------------------------------------------------------------ trimming (align.cpp:51-76, 418-435)
// Works on views: trimming only ever shortens a read. `arena` receives rewritten quality strings (non-default -z, or a
// quality string whose length differs from the sequence's).
static void trim_read(ReadView &r, const Options &O, std::deque<std::string> &arena, const char *fasta_qual) {
    r.too_short = false;
    u32 qlen = r.raw_len; r.raw_len = r.len;
    for (size_t a = 0, cut = 0; a < O.adapters.size() && !cut; a++) {                                    // TrimAdapter
        const std::string &ad = O.adapters[a];
        for (u32 pos = O.P.seed_size + O.P.index_interval - 1; (size_t)pos + 4 < r.len && r.len >= 4; pos++) {
            u32 m0 = 0, k = 0;
            for (; k < ad.size() && k < 15 && pos + k < r.len; k++) if ((m0 += (ad[k] != r.seq[pos + k])) > 4) break;
            if (k >= m0 * 5 && k > 3) { r.len = pos; if (qlen > pos) qlen = pos; cut = 1; break; }
        }
    }
    // TrimLowQual
    if (r.len != qlen) { r.qual = fasta_qual; qlen = r.len; }                                             // string(seq.size(), zero_qual + default_qual)
    u8 qual_thres = (u8)(O.zero_qual + O.qual_threshold);
    if (O.zero_qual != '!') {
        arena.emplace_back(r.qual, qlen); std::string &q = arena.back();
        for (char &c : q) c = (char)(c - (O.zero_qual - '!'));
        r.qual = q.data(); qual_thres = (u8)(qual_thres - (O.zero_qual - '!'));
    }
    if (O.qual_threshold == 0) return;
    size_t i = qlen;
    while (i > 0 && !((u8)r.qual[i - 1] > qual_thres)) i--;
    if (i < O.P.seed_size + O.P.index_interval - 1) { r.too_short = true; return; }
    r.len = (u32)i;
}

// ------------------------------------------------------------------------------------------------ SAM text
static inline char rev_char(char c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
                 case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a'; default: return 'N'; }
}
struct RevTable { char t[256]; RevTable() { for (int i = 0; i < 256; i++) t[i] = rev_char((char)i); } };
static const RevTable kRevTab;
static void append_seq(std::string &os, const ReadView &r, bool rev) {
    if (!rev) { os.append(r.seq, r.len); return; }
    const size_t n = os.size(); os.resize(n + r.len); char *d = &os[n];
    for (u32 i = 0; i < r.len; i++) d[i] = kRevTab.t[(u8)r.seq[r.len - 1 - i]];
}
static void append_qual(std::string &os, const ReadView &r, bool rev) {
    if (!rev) { os.append(r.qual, r.len); return; }
    const size_t n = os.size(); os.resize(n + r.len); char *d = &os[n];
    for (u32 i = 0; i < r.len; i++) d[i] = r.qual[r.len - 1 - i];
}
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
    static void name(std::string &os, const ReadView &r) { os.append(r.name, r.name_len); }
    void unaligned(std::string &os, const ReadView &r, int flag) const {
        name(os, r); os.push_back('\t'); append_i(os, flag); os += "\t*\t0\t0\t*\t*\t0\t0\t"; os.append(r.seq, r.len); os.push_back('\t'); os.append(r.qual, r.len); os.push_back('\n');
    }
    // s_OutHit (align.cpp:616-669): n<0 filtered, n==0 unmapped, else hit count
    void single(std::string &os, const ReadView &r, u32 readset, const bsl_hit &h, int n) const {
        int flag = 0x40 * (int)readset;
        if (n <= 0) { if (!O.unmap) return; unaligned(os, r, flag | (n < 0 ? 0x204 : 0x4)); return; }
        bool rev = (h.read_chain ^ (h.chr & 1)) != 0;
        if (n > 1) flag |= 0x100;
        if (rev) flag |= 0x10;
        name(os, r); os.push_back('\t'); append_i(os, flag); os.push_back('\t'); os += R.names[h.chr >> 1]; os.push_back('\t'); append_u(os, h.loc + 1);
        os += "\t255\t"; append_cigar(os, h); os += "\t*\t0\t0\t"; append_seq(os, r, rev); os.push_back('\t'); append_qual(os, r, rev);
        os += "\tNM:i:"; append_i(os, h.nm);
        if (O.outref) xr(os, h);
        os += "\tZS:Z:"; os.push_back("+-"[h.chr & 1]); os.push_back("+-"[h.read_chain]); os.push_back('\n');
    }
    // s_OutHitPair (pairs.cpp:307-416)
    void pair(std::string &os, const ReadView &ra, const ReadView &rb, const bsl_hit &a, const bsl_hit &b, u32 chain, u32 insert, int n) const {
        for (int side = 0; side < 2; side++) {
            const bsl_hit &me = side ? b : a, &mate = side ? a : b; const ReadView &r = side ? rb : ra;
            u32 ch = side ? !chain : chain; bool rev = (ch ^ (me.chr & 1)) != 0;
            int flag = 0x3; if (n > 1) flag |= 0x100; long long ins;
            if (rev) { flag |= 0x10; ins = -(long long)(int)insert; } else { flag |= 0x20; ins = (int)insert; }
            flag |= 0x40 * (side + 1);
            name(os, r); os.push_back('\t'); append_i(os, flag); os.push_back('\t'); os += R.names[me.chr >> 1]; os.push_back('\t'); append_u(os, me.loc + 1);
            os += "\t255\t"; append_cigar(os, me); os += "\t=\t"; append_u(os, mate.loc + 1); os.push_back('\t'); append_i(os, ins); os.push_back('\t');
            append_seq(os, r, rev); os.push_back('\t'); append_qual(os, r, rev); os += "\tNM:i:"; append_i(os, me.nm);
            if (O.outref) xr(os, me);
            os += "\tZS:Z:"; os.push_back("+-"[me.chr & 1]); os.push_back("+-"[ch]); os.push_back('\n');
        }
    }
    // s_OutHitUnpair (pairs.cpp:418-485)
    void unpair(std::string &os, const ReadView &r, int side, u32 chain_a, u32 chain_b, int ma, u32 na, const bsl_hit &ha, int mb, const bsl_hit &hb) const {
        int flag = 1 | (0x40 * (side + 1)); bool rev = (chain_a ^ (ha.chr & 1)) != 0;
        if (ma <= 0) {
            if (ma < 0) flag |= 0x204;
            if (ma == 0) flag |= 0x4;
            if (mb <= 0) { unaligned(os, r, flag | 0x8); return; }
            if (chain_b ^ (hb.chr & 1)) flag |= 0x20;
            name(os, r); os.push_back('\t'); append_i(os, flag); os += "\t*\t0\t0\t*\t"; os += R.names[hb.chr >> 1]; os.push_back('\t'); append_u(os, hb.loc + 1);
            os += "\t0\t"; os.append(r.seq, r.len); os.push_back('\t'); os.append(r.qual, r.len); os.push_back('\n'); return;
        }
        if (ma > 1) flag |= 0x100;
        if (rev) flag |= 0x10;
        if (mb <= 0) flag |= 0x8; else if (chain_b ^ (hb.chr & 1)) flag |= 0x20;
        name(os, r); os.push_back('\t'); append_i(os, flag); os.push_back('\t'); os += R.names[ha.chr >> 1]; os.push_back('\t'); append_u(os, ha.loc + 1);
        os += "\t255\t"; append_cigar(os, ha);
        if (mb <= 0) os += "\t*\t0\t0\t"; else { os.push_back('\t'); os += R.names[hb.chr >> 1]; os.push_back('\t'); append_u(os, hb.loc + 1); os += "\t0\t"; }
        append_seq(os, r, rev); os.push_back('\t'); append_qual(os, r, rev); os += "\tNM:i:"; append_i(os, na);
        if (O.outref) xr(os, ha);
        os += "\tZS:Z:"; os.push_back("+-"[ha.chr & 1]); os.push_back("+-"[chain_a]); os.push_back('\n');
    }
};

// FixPairReadName (pairs.cpp:487-507): both names are cut after the last digit of their common prefix
static void fix_pair_names(ReadView &a, ReadView &b) {
    if (a.name_len == b.name_len && memcmp(a.name, b.name, a.name_len) == 0) return;
    int d = -1; size_t i, n = std::min(a.name_len, b.name_len);
    for (i = 0; i < n; i++) { if (a.name[i] != b.name[i]) break; else if (isdigit((unsigned char)a.name[i])) d = (int)i; }
    if (i > 0) { if (d < 0) d = (int)i - 1; a.name_len = std::min<u32>(a.name_len, (u32)d + 1); b.name_len = std::min<u32>(b.name_len, (u32)d + 1); }
    else { fprintf(stderr, "Error: Paired reads name not match:\n%.*s\n%.*s\n", (int)a.name_len, a.name, (int)b.name_len, b.name); exit(1); }
}

// ------------------------------------------------------------------------------------------------ pipeline
//   splitter thread per input file  ->  queue of text blocks (whole records)
//   worker threads: take block k of every file (ticket k), tokenise + trim + pack the bases into pinned memory,
//                   bsl_align_se / bsl_align_pe on their GPU, SAM text (or BGZF blocks) into their own buffer
//   output: tickets reserve their byte range in input order, the bytes are written by the workers in parallel (pwrite);
//           a pipe / stdout is written in ticket order
struct BlockQueue {
    std::mutex mu; std::condition_variable cv_put, cv_get; std::deque<TextBlock> q; bool done = false; size_t cap = 8;
    void put(TextBlock &&b) { std::unique_lock<std::mutex> g(mu); cv_put.wait(g, [&] { return q.size() < cap; }); q.push_back(std::move(b)); cv_get.notify_one(); }
    void finish() { std::lock_guard<std::mutex> g(mu); done = true; cv_get.notify_all(); }
    bool get(TextBlock &b) { std::unique_lock<std::mutex> g(mu); cv_get.wait(g, [&] { return !q.empty() || done; }); if (q.empty()) return false; b = std::move(q.front()); q.pop_front(); cv_put.notify_one(); return true; }
};

struct OutSink {
    int fd = -1; bool seekable = false;
    std::mutex mu; std::condition_variable cv; u64 next_ticket = 0; u64 offset = 0;
    // A regular file grows in 256 MB windows that are mapped once (MAP_SHARED) and never moved: a ticket reserves its byte range in
    // input order, then the worker copies its bytes into the mapping — in parallel with the other workers (write(2) / pwrite(2)
    // on one file serialise on the inode lock). A file system that cannot map falls back to pwrite.
    static constexpr u64 kWin = 256ull << 20;
    std::vector<char *> wins; bool can_map = true;
    static bool write_all(int fd, const char *p, size_t n) { while (n) { ssize_t k = ::write(fd, p, n); if (k < 0) { if (errno == EINTR) continue; return false; } p += k; n -= (size_t)k; } return true; }
    static bool pwrite_all(int fd, const char *p, size_t n, u64 off) { while (n) { ssize_t k = ::pwrite(fd, p, n, (off_t)off); if (k < 0) { if (errno == EINTR) continue; return false; } p += k; n -= (size_t)k; off += (u64)k; } return true; }
    bool map_upto(u64 end) {                                             // caller holds mu
        while (can_map && (u64)wins.size() * kWin < end) {
            const u64 at = (u64)wins.size() * kWin;
            if (ftruncate(fd, (off_t)(at + kWin)) != 0) { can_map = false; break; }
            void *m = mmap(nullptr, kWin, PROT_READ | PROT_WRITE, MAP_SHARED, fd, (off_t)at);
            if (m == MAP_FAILED) { can_map = false; break; }
            wins.push_back((char *)m);
#ifdef MADV_POPULATE_WRITE
            if (wins.size() > 1) std::thread([m]() { madvise(m, kWin, MADV_POPULATE_WRITE); }).detach();   // a long output: fault the window in ahead of the workers' copies
#endif
        }
        return can_map;
    }
    void copy_in(const char *p, size_t n, u64 off) { while (n) { const u64 w = off / kWin, o = off % kWin; const size_t k = (size_t)std::min<u64>(n, kWin - o); memcpy(wins[w] + o, p, k); p += k; n -= k; off += k; } }
    bool put(const char *p, size_t n, u64 off, bool mapped) { if (mapped) { copy_in(p, n, off); return true; } return pwrite_all(fd, p, n, off); }
    bool head(const std::string &s) {
        if (s.empty()) return true;
        if (!seekable) return write_all(fd, s.data(), s.size());
        std::unique_lock<std::mutex> g(mu); const u64 off = offset; offset += s.size(); const bool mapped = map_upto(offset); g.unlock();
        return put(s.data(), s.size(), off, mapped);
    }
    // the bytes of ticket t; blocks until every earlier ticket has reserved its range (regular file) or has been written (pipe)
    bool submit(u64 t, const char *p, size_t n) {
        std::unique_lock<std::mutex> g(mu);
        cv.wait(g, [&] { return next_ticket == t; });
        bool ok = true;
        if (seekable) { const u64 off = offset; offset += n; const bool mapped = map_upto(offset); next_ticket++; g.unlock(); cv.notify_all(); ok = put(p, n, off, mapped); }
        else { ok = write_all(fd, p, n); next_ticket++; g.unlock(); cv.notify_all(); }
        return ok;
    }
    void skip(u64 t) { std::unique_lock<std::mutex> g(mu); cv.wait(g, [&] { return next_ticket == t; }); next_ticket++; g.unlock(); cv.notify_all(); }
    void finish() { for (char *w : wins) munmap(w, kWin); if (!wins.empty()) { if (ftruncate(fd, (off_t)offset) != 0) perror("ftruncate"); } wins.clear(); }
};

struct Counters { u64 al = 0, un = 0, mu = 0, pal = 0, pun = 0, pmu = 0, aal = 0, aun = 0, amu = 0, bal = 0, bun = 0, bmu = 0; };

// Earlier reads a batch must be preceded by so that the aligner state the reference carries from read to read
// (xseed_start_offset, xseed_array: align.cpp:476-480, 79-150; SURVEY trap 3) is what a -p 1 run of the reference would
// have: per mate, the last unfiltered read with a non-empty start-offset range, and every unfiltered read longer than all
// unfiltered reads after it (the reads whose seed hashes are still visible beyond the end of a shorter read).
struct CtxPair { std::string a, b; u32 raw_a = 0, raw_b = 0; bool short_a = false, short_b = false; };
typedef std::vector<CtxPair> Tail;

struct Pipeline {
    const Options &O; const Reference &R; const TextRule &T; bool pe;
    std::vector<bsl_ctx *> ctx;
    ReadSource fa, fb; BlockQueue qa, qb;
    std::mutex in_mu; u64 next_ticket = 0; u32 next_index; u64 reads_left;
    OutSink sink;
    bool bam_native = false; bam::Refs refs;          // -o x.bam without an external samtools
    Counters total; std::mutex total_mu;
    std::atomic<int> failed{0};
    u32 batch_reads; std::string fasta_qual;
    // tails: tail[k] = context for batch k+1, published by the worker of batch k
    std::mutex tail_mu; std::condition_variable tail_cv; std::map<u64, std::shared_ptr<Tail>> tails;

    Pipeline(const Options &o, const Reference &r, const TextRule &t) : O(o), R(r), T(t), pe(!o.b.empty()), next_index(o.read_start - 1) {
        const char *e = getenv("BASAL_BATCH"); batch_reads = e ? (u32)std::max(1l, atol(e)) : (pe ? 131072u : 262144u);
        reads_left = O.read_end > O.read_start - 1 ? (u64)O.read_end - (O.read_start - 1) : 0;
        fasta_qual.assign(512, (char)(O.zero_qual + O.default_qual));
    }

    void splitter(ReadSource &src, BlockQueue &q) {
        u64 skip = O.read_start - 1, left = reads_left;                        // InitIndex (reads.cpp:13-40) / -E
        TextBlock b;
        while (skip) { const u32 k = (u32)std::min<u64>(skip, 1u << 20); if (!src.next(b, k)) { q.finish(); return; } skip -= b.records; if (b.records < k) { q.finish(); return; } }
        while (left && !failed) { const u32 k = (u32)std::min<u64>(left, batch_reads); if (!src.next(b, k)) break; left -= b.records; q.put(std::move(b)); }
        q.finish();
    }

    struct Batch {
        u64 ticket = 0; u32 first_index = 0, n = 0; TextBlock ba, bb;
        std::vector<ReadView> a, b; std::deque<std::string> arena; std::string text, blk;
    };
    // per-worker staging in pinned memory (bsl_host_alloc), grown on demand and reused from batch to batch
    struct Staging {
        u8 *bases[2] = {nullptr, nullptr}; size_t cap_bases[2] = {0, 0};
        bsl_hit *hits[2] = {nullptr, nullptr}; bsl_pair *pairs = nullptr; size_t cap_recs = 0;
        std::vector<u64> off[2]; std::vector<u16> raw[2];
        bool need(int m, size_t bytes) { if (bytes <= cap_bases[m]) return true; bsl_host_free(bases[m]); cap_bases[m] = bytes + bytes / 4 + 4096; bases[m] = (u8 *)bsl_host_alloc(cap_bases[m]); return bases[m] != nullptr; }
        bool need_recs(size_t n, bool pe) { if (n <= cap_recs) return true; for (int m = 0; m < 2; m++) bsl_host_free(hits[m]); bsl_host_free(pairs); cap_recs = n + n / 4 + 64;
            hits[0] = (bsl_hit *)bsl_host_alloc(cap_recs * sizeof(bsl_hit)); hits[1] = pe ? (bsl_hit *)bsl_host_alloc(cap_recs * sizeof(bsl_hit)) : nullptr; pairs = pe ? (bsl_pair *)bsl_host_alloc(cap_recs * sizeof(bsl_pair)) : nullptr;
            return hits[0] && (!pe || (hits[1] && pairs)); }
        ~Staging() { for (int m = 0; m < 2; m++) { bsl_host_free(bases[m]); bsl_host_free(hits[m]); } bsl_host_free(pairs); }
    };

    // a pool of staging sets (one per worker + 2), allocated by a background thread while the index is built and the first
    // batches are parsed: pinned memory is expensive to allocate
    std::vector<std::unique_ptr<Staging>> pool; std::vector<Staging *> pool_free; std::mutex pool_mu; std::condition_variable pool_cv;
    void make_pool(size_t n_sets) {
        const size_t guess = (size_t)batch_reads * 168 + 4096;                    // 150-bp reads fit without a second allocation
        for (size_t k = 0; k < n_sets; k++) {
            std::unique_ptr<Staging> s(new Staging());
            s->need(0, guess); if (pe) s->need(1, guess); s->need_recs(batch_reads + 64, pe);
            std::lock_guard<std::mutex> g(pool_mu); pool_free.push_back(s.get()); pool.push_back(std::move(s)); pool_cv.notify_one();
        }
    }
    Staging *acquire() { std::unique_lock<std::mutex> g(pool_mu); pool_cv.wait(g, [&] { return !pool_free.empty(); }); Staging *s = pool_free.back(); pool_free.pop_back(); return s; }
    void release(Staging *s) { std::lock_guard<std::mutex> g(pool_mu); pool_free.push_back(s); pool_cv.notify_one(); }
    // phase clocks (seconds summed over the workers), printed with $BASAL_TIMING
    std::atomic<long long> ns_parse{0}, ns_wait_stage{0}, ns_pack{0}, ns_gpu{0}, ns_format{0}, ns_out{0};
    static long long now_ns() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (long long)ts.tv_sec * 1000000000ll + ts.tv_nsec; }

    bool load(Batch &B) {
        std::lock_guard<std::mutex> g(in_mu);
        if (!qa.get(B.ba)) return false;
        if (pe && !qb.get(B.bb)) return false;
        B.n = pe ? std::min(B.ba.records, B.bb.records) : B.ba.records;
        if (!B.n) return false;
        B.ticket = next_ticket++; B.first_index = next_index; next_index += B.n;
        return true;
    }

    static int mate_count(const bsl_hit &h) { return h.status == BSL_ST_FILTERED ? -1 : (h.status == BSL_ST_UNMAPPED ? 0 : (int)h.n_hits); }

    // ---- carried aligner state across batches (SURVEY trap 3)
    bool defining(u32 len) const { return len + 1 >= O.P.index_interval && (len + 1 - O.P.index_interval) % O.P.seed_size != 0; }
    bool filtered(const char *s, u32 len, bool too_short) const {                 // FilterReads (align.cpp:557-560)
        if (too_short || len == 0 || len < O.P.min_read_size) return true;
        u32 ns = 0; for (u32 i = 0; i < len; i++) { const char c = (char)(s[i] & 0xDF); ns += !(c == 'A' || c == 'C' || c == 'G' || c == 'T'); }
        return ns > O.P.max_ns;
    }
    std::shared_ptr<Tail> wait_tail(u64 ticket) {                                // the context in front of batch `ticket`
        if (ticket == 0) return std::make_shared<Tail>();
        std::unique_lock<std::mutex> g(tail_mu);
        tail_cv.wait(g, [&] { return tails.count(ticket - 1) != 0 || failed; });
        auto it = tails.find(ticket - 1); return it == tails.end() ? std::make_shared<Tail>() : it->second;
    }
    void publish_tail(Batch &B) {
        // walk back through this batch, then through the context it was given, until both mates have their last defining
        // read and nothing longer can follow
        std::shared_ptr<Tail> prev;                                             // fetched only if this batch does not settle it
        u32 maxlen[2] = {0, 0}; bool have_def[2] = {false, !pe};
        u32 top = 0; for (u32 i = 0; i < B.n; i++) { top = std::max(top, B.a[i].len); if (pe) top = std::max(top, B.b[i].len); }
        std::vector<CtxPair> picked;                                            // newest first
        auto consider = [&](const char *sa, u32 la, u32 ra, bool ta, const char *sb, u32 lb, u32 rb, bool tb) {
            bool want = false;
            for (int m = 0; m < (pe ? 2 : 1); m++) {
                const char *s = m ? sb : sa; const u32 l = m ? lb : la; const bool ts = m ? tb : ta;
                const bool cand = l > maxlen[m] || (!have_def[m] && defining(l));
                if (!cand || filtered(s, l, ts)) continue;
                if (l > maxlen[m]) maxlen[m] = l;
                if (defining(l)) have_def[m] = true;
                want = true;
            }
            if (want) { CtxPair c; c.a.assign(sa, la); c.raw_a = ra; c.short_a = ta; if (pe) { c.b.assign(sb, lb); c.raw_b = rb; c.short_b = tb; } picked.push_back(std::move(c)); }
        };
        auto settled = [&](u32 cap) { for (int m = 0; m < (pe ? 2 : 1); m++) if (!have_def[m] || maxlen[m] < cap) return false; return true; };
        for (u32 i = B.n; i-- > 0 && !settled(480);) {
            const ReadView &a = B.a[i]; const ReadView *b = pe ? &B.b[i] : nullptr;
            consider(a.seq, a.too_short ? 0 : a.len, a.raw_len, a.too_short, b ? b->seq : nullptr, b ? (b->too_short ? 0 : b->len) : 0, b ? b->raw_len : 0, b ? b->too_short : false);
            if (i == 0 || settled(480)) break;
            if (settled(top) && i + 1 < B.n) { /* nothing in this batch can add to the skyline any more; older context may */ break; }
        }
        if (!settled(480)) {
            prev = wait_tail(B.ticket);
            for (size_t i = prev->size(); i-- > 0 && !settled(480);) { const CtxPair &c = (*prev)[i]; consider(c.a.data(), (u32)c.a.size(), c.raw_a, c.short_a, c.b.data(), (u32)c.b.size(), c.raw_b, c.short_b); }
        }
        auto t = std::make_shared<Tail>(picked.rbegin(), picked.rend());
        { std::lock_guard<std::mutex> g(tail_mu); tails[B.ticket] = t; if (B.ticket >= 2) tails.erase(B.ticket - 2); }
        tail_cv.notify_all();
    }

    bool process(Batch &B, bsl_ctx *c, Counters &cn) {
        Staging *Sp = nullptr;
        const bool ok = process_with(B, Sp, c, cn);
        if (Sp) release(Sp);
        return ok;
    }
    bool process_with(Batch &B, Staging *&Sp, bsl_ctx *c, Counters &cn) {
        const u32 n = B.n; Formatter F(O, R, T);
        long long t0 = now_ns(), t1;
        B.arena.clear();
        parse_block(B.ba, fa.fastq(), O.max_readlen, fasta_qual.data(), B.a);
        if (pe) parse_block(B.bb, fb.fastq(), O.max_readlen, fasta_qual.data(), B.b);
        if (B.a.size() < n || (pe && B.b.size() < n)) { fprintf(stderr, "\ninternal error: block holds fewer records than counted\n"); return false; }
        bool need_ctx = false;
        for (int m = 0; m < (pe ? 2 : 1); m++) for (ReadView &r : (m ? B.b : B.a)) { trim_read(r, O, B.arena, fasta_qual.data()); if (!r.too_short && r.len >= O.P.min_read_size && !defining(r.len)) need_ctx = true; }
        if (pe) for (u32 i = 0; i < n; i++) fix_pair_names(B.a[i], B.b[i]);
        // ---- context reads in front of the batch (only when a read of this batch can inherit state at all)
        std::shared_ptr<Tail> tin; u32 nctx = 0;
        if (need_ctx) { tin = wait_tail(B.ticket); nctx = (u32)tin->size(); }
        publish_tail(B);
        t1 = now_ns(); ns_parse += t1 - t0; t0 = t1;
        Sp = acquire(); Staging &S = *Sp;
        t1 = now_ns(); ns_wait_stage += t1 - t0; t0 = t1;
        // ---- pack: bases of (context +) batch, back to back, into pinned memory
        const u32 nt = nctx + n;
        for (int m = 0; m < (pe ? 2 : 1); m++) {
            std::vector<ReadView> &v = m ? B.b : B.a;
            size_t tot = 0; for (u32 k = 0; k < nctx; k++) tot += m ? (*tin)[k].b.size() : (*tin)[k].a.size();
            for (u32 i = 0; i < n; i++) tot += v[i].too_short ? 0 : v[i].len;
            if (!S.need(m, tot + 64)) { fprintf(stderr, "\npinned host allocation failed\n"); return false; }
            S.off[m].resize((size_t)nt + 1); S.raw[m].resize(nt); size_t p = 0; u8 *dst = S.bases[m];
            for (u32 k = 0; k < nctx; k++) { const CtxPair &cp = (*tin)[k]; const std::string &s = m ? cp.b : cp.a; S.off[m][k] = p; memcpy(dst + p, s.data(), s.size()); p += s.size(); S.raw[m][k] = (u16)(m ? cp.raw_b : cp.raw_a); }
            for (u32 i = 0; i < n; i++) { S.off[m][nctx + i] = p; if (!v[i].too_short) { memcpy(dst + p, v[i].seq, v[i].len); p += v[i].len; } S.raw[m][nctx + i] = (u16)v[i].raw_len; }
            S.off[m][nt] = p;
        }
        if (!S.need_recs(nt, pe)) { fprintf(stderr, "\npinned host allocation failed\n"); return false; }
        bsl_batch qa; memset(&qa, 0, sizeof qa); qa.n = nt; qa.readset = pe ? 1 : 0; qa.bases = S.bases[0]; qa.offsets = S.off[0].data(); qa.first_index = B.first_index - nctx; qa.raw_len = S.raw[0].data(); qa.n_context = nctx;
        const bool all = O.P.report_repeat_hits == 2;
        std::vector<bsl_hit> alla, allb; u64 n_all = 0; u64 all_cap = all ? std::max<u64>((u64)nt * 8, 1u << 20) : 0;
        t1 = now_ns(); ns_pack += t1 - t0; t0 = t1;
        int rc;
        for (;;) {
            if (all) { alla.resize(all_cap); if (pe) allb.resize(all_cap); }
            if (!pe) rc = bsl_align_se(c, &qa, S.hits[0], all ? alla.data() : nullptr, all_cap, &n_all);
            else { bsl_batch qb = qa; qb.readset = 2; qb.bases = S.bases[1]; qb.offsets = S.off[1].data(); qb.raw_len = S.raw[1].data();
                rc = bsl_align_pe(c, &qa, &qb, S.hits[0], S.hits[1], S.pairs, all ? alla.data() : nullptr, all ? allb.data() : nullptr, all_cap, &n_all); }
            if (rc == 0 && all && n_all > all_cap) { all_cap = n_all + 16; continue; }      // grow the all-hits buffer and redo the batch
            break;
        }
        if (rc != 0) { fprintf(stderr, "GPU alignment failed (%d): %s\n", rc, bsl_last_error(c)); return false; }
        t1 = now_ns(); ns_gpu += t1 - t0; t0 = t1;
        const bsl_hit *ha = S.hits[0] + nctx, *hb = pe ? S.hits[1] + nctx : nullptr; const bsl_pair *hp = pe ? S.pairs + nctx : nullptr;
        std::string &os = B.text; os.clear(); os.reserve((size_t)n * (pe ? 900 : 420));
        for (u32 i = 0; i < n; i++) {
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
                if (O.P.report_repeat_hits == 1) { cn.aal++; F.unpair(os, B.a[i], 0, ca, cb, ma, a.nm, a, mb1, b); }
                else if (O.P.report_repeat_hits == 2) { cn.aal++; for (int k = 0; k < ma; k++) { const bsl_hit &x = alla[a.all_first + k]; F.unpair(os, B.a[i], 0, x.read_chain, cb, ma, a.nm, x, mb1, b); } }   // pairs.cpp:270-274
                else if (O.unmap) F.unpair(os, B.a[i], 0, 0, cb, 0, 0, a, mb1, b); }
            if (mb <= 0) { if (O.unmap) F.unpair(os, B.b[i], 1, 0, ca, mb, 0, b, ma1, a); }
            else if (mb == 1) { cn.bal++; cn.bun++; F.unpair(os, B.b[i], 1, cb, ca, 1, b.nm, b, ma1, a); }
            else { cn.bmu++;
                if (O.P.report_repeat_hits == 1) { cn.bal++; F.unpair(os, B.b[i], 1, cb, ca, mb, b.nm, b, ma1, a); }
                else if (O.P.report_repeat_hits == 2) { cn.bal++; for (int k = 0; k < mb; k++) { const bsl_hit &x = allb[b.all_first + k]; F.unpair(os, B.b[i], 1, x.read_chain, cb, mb, b.nm, x, ma1, a); } }   // pairs.cpp:293-297 (passes cb, not ca)
                else if (O.unmap) F.unpair(os, B.b[i], 1, 0, ca, 0, 0, b, ma1, a); }
        }
        ns_format += now_ns() - t0;
        return true;
    }

    void worker(int wid) {
        bsl_ctx *c = ctx[wid % ctx.size()]; Counters cn; Batch B;
        while (!failed && load(B)) {
            bool ok = process(B, c, cn);
            const long long tw = now_ns();
            const std::string *out = &B.text;
            if (ok && bam_native) { B.blk.clear(); if (!bam::text_to_blocks(B.text, refs, B.blk)) { fprintf(stderr, "\ninternal error: malformed SAM record in BAM conversion\n"); ok = false; } out = &B.blk; }
            if (!ok) { failed = 1; tail_cv.notify_all(); sink.skip(B.ticket); break; }
            if (!sink.submit(B.ticket, out->data(), out->size())) { fprintf(stderr, "\nwrite failed: %s\n", strerror(errno)); failed = 1; break; }
            ns_out += now_ns() - tw;
            if (O.verbose >= 2) fprintf(stderr, "[BASAL @%s] batch %llu finished. %ld secs passed\n", now_str(), (unsigned long long)B.ticket + 1, secs_passed());
        }
        if (failed) { qa.finish(); qb.finish(); TextBlock d; while (qa.get(d)) {} while (qb.get(d)) {} }      // unblock the splitters
        std::lock_guard<std::mutex> g(total_mu);
        total.al += cn.al; total.un += cn.un; total.mu += cn.mu; total.pal += cn.pal; total.pun += cn.pun; total.pmu += cn.pmu;
        total.aal += cn.aal; total.aun += cn.aun; total.amu += cn.amu; total.bal += cn.bal; total.bun += cn.bun; total.bmu += cn.bmu;
    }
};

static void check_input(const std::string &path, const char *msg) {
    FILE *f = fopen(path.c_str(), "rb"); if (!f) { fprintf(stderr, "\n%s%s\n", msg, path.c_str()); exit(1); } fclose(f);
}
static double wall() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }

int main(int argc, char **argv) {
    if (argc == 1) usage();
    g_t0 = time(nullptr);
    const double t_start = wall();
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
    const double t_ref = wall();
    const bool pe = !O.b.empty();

    // ---- $BASAL_PARSE_ONLY: run the read loader alone (no GPU) and print what it parsed as FASTQ, mates interleaved:
    //      the CPU test hook of the host loader (FASTA / FASTQ / gz / BAM input, -L, -B / -E, block boundaries)
    if (getenv("BASAL_PARSE_ONLY")) {
        Pipeline P(O, R, T);
        if (!P.fa.open(O.a, O.max_readlen) || P.fa.format < 0 || (pe && (!P.fb.open(O.b, O.max_readlen) || P.fb.format != P.fa.format))) { fprintf(stderr, "\t(format: unknown)\nUnknown input format.\n"); return 1; }
        if (pe) { P.fa.mate = 1; P.fb.mate = 2; }
        fprintf(stderr, "format: %s\n", P.fa.format_name());
        const bool count_only = strcmp(getenv("BASAL_PARSE_ONLY"), "count") == 0;      // loader throughput without the printing
        std::thread ta([&] { P.splitter(P.fa, P.qa); }), tb; if (pe) tb = std::thread([&] { P.splitter(P.fb, P.qb); });
        Pipeline::Batch B; std::string o; u64 n_rec = 0, n_base = 0;
        while (P.load(B)) {
            parse_block(B.ba, P.fa.fastq(), O.max_readlen, P.fasta_qual.data(), B.a); if (pe) parse_block(B.bb, P.fb.fastq(), O.max_readlen, P.fasta_qual.data(), B.b);
            for (u32 i = 0; i < B.n; i++) for (int m = 0; m < (pe ? 2 : 1); m++) {
                const ReadView &r = m ? B.b[i] : B.a[i]; n_rec++; n_base += r.len;
                if (count_only) continue;
                const u32 ql = std::min(r.raw_len, r.len);                                  // raw_len = quality length before trim_read
                o.push_back('@'); o.append(r.name, r.name_len); o.push_back('\n'); o.append(r.seq, r.len); o += "\n+\n"; o.append(r.qual, ql); o.push_back('\n');
            }
            if (o.size() > (1u << 20)) { fwrite(o.data(), 1, o.size(), stdout); o.clear(); }
        }
        ta.join(); if (pe) tb.join();
        fwrite(o.data(), 1, o.size(), stdout);
        fprintf(stderr, "parsed %llu reads, %llu bases\n", (unsigned long long)n_rec, (unsigned long long)n_base);
        return 0;
    }

    // ---- GPUs: every visible device holds a replica of the index
    int ngpu = 1; { const char *e = getenv("BASAL_GPUS"); if (e) ngpu = std::max(1, atoi(e)); else { const char *v = getenv("BASAL_ALL_GPUS"); if (v && atoi(v)) ngpu = 64; } }
    Pipeline P(O, R, T);
    // a short input is cut into more, smaller batches so that every worker gets some (the default suits long runs)
    if (!getenv("BASAL_BATCH")) {
        struct stat st; if (stat(O.a.c_str(), &st) == 0 && S_ISREG(st.st_mode)) {
            const u64 est = (u64)st.st_size / 330 + 1;                           // ~reads in a plain 150-bp FASTQ file; only a heuristic
            const u64 want = est / (4ull * std::max(O.procs, 3)) + 1;
            P.batch_reads = (u32)std::min<u64>(P.batch_reads, std::max<u64>(want, 16384));
        }
    }
    std::thread pool_thread;
    {
        std::vector<std::thread> th; std::vector<int> rcs; std::vector<bsl_ctx *> cs;
        for (int g = 0; g < ngpu; g++) {
            bsl_ctx *c = nullptr; int rc = bsl_ctx_create(&c, g, &O.P);
            if (rc != 0) { if (g == 0) { fprintf(stderr, "cannot open GPU 0 (%d): %s\n", rc, bsl_last_error(nullptr)); exit(1); } break; }
            cs.push_back(c);
        }
        rcs.assign(cs.size(), 0);
        pool_thread = std::thread([&P, &O, n = cs.size()]() { P.make_pool((size_t)std::max<int>(std::max(O.procs, 1), (int)n * 4) + 2); });   // pinned staging (one set per worker) while the GPUs build their index
        for (size_t g = 0; g < cs.size(); g++) th.emplace_back([&, g]() { rcs[g] = bsl_index_build(cs[g], R.cat.data(), R.off.data(), R.len.data(), (u32)R.len.size()); });
        for (auto &t : th) t.join();
        for (size_t g = 0; g < cs.size(); g++) if (rcs[g] != 0) { fprintf(stderr, "index build failed on GPU %zu (%d): %s\n", g, rcs[g], bsl_last_error(cs[g])); exit(1); }
        P.ctx = cs;
    }
    if (O.verbose >= 1) fprintf(stderr, "[BASAL @%s] create seed table. %ld secs passed\n", now_str(), secs_passed());
    const double t_index = wall();

    // ---- RunProcess (main.cpp:409-614)
    if (O.o.size() > 4) { if (O.o.compare(O.o.size() - 4, 4, ".sam") == 0) O.out_sam = 1; else if (O.o.compare(O.o.size() - 4, 4, ".bam") == 0) O.out_sam = 2; }
    if (O.verbose >= 2) {
        if (O.P.max_snp_num < 100) fprintf(stderr, "\tmax number of mismatches: %u", O.P.max_snp_num); else fprintf(stderr, "\tmax number of mismatches: read_length * %u%% ", O.P.max_snp_num - 100);
        fprintf(stderr, " \tmax gap size: %u \tkmer cut-off ratio: %g \tmax multi-hits: %u\n", O.P.gap, (double)O.P.max_kmer_ratio, O.P.max_num_hits);
        fprintf(stderr, "\tquality cutoff: %d \tbase quality char: '%c' \tmax Ns: %u\n", O.qual_threshold, O.zero_qual, O.P.max_ns);
        fprintf(stderr, "\twildcard mapping approach \tseed size: %u \tindex interval: %u\n", O.P.seed_size, O.P.index_interval);
    }
    const int nw = std::max<int>(std::max(O.procs, 1), (int)P.ctx.size() * 4);                 // at least four calls in flight per GPU context (six lanes each)
    if (O.verbose >= 1) fprintf(stderr, "[BASAL @%s] %s alignment(%zu GPU(s), %d host threads),\n", now_str(), pe ? "Pair-end" : "Single-end", P.ctx.size(), nw);
    check_input(O.a, pe ? "failed to open read file #1 (check -a option): " : "failed to open read file (check -a option): ");
    if (!P.fa.open(O.a, O.max_readlen) || P.fa.format < 0) { fprintf(stderr, "\t(format: unknown)\nUnknown input format.\n"); exit(1); }
    if (pe) P.fa.mate = 1;
    if (O.verbose >= 1) fprintf(stderr, "\tInput read file%s: %s \t(format: %s)\n", pe ? " #1" : "", O.a.c_str(), P.fa.format_name());
    if (pe) {
        check_input(O.b, "failed to open read file #2 (check -b option): ");
        if (!P.fb.open(O.b, O.max_readlen) || P.fb.format < 0) { fprintf(stderr, "\t(format: unknown)\nUnknown input format.\n"); exit(1); }
        if (P.fb.format != P.fa.format) { fprintf(stderr, "Input read file #1 and #2 should be in same format.\n"); exit(1); }
        P.fb.mate = 2;
        if (O.verbose >= 1) fprintf(stderr, "\tInput read file #2: %s \t(format: %s)\n", O.b.c_str(), P.fb.format_name());
    }
    FILE *piped = nullptr;
    if (O.to_stdout) { P.sink.fd = 1; if (O.verbose >= 1) fprintf(stderr, "\tOutput: STDOUT\t (format: SAM)\n"); }
    else {
        if (O.verbose >= 1 || pe) fprintf(stderr, "\tOutput file: %s\t (format: SAM%s)\n", O.o.c_str(), O.out_sam == 2 ? ", automatically convert to BAM" : "");
        const int fd = ::open(O.o.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
        if (fd < 0) { fprintf(stderr, "\nfailed to open output file (check -o option): %s\n", O.o.c_str()); exit(1); }
        if (O.out_sam == 2 && getenv("BASAL_SAMTOOLS")) { ::close(fd); std::string cmd = "samtools view -bS - >" + O.o; piped = popen(cmd.c_str(), "w"); if (piped) P.sink.fd = fileno(piped); }   // main.cpp:505
        if (!piped) { P.sink.fd = piped ? P.sink.fd : (O.out_sam == 2 && getenv("BASAL_SAMTOOLS") ? ::open(O.o.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644) : fd);
            struct stat st; P.sink.seekable = fstat(P.sink.fd, &st) == 0 && S_ISREG(st.st_mode); }
        if (O.out_sam == 2 && !piped) { P.bam_native = true; for (size_t i = 0; i < R.names.size(); i++) P.refs.add(R.names[i], R.len[i]); }
    }
    {
        std::string h;
        if (O.header) {                                                                                  // main.cpp:516-526
            h = "@HD\tVN:1.0\n";
            for (size_t i = 0; i < R.names.size(); i++) { h += "@SQ\tSN:" + R.names[i] + "\tLN:"; append_u(h, R.len[i]); h.push_back('\n'); }
            h += std::string("@PG\tID:BASAL\tVN:") + kVersion + "\tCL:\"" + O.cmdline + "\"\n";
        }
        if (P.bam_native) { const std::string raw = bam::header_bytes(h, P.refs); std::string blk; bam::bgzf_append(raw.data(), raw.size(), blk); P.sink.head(blk); }
        else P.sink.head(h);
    }
    const double t_map0 = wall();
    {
        P.qa.cap = P.qb.cap = (size_t)nw + 2;
        std::thread ta([&] { P.splitter(P.fa, P.qa); }), tb; if (pe) tb = std::thread([&] { P.splitter(P.fb, P.qb); });
        std::vector<std::thread> th; for (int w = 0; w < nw; w++) th.emplace_back([&, w]() { P.worker(w); });
        for (auto &t : th) t.join();
        P.qa.finish(); P.qb.finish(); { TextBlock d; while (P.qa.get(d)) {} while (P.qb.get(d)) {} }
        ta.join(); if (pe) tb.join();
    }
    pool_thread.join();
    if (P.bam_native) { std::string e; bam::bgzf_eof(e); P.sink.head(e); }
    P.sink.finish();
    if (piped) pclose(piped); else if (P.sink.fd != 1) ::close(P.sink.fd);
    const double t_map1 = wall();
    for (bsl_ctx *c : P.ctx) bsl_ctx_destroy(c);
    if (P.failed) return 2;
    const double tot = (double)(P.next_index - (O.read_start - 1));
    if (getenv("BASAL_TIMING"))                                          // machine-readable phase clocks (bench.py: cli_e2e)
        fprintf(stderr, "[timing] load_ref_s=%.4f index_s=%.4f map_s=%.4f total_s=%.4f reads=%.0f gpus=%zu threads=%d batch=%u parse_s=%.3f stage_wait_s=%.3f pack_s=%.3f gpu_s=%.3f format_s=%.3f out_s=%.3f\n",
                t_ref - t_start, t_index - t_ref, t_map1 - t_map0, wall() - t_start, tot * (pe ? 2 : 1), P.ctx.size(), nw, P.batch_reads,
                1e-9 * P.ns_parse, 1e-9 * P.ns_wait_stage, 1e-9 * P.ns_pack, 1e-9 * P.ns_gpu, 1e-9 * P.ns_format, 1e-9 * P.ns_out);
    if (O.verbose >= 1) {                                                                                // main.cpp:536-552, 606-612
        const Counters &c = P.total; const char *sup = O.P.report_repeat_hits == 0 ? "suppressed " : "";
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
