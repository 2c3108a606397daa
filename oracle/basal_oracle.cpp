// basal_oracle.cpp — CPU restatement of BASAL v1.8.1's read-mapping path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (basal_b200/, include/, the
// `basal` CLI) may include, link or execute this file; only tests/, bench.py's
// cpu_baseline / --impl reference legs and __graft_entry__.smoke() use it, as the
// checker.  It is deliberately written over one-byte-per-base code arrays (not the
// packed words the reference and the CUDA path use) so that it is an independent
// statement of the *rules*, cf. SURVEY.md Appendix A/D.
//
// Parity pin: tests/test_oracle_golden.py diffs this program's SAM against the
// unmodified reference binary (oracle/_ref/basal, built by oracle/Makefile.ref)
// and against the committed fixtures in tests/golden/ (made by
// tests/golden/make_golden.py from that binary); tools/fuzz_oracle.py does the same
// on random flag combinations.
//
// Each function cites the reference file:line (under /root/reference) it follows.
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "../include/basal_gpu.h"   // record layouts only (bsl_params/bsl_batch/bsl_hit/bsl_pair)

namespace orc {

typedef uint32_t u32; typedef uint64_t u64; typedef uint8_t u8;
static const u32 MAXSNPS = 15;      // param.h:18
static const u32 MARGIN  = 400;     // refbase.h:16 REF_MARGIN (words)

// ---------------------------------------------------------------- rule (param.cpp:163-263)
struct Rule {
    u8 code[256], rcode[256], reg[256], conv[256], rconv[256];
    char letter[4];          // useful_nt: code -> letter
    bool single;             // exactly one convert-to base and it is not '-'
    char from; std::string to;
    std::string err;
};

static int base_idx(int c) { switch (toupper(c)) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; } return -1; }

static bool make_rule(const std::string &txt, Rule &r) {
    static const char NT[5] = {'A','C','G','T','-'}, CNT[5] = {'T','G','C','A','-'};
    if (txt.size() < 2 || txt[1] != ':') { r.err = "invalid -M, ref base(one letter in A/C/G/T) should be assigned first before :"; return false; }
    r.from = (char)toupper(txt[0]);
    if (base_idx(r.from) < 0) { r.err = std::string("invalid -M, ref base ") + txt[0] + " not in A/C/G/T"; return false; }
    r.to.clear();
    for (size_t i = 2; i < txt.size(); i++) {
        char t = (char)toupper(txt[i]);
        if (t == r.from) { r.err = "invalid -M, read base equal to ref base"; return false; }
        if (!memchr(NT, t, 5)) { r.err = "invalid -M, read base not in A/C/G/T/-"; return false; }
        if (r.to.find(t) == std::string::npos) r.to.push_back(t);           // param.cpp:193-196 duplicates ignored
    }
    for (int i = 0; i < 256; i++) {
        bool acgt = base_idx(i) >= 0 && isalpha(i);
        r.reg[i] = acgt ? 3 : 0;                                            // param.cpp:130-139
        r.conv[i] = r.rconv[i] = r.reg[i];
    }
    for (char t : r.to) {                                                   // param.cpp:202-215
        r.conv[(u8)t] = r.conv[(u8)tolower(t)] = 1;
        for (int j = 0; j < 5; j++) if (NT[j] == t && t != '-') { r.rconv[(u8)CNT[j]] = r.rconv[(u8)tolower(CNT[j])] = 1; }
    }
    int bit[4] = {-1, -1, -1, -1};
    bit[base_idx(r.from)] = 1;                                              // param.cpp:219
    r.single = (r.to.size() == 1 && r.to[0] != '-');
    if (r.single) bit[base_idx(r.to[0])] = 3;                               // param.cpp:222-227
    static const int rest[3] = {0, 2, 3};
    for (int i = 0, j = 0; i < 4; i++) if (bit[i] < 0) bit[i] = rest[j++];  // param.cpp:228-233
    memset(r.code, 0, 256); memset(r.rcode, 0, 256);
    for (int i = 0; i < 4; i++) {                                           // param.cpp:238-253
        r.code[(u8)NT[i]] = r.code[(u8)tolower(NT[i])] = (u8)bit[i];
        r.rcode[(u8)NT[i]] = r.rcode[(u8)tolower(NT[i])] = (u8)bit[3 - i];
        r.letter[bit[i]] = NT[i];                                           // param.cpp:259
    }
    return true;
}

// ---------------------------------------------------------------- myrand (utilities.cpp:38-48)
static u32 myrand(u32 index, u32 seed) {
    u32 add = seed * 1000000u;                      // 32-bit wrap of param.randseed*1000000
    u64 v = ((u64)(int64_t)(int)index + (u64)add) * 3935559000370003845ULL + 2691343689449507681ULL;
    v ^= v >> 21; v ^= v << 37; v ^= v >> 4;
    v *= 4768777513237032717ULL;
    v ^= v << 20; v ^= v >> 41; v ^= v << 5;
    return (u32)v;
}

// ---------------------------------------------------------------- reference + index
struct Hit { u32 loc; u32 chr; int gap; u32 gp; };       // gHit (param.h:35-42); chr = 2*seq+strand

// What a SingleAlign object keeps from read to read (align.h:73-77): the reference never clears xseed_start_offset,
// xseed_array or xseedreg_array, so a read whose start-offset range is empty ((len - I + 1) % s == 0, align.cpp:476-480)
// inherits the start offset of the last read (same object, same chain) that had a range, and seed positions beyond
// len - s still hold the hashes of the last read that was long enough to write them (align.cpp:79-150). With -p 1 one
// object (two for pairs: _sa / _sb) sees every read in input order, which is the deterministic behaviour restated here;
// memory starts zeroed (the object lives on a fresh thread stack).
struct Carry { u32 st0[2]; u32 xseed[2][512]; u8 xreg[2][512]; };

struct Context {
    bsl_params P; Rule R;
    Carry carry[3];                       // aligner objects by ReadInf::readset: 0 single-end, 1 mate #1 (_sa), 2 mate #2 (_sb)
    bool keep_state = false;              // false: every orc_align_* call starts with fresh objects (like one bsl_align_* call)
    // reference (refbase.cpp:186-252)
    std::vector<std::string> names; std::vector<u32> len, rcoff, anchor;
    std::vector<u8> G[2];                 // one code per base, concatenated with margins, both strands
    u64 sum_length = 0; u64 n_words = 0;
    // index (refbase.cpp:261-439)
    u32 K = 0, maxk = 0; std::vector<u32> start, nf, loc;
    u32 prof[MAXSNPS + 1][16];
    bsl_stats st;
    std::string err;
};

struct Block { u32 id, b, e; };

// seed hash of s codes: code 3 folds onto 1, base-3 number, first base most significant (param.h:107-116)
static inline u32 seed_hash(const u8 *c, u32 s) {
    u32 h = 0; for (u32 i = 0; i < s; i++) { u32 d = c[i] == 3 ? 1 : c[i]; h = h * 3 + d; } return h;
}

static void build_index(Context &C, const u8 *cat, const u64 *off, const u32 *lens, u32 nchr) {
    const Rule &R = C.R; const u32 s = C.P.seed_size, I = C.P.index_interval;
    C.len.assign(lens, lens + nchr); C.rcoff.resize(nchr); C.anchor.resize(nchr + 1);
    u64 words = 0; C.anchor[0] = MARGIN * 32; C.sum_length = 0;
    for (u32 c = 0; c < nchr; c++) {                                        // refbase.cpp:64,222-226
        u32 nw = (lens[c] + 31) / 32 + 2; C.rcoff[c] = nw * 32; words += nw;
        C.anchor[c + 1] = (u32)((words + MARGIN) * 32); C.sum_length += lens[c];
    }
    C.n_words = words + 2 * MARGIN;
    C.G[0].assign(C.n_words * 32, 0); C.G[1].assign(C.n_words * 32, 0);
    std::vector<Block> blocks;
    for (u32 c = 0; c < nchr; c++) {
        const u8 *q = cat + off[c]; u32 n = lens[c], P = C.rcoff[c];
        u8 *F = &C.G[0][C.anchor[c]], *Rv = &C.G[1][C.anchor[c]];
        for (u32 k = 0; k < n; k++) { F[k] = R.code[q[k]]; Rv[P - 1 - k] = R.rcode[q[k]]; }   // refbase.cpp:63-101
        // UnmaskRegion (refbase.cpp:103-128): runs from an ACGT char to the next N/X char, >=16 long
        u32 e = 0;
        while (e < n) {
            u32 b = e; while (b < n && !R.reg[q[b]]) b++;
            if (b >= n) break;
            e = b; while (e < n && !(q[e] == 'N' || q[e] == 'X' || q[e] == 'n' || q[e] == 'x')) e++;
            if (e - b < 16) continue;
            blocks.push_back({2 * c, b, e}); blocks.push_back({2 * c + 1, P - e, P - b});
        }
    }
    std::stable_sort(blocks.begin(), blocks.end(), [](const Block &a, const Block &b) { return a.id < b.id || (a.id == b.id && a.b < b.b); });
    C.K = 1; for (u32 i = 0; i < s; i++) C.K *= 3;
    std::vector<u32> cnt0(C.K, 0), cnt1(C.K, 0);
    auto each = [&](u32 strand, auto &&fn) {                                // refbase.cpp:303-325 / 419-439
        for (const Block &bk : blocks) {
            if (bk.id % 2 != strand) continue;
            u32 c = bk.id / 2; const u8 *g = &C.G[strand][C.anchor[c]];
            u32 last = ((bk.e - s) / I) * I;
            for (u32 p = (bk.b / I) * I; p <= last; p += I) fn(seed_hash(g + p, s), C.anchor[c] + p);
        }
    };
    each(0, [&](u32 k, u32) { cnt0[k]++; }); each(1, [&](u32 k, u32) { cnt1[k]++; });
    C.start.assign(C.K + 1, 0); C.nf.assign(C.K, 0);
    u64 run = 0; for (u32 k = 0; k < C.K; k++) { C.start[k] = (u32)run; C.nf[k] = cnt0[k]; run += cnt0[k] + cnt1[k]; }
    C.start[C.K] = (u32)run; C.loc.assign(run, 0);
    std::vector<u32> fill(C.K);
    for (u32 k = 0; k < C.K; k++) fill[k] = C.start[k];
    each(0, [&](u32 k, u32 g) { C.loc[fill[k]++] = g; });                   // forward entries first ...
    each(1, [&](u32 k, u32 g) { C.loc[fill[k]++] = g; });                   // ... then reverse (refbase.cpp:433)
    // over-represented k-mer cut-off (refbase.cpp:362-363): only the first K-1 counts are sorted,
    // rank computed in float32.
    std::vector<u32> tot(C.K); for (u32 k = 0; k < C.K; k++) tot[k] = cnt0[k] + cnt1[k];
    u32 rank = (u32)((float)C.K * (1 - C.P.max_kmer_ratio)) - 1;
    if (rank < C.K - 1) { std::nth_element(tot.begin(), tot.begin() + rank, tot.begin() + (C.K - 1)); }
    C.maxk = tot[rank];
    for (u32 i = 0; i < I; i++) for (u32 j = 0; j <= MAXSNPS; j++) C.prof[j][i] = ((j * s + i + I - 1) / I) * I;   // param.cpp:70-74
}

// ---------------------------------------------------------------- one read, one chain view
struct Shared {                 // state shared by the two chain views of a read (one SingleAlign object)
    u32 thr; std::vector<Hit> lists[2][MAXSNPS + 1];
    std::set<std::pair<u32, u32> > seen, gseen;
    u32 L, B, Rnd, nseg; bool on[2];
};

struct View {
    const Context *C; Shared *S; u32 c;        // read chain
    std::vector<u8> q; std::vector<u8> isN, conv;
    u32 *sh; u8 *sN; u32 *st0;                 // xseed_array / xseedreg_array / xseed_start_offset of this chain: live in the Carry
    u32 st[MAXSNPS + 1]; std::pair<int, int> rank[MAXSNPS + 1];
};

// FilterReads budget (align.cpp:550-561)
static u32 budget(const bsl_params &P, u32 raw_len, u32 len) {
    u32 B = P.max_snp_num < 100 ? P.max_snp_num : (u32)((P.max_snp_num - 100) / 100.0 * raw_len + 0.5);
    if (P.gap > 0) B = B + 1 + P.gap;
    if (B > MAXSNPS) B = MAXSNPS;
    return (B + 1) * (len - 1) / raw_len;
}

static void make_view(View &v, const Context &C, Carry &K, Shared &S, const u8 *seq, u32 chain) {
    const Rule &R = C.R; u32 L = S.L, s = C.P.seed_size;
    v.C = &C; v.S = &S; v.c = chain; v.q.resize(L); v.isN.resize(L); v.conv.resize(L);
    v.sh = K.xseed[chain]; v.sN = K.xreg[chain]; v.st0 = &K.st0[chain];
    for (u32 k = 0; k < L; k++) {                                           // align.cpp:79-226
        u8 ch = chain ? seq[L - 1 - k] : seq[k];
        v.q[k] = chain ? R.rcode[ch] : R.code[ch];
        v.isN[k] = !R.reg[ch];
        v.conv[k] = (chain ? R.rconv[ch] : R.conv[ch]) == 1;
    }
    for (u32 p = 0; p + s <= L; p++) {                                      // entries beyond L - s keep what an earlier read wrote
        v.sh[p] = seed_hash(&v.q[p], s); u8 any = 0; for (u32 k = 0; k < s; k++) any |= v.isN[p + k]; v.sN[p] = any;
    }
}

// CountSeeds (align.cpp:526-540)
static int count_seeds(const View &v, u32 j, u32 st) {
    const Context &C = *v.C; u32 I = C.P.index_interval, total = 0, k = 0;
    for (u32 i = 0; i < I; i++) {
        u32 p = C.prof[j][i] + st - i;
        if (v.sN[p]) k = 12;
        u32 h = v.sh[p]; total += (C.start[h + 1] - C.start[h]) << k;
    }
    if (total == 0) total = 9999999;
    return (int)total;
}

// ReorderSeed + AdjustSeedStartArray (align.cpp:468-524)
static void schedule(View &v) {
    const Context &C = *v.C; u32 L = v.S->L, I = C.P.index_interval, s = C.P.seed_size, nseg = v.S->nseg;
    u32 ii = (L - I + 1) % s, best = 0xffffffffu, st0 = *v.st0;  // empty range: the value the previous read left (SURVEY trap 3)
    for (u32 i = 0; i < ii; i++) {
        u32 tt = 0; for (u32 j = 0; j < nseg; j++) tt += (u32)count_seeds(v, j, i);
        if (tt < best) { best = tt; st0 = i; }
    }
    *v.st0 = st0;
    for (u32 j = 0; j < nseg; j++) v.st[j] = st0;
    for (u32 t = 0; t < nseg; t++) {
        u32 ptr = (t % 2 == 0) ? t / 2 : nseg - 1 - t / 2;
        u32 lo = ptr == 0 ? 0 : v.st[ptr - 1], hi = ptr == nseg - 1 ? ii : v.st[ptr + 1];
        v.st[ptr] = lo; u32 b = 0xffffffffu;
        for (u32 x = lo; x <= hi; x++) { u32 tt = (u32)count_seeds(v, ptr, x); if (tt < b) { b = tt; v.st[ptr] = x; } }
    }
    for (u32 j = 0; j < nseg; j++) v.rank[j] = std::make_pair(count_seeds(v, j, v.st[j]), (int)j);
    std::sort(v.rank, v.rank + nseg);
}

// mismatch rule at one base (align.h:118-131 single conversion, :199-239 multi/'-'); N-mask applied by the caller
static inline bool mism(const View &v, u32 k, u8 ref) {
    u8 q = v.q[k];
    if (v.C->R.single) return ref == 1 ? !(q == 1 || q == 3) : q != ref;
    if (ref == 1 && (v.conv[k] || v.isN[k])) return false;
    return q != ref;
}

// AddHit + int2hit (align.h:329-347, align.cpp:319-346). returns 1 = abort this SnpAlign
static int add_hit(View &v, u32 level, u32 g, u32 sig, int sh, u32 gp) {
    const Context &C = *v.C; Shared &S = *v.S; u32 L = S.L;
    u32 n = (u32)C.len.size(), left = 0, right = n;
    while (left + 1 < right) { u32 mid = (left + right) / 2; if (g >= C.anchor[mid]) left = mid; else right = mid; }
    u32 x = g - C.anchor[left];
    if (sig) { x = C.rcoff[left] - L - x; gp = (u32)((int)L + (sh < 0 ? sh : 0) - (int)gp) & 511; x -= (u32)sh; }
    if ((int)x < 0) return 0;
    if (x + L > C.len[left]) return 0;
    auto &ss = sh ? S.gseen : S.seen;
    if (!ss.insert(std::make_pair(left, x)).second) return 0;
    S.lists[v.c][level].push_back({x, left * 2 + sig, sh, gp & 511});
    if (S.lists[0][level].size() + S.lists[1][level].size() >= C.P.max_num_hits) {
        if (level == 0) return 1;
        S.thr = level - 1;
    }
    return 0;
}

// GapAlign (align.cpp:348-410) over MismatchPattern0/1 (align.h:133-196, 241-327; no N-mask)
static int gap_align(View &v, u32 g, u32 sig, u32 h) {
    const Context &C = *v.C; Shared &S = *v.S; u32 L = S.L, thr = S.thr, s = C.P.seed_size;
    if (thr < 2) return 0;
    const u8 *ref = &C.G[sig][0];
    std::vector<u32> P0; u32 ret0 = L;
    for (u32 k = 0; k < L && P0.size() < thr - 1; k++) if (mism(v, k, ref[g + k])) { P0.push_back(k); if (P0.size() == thr - 1) ret0 = k; }
    P0.resize(thr - 1, L);
    if (ret0 < h + s) return 0;
    for (u32 tt = 1; tt <= 2 * C.P.gap; tt++) {
        u32 t = (tt + 1) / 2; int sh = (tt % 2) ? -(int)t : (int)t; int sh1 = sh < 0 ? sh : 0;
        if (thr < 1 + t) break;
        u32 g1 = g + (u32)sh; std::vector<u32> PR;
        for (u32 k = 0; k < L && PR.size() < thr - 1; k++) if (mism(v, L - 1 - k, ref[g1 + L - 1 - k])) PR.push_back(k);
        PR.resize(thr - 1, L);
        u32 rl = L - t - 1;
        for (u32 i = 0; i < thr - t; i++) {
            u32 gp = P0[i];
            if (gp < 6 || gp >= rl) continue;
            for (u32 j = 0; j < thr - t - i; j++) {
                u32 m2 = PR[j];
                if (m2 < 6 || m2 >= rl) continue;
                if ((int)gp + (int)m2 - sh1 < (int)L) continue;
                int clip = (int)gp + 6 - (int)L - sh1;
                if (clip > 0) gp -= (u32)clip;
                return add_hit(v, i + j + t, g, sig, sh, gp);
            }
        }
    }
    return 0;
}

// SnpAlign, WGBS branch (align.cpp:274-316) for one chain view
static int snp_align(View &v, u32 r, bsl_stats &st) {
    const Context &C = *v.C; Shared &S = *v.S; u32 L = S.L, I = C.P.index_interval;
    u32 j = (u32)v.rank[r].second;
    for (u32 i = 0; i < I; i++) {
        u32 h = C.prof[j][i] + v.st[j] - i, k = v.sh[h], m = C.start[k + 1] - C.start[k];
        st.seed_lookups++;
        if (m == 0 || m > C.maxk) continue;
        st.candidates += m;
        int mc = (int)C.nf[k] - 1; u32 base = C.start[k], jj = S.Rnd % m;
        for (u32 t = 0; t < m; t++, jj++) {
            if (jj >= m) jj -= m;
            u32 sig = ((int)jj > mc) ? 1 : 0, g = C.loc[base + jj] - h;
            const u8 *ref = &C.G[sig][g];
            u32 snp = 0;
            for (u32 x = 0; x < L && snp <= S.thr; x++) if (!v.isN[x] && mism(v, x, ref[x])) snp++;
            if (snp <= S.thr) { if (add_hit(v, snp, g, sig, 0, 0)) return 1; }
            if (C.P.gap > 0) { if (gap_align(v, g, sig, h)) return 1; }
        }
    }
    return 0;
}

struct ReadState { Shared S; View v[2]; bool filtered; };

static void chain_flags(const bsl_params &P, u32 readset, bool on[2]) {     // align.cpp:83-84
    on[0] = (P.chains == 1) || ((P.chains <= 1) == (readset < 2));
    on[1] = (P.chains == 1) || ((P.chains <= 1) == (readset == 2));
}

static bool prepare(ReadState &rs, Context &C, const u8 *seq, u32 len, u32 raw_len, u32 index, u32 readset) {
    const bsl_params &P = C.P; Shared &S = rs.S;
    rs.filtered = true; S.L = len; S.B = 0;
    if (len < P.min_read_size || len == 0) return false;
    u32 ns = 0; for (u32 k = 0; k < len; k++) ns += !C.R.reg[seq[k]];
    if (ns > P.max_ns) return false;                                        // align.cpp:559-560
    rs.filtered = false;
    S.B = budget(P, raw_len, len); S.thr = S.B; S.Rnd = myrand(index, P.randseed);
    S.nseg = std::min((u32)((len - P.index_interval + 1) / P.seed_size), S.B + 1);   // align.cpp:450
    chain_flags(P, readset, S.on);
    for (u32 c = 0; c < 2; c++) if (S.on[c]) { make_view(rs.v[c], C, C.carry[readset < 3 ? readset : 0], S, seq, c); schedule(rs.v[c]); }
    return true;
}

static int snp_align_both(ReadState &rs, u32 r, bsl_stats &st) {
    for (u32 c = 0; c < 2; c++) if (rs.S.on[c]) if (snp_align(rs.v[c], r, st)) return 1;
    return 0;
}

static void fill_hit(bsl_hit &o, const Hit &h, u32 chain, u32 level) {
    o.loc = h.loc; o.chr = h.chr; o.gap_size = h.gap; o.gap_pos = (uint16_t)h.gp; o.nm = (u8)level; o.read_chain = (u8)chain;
}

// lowest non-empty level + -r 1 pick (align.cpp:583-612, pairs.cpp:232-259)
static void report_single(const ReadState &rs, bsl_hit &o, std::vector<bsl_hit> *all) {
    const Shared &S = rs.S; memset(&o, 0, sizeof o);
    o.read_len = (uint16_t)S.L; o.max_snp = (u8)S.B;
    if (rs.filtered) { o.status = BSL_ST_FILTERED; return; }
    for (u32 l = 0; l <= S.B; l++) {
        u32 n0 = (u32)S.lists[0][l].size(), n = n0 + (u32)S.lists[1][l].size();
        if (!n) continue;
        u32 pick = n == 1 ? 0 : S.Rnd % n; u32 chain = pick < n0 ? 0 : 1;
        fill_hit(o, S.lists[chain][l][pick - (chain ? n0 : 0)], chain, l);
        o.n_hits = n; o.n_chain0 = n0; o.status = n == 1 ? BSL_ST_UNIQUE : BSL_ST_MULTI;
        if (all) { o.all_first = (u32)all->size();
            for (u32 c = 0; c < 2; c++) for (const Hit &h : S.lists[c][l]) { bsl_hit a = o; fill_hit(a, h, c, l); all->push_back(a); } }
        return;
    }
    o.status = BSL_ST_UNMAPPED;
}

// RunAlign (align.cpp:446-466)
static void run_single(ReadState &rs, bsl_stats &st) {
    Shared &S = rs.S;
    for (u32 r = 0; r < S.nseg; r++) {
        snp_align_both(rs, r, st);
        for (u32 l = 0; l <= r; l++) if (!S.lists[0][l].empty() || !S.lists[1][l].empty()) return;
    }
}

// ---------------------------------------------------------------- pairs (pairs.cpp:29-177)
struct PairRec { u32 chain, na, nb, insert; Hit a, b; };

static int get_pairs(ReadState &A, ReadState &B, u32 na, u32 nb, std::vector<PairRec> ph[], const bsl_params &P) {
    if (na > A.S.B || nb > B.S.B) return 0;
    int npair = 0;
    for (u32 chain = 0; chain < 2; chain++) {
        std::vector<Hit> &al = A.S.lists[chain][na], &bl = B.S.lists[1 - chain][nb];
        u32 chra = ~0u; size_t bs = 0, be = 0;
        for (size_t i = 0; i < al.size(); i++) {
            if (chra != al[i].chr) {
                chra = al[i].chr;
                for (bs = be; bs < bl.size(); bs++) if (bl[bs].chr >= chra) break;
                for (be = bs; be < bl.size(); be++) if (bl[be].chr > chra) break;
            }
            for (size_t j = bs; j < be; j++) {
                u32 s0, e0; bool a_first = chain == 0 ? !(chra & 1) : (chra & 1);
                if (a_first) { s0 = al[i].loc; e0 = bl[j].loc + B.S.L; } else { s0 = bl[j].loc; e0 = al[i].loc + A.S.L; }
                u32 ins = e0 - s0;
                if (ins >= P.min_insert && ins <= P.max_insert) {
                    ph[na + nb].push_back({chain, na, nb, ins, al[i], bl[j]}); npair++;
                    if (ph[na + nb].size() >= P.max_num_hits) return npair;
                }
            }
        }
    }
    return npair;
}

static bool hit_less(const Hit &a, const Hit &b) { return a.chr < b.chr || (a.chr == b.chr && a.loc < b.loc); }   // utilities.cpp:51

static int run_pair(ReadState &A, ReadState &B, std::vector<PairRec> ph[], const bsl_params &P, bsl_stats &st) {
    u32 maxi = std::max(A.S.B, B.S.B); int n = 0;
    for (u32 i = 0; i <= maxi; i++) {
        if (i < A.S.nseg) snp_align_both(A, i, st);
        if (i < B.S.nseg) snp_align_both(B, i, st);
        if (i <= A.S.B) for (u32 c = 0; c < 2; c++) std::sort(A.S.lists[c][i].begin(), A.S.lists[c][i].end(), hit_less);   // align.cpp:412-416
        if (i <= B.S.B) for (u32 c = 0; c < 2; c++) std::sort(B.S.lists[c][i].begin(), B.S.lists[c][i].end(), hit_less);
        n += get_pairs(A, B, i, i, ph, P);
        for (u32 j = 0; j < i; j++) { n += get_pairs(A, B, i, j, ph, P); n += get_pairs(A, B, j, i, ph, P); }
        if (n > 0) return 1;
    }
    return n;
}

} // namespace orc

// ================================================================== C interface (ctypes / tests)
using namespace orc;
struct orc_ctx { Context C; };

extern "C" {

int orc_ctx_create(orc_ctx **out, const bsl_params *p) {
    orc_ctx *c = new orc_ctx(); c->C.P = *p; memset(&c->C.st, 0, sizeof c->C.st); memset(c->C.carry, 0, sizeof c->C.carry);
    std::string rule = std::string(1, p->from_base) + ":" + p->to_bases;
    if (!make_rule(rule, c->C.R)) { fprintf(stderr, "%s\n", c->C.R.err.c_str()); delete c; return BSL_EINVAL; }
    *out = c; return 0;
}
void orc_ctx_destroy(orc_ctx *c) { delete c; }

int orc_rule_tables(const orc_ctx *c, uint8_t *code, uint8_t *rcode, uint8_t *conv, uint8_t *rconv, char *letter4) {
    memcpy(code, c->C.R.code, 256); memcpy(rcode, c->C.R.rcode, 256); memcpy(conv, c->C.R.conv, 256); memcpy(rconv, c->C.R.rconv, 256);
    memcpy(letter4, c->C.R.letter, 4); return c->C.R.single;
}

int orc_index_build(orc_ctx *c, const uint8_t *cat, const uint64_t *off, const uint32_t *len, uint32_t n) {
    build_index(c->C, cat, off, len, n); return 0;
}
int orc_index_info_get(const orc_ctx *c, bsl_index_info *o) {
    memset(o, 0, sizeof *o); o->n_seq = (u32)c->C.len.size(); o->n_kmers = c->C.K; o->sum_length = c->C.sum_length;
    o->n_words = c->C.n_words; o->n_entries = c->C.loc.size(); o->max_kmer_num = c->C.maxk; return 0;
}
// planes are returned packed exactly like the reference's refcat/crefcat (base k of a word at bits 63-2k..62-2k)
int orc_index_download(const orc_ctx *c, uint32_t *bucket_start, uint32_t *n_fwd, uint32_t *loc, uint64_t *fwd, uint64_t *rc) {
    const Context &C = c->C;
    if (bucket_start) memcpy(bucket_start, C.start.data(), C.start.size() * 4);
    if (n_fwd) memcpy(n_fwd, C.nf.data(), C.nf.size() * 4);
    if (loc) memcpy(loc, C.loc.data(), C.loc.size() * 4);
    for (int s = 0; s < 2; s++) { uint64_t *o = s ? rc : fwd; if (!o) continue;
        for (u64 w = 0; w < C.n_words; w++) { u64 x = 0; for (int k = 0; k < 32; k++) x = (x << 2) | C.G[s][w * 32 + k]; o[w] = x; } }
    return 0;
}

uint32_t orc_read_budget(const bsl_params *p, uint32_t raw_len, uint32_t len) { return budget(*p, raw_len, len); }
uint32_t orc_myrand(uint32_t index, uint32_t seed) { return myrand(index, seed); }

static inline u32 rd_index(const bsl_batch *b, u32 i) { return b->index ? b->index[i] : b->first_index + i; }
static inline u32 rd_raw(const bsl_batch *b, u32 i, u32 len) { return b->raw_len ? b->raw_len[i] : len; }

// keep = 1: the aligner objects live as long as the context (the stand-alone CLI maps one read per call);
// keep = 0 (default): every call starts with fresh objects, which is what one bsl_align_se / bsl_align_pe call does
void orc_keep_state(orc_ctx *c, int keep) { c->C.keep_state = keep != 0; }

int orc_align_se(orc_ctx *c, const bsl_batch *b, bsl_hit *out, bsl_hit *all, uint64_t all_cap, uint64_t *n_all) {
    Context &C = c->C; std::vector<bsl_hit> allv; bool want_all = C.P.report_repeat_hits == 2 && all;
    if (!C.keep_state) memset(C.carry, 0, sizeof C.carry);
    for (u32 i = 0; i < b->n; i++) {
        const u8 *seq = b->bases + b->offsets[i]; u32 len = (u32)(b->offsets[i + 1] - b->offsets[i]);
        ReadState rs;
        if (i < b->n_context) {                                             // context read: leaves its state behind, is not mapped
            prepare(rs, C, seq, len, rd_raw(b, i, len), rd_index(b, i), b->readset);
            memset(&out[i], 0, sizeof out[i]); out[i].status = BSL_ST_FILTERED; continue;
        }
        C.st.reads++;
        if (prepare(rs, C, seq, len, rd_raw(b, i, len), rd_index(b, i), b->readset)) run_single(rs, C.st);
        report_single(rs, out[i], want_all ? &allv : nullptr);
    }
    if (n_all) *n_all = allv.size();
    if (want_all) memcpy(all, allv.data(), std::min<u64>(allv.size(), all_cap) * sizeof(bsl_hit));
    return 0;
}

int orc_align_pe(orc_ctx *c, const bsl_batch *a, const bsl_batch *b, bsl_hit *oa, bsl_hit *ob, bsl_pair *op,
                 bsl_hit *all_a, bsl_hit *all_b, uint64_t all_cap, uint64_t *n_all) {
    Context &C = c->C; if (a->n != b->n) return BSL_EINVAL;
    std::vector<bsl_hit> va, vb; bool want_all = C.P.report_repeat_hits == 2 && all_a && all_b;
    if (!C.keep_state) memset(C.carry, 0, sizeof C.carry);
    for (u32 i = 0; i < a->n; i++) {
        const u8 *sa = a->bases + a->offsets[i], *sb = b->bases + b->offsets[i];
        u32 la = (u32)(a->offsets[i + 1] - a->offsets[i]), lb = (u32)(b->offsets[i + 1] - b->offsets[i]);
        ReadState A, B;
        bool okA = prepare(A, C, sa, la, rd_raw(a, i, la), rd_index(a, i), a->readset);
        bool okB = prepare(B, C, sb, lb, rd_raw(b, i, lb), rd_index(b, i), b->readset);
        memset(&op[i], 0, sizeof op[i]);
        if (i < a->n_context) {                                             // context pair: leaves its state behind, is not mapped
            memset(&oa[i], 0, sizeof oa[i]); memset(&ob[i], 0, sizeof ob[i]); oa[i].status = ob[i].status = BSL_ST_FILTERED; continue;
        }
        C.st.reads += 2;
        std::vector<PairRec> ph[2 * MAXSNPS + 1]; int paired = 0;
        if (okA && okB) paired = run_pair(A, B, ph, C.P, C.st);
        else { if (okA) run_single(A, C.st); if (okB) run_single(B, C.st); }
        bool reported = false;
        if (paired) for (u32 l = 0; l <= 2 * MAXSNPS; l++) if (!ph[l].empty()) {        // pairs.cpp:204-230
            u32 k = (u32)ph[l].size(); op[i].n_pairs = k;
            if (k > 1 && C.P.report_repeat_hits == 0) break;                             // suppressed: mates reported unpaired
            u32 pick = k == 1 ? 0 : A.S.Rnd % k; const PairRec &pr = ph[l][pick];
            memset(&oa[i], 0, sizeof oa[i]); memset(&ob[i], 0, sizeof ob[i]);
            oa[i].read_len = (uint16_t)A.S.L; ob[i].read_len = (uint16_t)B.S.L; oa[i].max_snp = (u8)A.S.B; ob[i].max_snp = (u8)B.S.B;
            fill_hit(oa[i], pr.a, pr.chain, pr.na); fill_hit(ob[i], pr.b, 1 - pr.chain, pr.nb);
            oa[i].status = ob[i].status = BSL_ST_PAIRED; oa[i].n_hits = ob[i].n_hits = k;
            op[i].insert = pr.insert; op[i].chain = (u8)pr.chain; op[i].na = (u8)pr.na; op[i].nb = (u8)pr.nb;
            if (want_all) { op[i].all_first = (u32)va.size();
                for (const PairRec &x : ph[l]) { bsl_hit ha = oa[i], hb = ob[i]; fill_hit(ha, x.a, x.chain, x.na); fill_hit(hb, x.b, 1 - x.chain, x.nb);
                    ha.all_first = hb.all_first = x.insert; va.push_back(ha); vb.push_back(hb); } }
            reported = true; break;
        }
        if (!reported) {                                                                 // pairs.cpp:232-305: -r 2 lists every hit of a multi-hit mate
            report_single(A, oa[i], want_all ? &va : nullptr); vb.resize(va.size());
            report_single(B, ob[i], want_all ? &vb : nullptr); va.resize(vb.size());
        }
    }
    if (n_all) *n_all = va.size();
    if (want_all) { u64 k = std::min<u64>(va.size(), all_cap); memcpy(all_a, va.data(), k * sizeof(bsl_hit)); memcpy(all_b, vb.data(), k * sizeof(bsl_hit)); }
    return 0;
}

int orc_stats_get(const orc_ctx *c, bsl_stats *st) { *st = c->C.st; return 0; }

} // extern "C"

// ================================================================== stand-alone CLI: SAM like the reference prints it
#ifdef ORACLE_MAIN
namespace {

struct Opts {
    std::string a, b, d, o, M; bsl_params P; bool header = true, unmap = false, outref = false;
    u32 max_readlen = 480, read_start = 1, read_end = ~0u; std::string cmdline;
};

struct Rd { std::string name, seq, qual; u32 index; };

// token-wise FASTA/FASTQ reader (reads.cpp:42-84)
static bool next_read(std::istream &in, bool fq, Rd &r, u32 max_len) {
    char c; in >> c; if (in.eof() || !in) return false;
    std::string rest; in >> r.name; std::getline(in, rest); in >> r.seq;
    if (fq) { std::string plus; in >> plus; std::getline(in, rest); in >> r.qual; }
    else r.qual = std::string(r.seq.size(), (char)('!' + 40));
    if (r.seq.size() > max_len) { r.seq.erase(max_len); r.qual.erase(max_len); }
    return true;
}

static std::string revcomp(const std::string &s) {
    std::string o(s.rbegin(), s.rend());
    for (char &ch : o) switch (ch) { case 'A': ch = 'T'; break; case 'C': ch = 'G'; break; case 'G': ch = 'C'; break; case 'T': ch = 'A'; break;
        case 'a': ch = 't'; break; case 'c': ch = 'g'; break; case 'g': ch = 'c'; break; case 't': ch = 'a'; break; default: ch = 'N'; }
    return o;
}

static std::string cigar(const bsl_hit &h) {                                   // align.cpp:641-643
    char b[64]; int L = h.read_len;
    if (h.gap_size == 0) snprintf(b, sizeof b, "%uM", (unsigned)L);
    else if (h.gap_size > 0) snprintf(b, sizeof b, "%dM%dD%dM", (int)h.gap_pos, h.gap_size, L - (int)h.gap_pos);
    else snprintf(b, sizeof b, "%dM%dI%dM", (int)h.gap_pos, -h.gap_size, L - (int)h.gap_pos + h.gap_size);
    return b;
}

static std::string xr_tag(const Context &C, const bsl_hit &h) {                // align.cpp:646-658
    std::string m; const u8 *F = &C.G[0][C.anchor[h.chr >> 1]];
    for (u32 ii = 2; ii > 0; ii--) { if (h.loc < ii) continue; m.push_back((char)(C.R.letter[F[h.loc - ii]] + 32)); }
    for (u32 ii = 0; ii < (u32)h.read_len + 2; ii++) m.push_back(C.R.letter[F[h.loc + ii]]);
    m[m.size() - 1] += 32; m[m.size() - 2] += 32;
    return "\tXR:Z:" + m;
}

// s_OutHit (align.cpp:616-669)
static void out_single(std::string &os, const Opts &O, const Context &C, const Rd &r, u32 readset, const bsl_hit &h, int n) {
    char buf[4096]; int flag = 0x40 * readset;
    if (n <= 0) { if (!O.unmap) return; flag |= n < 0 ? 0x204 : 0x4;
        snprintf(buf, sizeof buf, "%s\t%d\t*\t0\t0\t*\t*\t0\t0\t%s\t%s\n", r.name.c_str(), flag, r.seq.c_str(), r.qual.c_str()); os += buf; return; }
    u32 rev = h.read_chain ^ (h.chr & 1);
    if (n > 1) flag |= 0x100;
    if (rev) flag |= 0x10;
    std::string seq = rev ? revcomp(r.seq) : r.seq, q = rev ? std::string(r.qual.rbegin(), r.qual.rend()) : r.qual;
    snprintf(buf, sizeof buf, "%s\t%d\t%s\t%u\t255\t%s\t*\t0\t0\t%s\t%s\tNM:i:%d", r.name.c_str(), flag, C.names[h.chr >> 1].c_str(), h.loc + 1,
             cigar(h).c_str(), seq.c_str(), q.c_str(), (int)h.nm);
    os += buf; if (O.outref) os += xr_tag(C, h);
    snprintf(buf, sizeof buf, "\tZS:Z:%c%c\n", "+-"[h.chr & 1], "+-"[h.read_chain]); os += buf;
}

// s_OutHitPair (pairs.cpp:307-416)
static void out_pair(std::string &os, const Opts &O, const Context &C, const Rd &ra, const Rd &rb, const bsl_hit &a, const bsl_hit &b, const bsl_pair &p, int n) {
    char buf[4096];
    for (int side = 0; side < 2; side++) {
        const bsl_hit &me = side ? b : a, &mate = side ? a : b; const Rd &r = side ? rb : ra;
        u32 chain = side ? !p.chain : p.chain; u32 rev = chain ^ (me.chr & 1);
        int flag = 0x3; if (n > 1) flag |= 0x100; int ins;
        if (rev) { flag |= 0x10; ins = -(int)p.insert; } else { flag |= 0x20; ins = (int)p.insert; }
        flag |= 0x40 * (side + 1);
        std::string seq = rev ? revcomp(r.seq) : r.seq, q = rev ? std::string(r.qual.rbegin(), r.qual.rend()) : r.qual;
        snprintf(buf, sizeof buf, "%s\t%d\t%s\t%u\t255\t%s\t=\t%u\t%d\t%s\t%s\tNM:i:%d", r.name.c_str(), flag, C.names[me.chr >> 1].c_str(), me.loc + 1,
                 cigar(me).c_str(), mate.loc + 1, ins, seq.c_str(), q.c_str(), (int)me.nm);
        os += buf; if (O.outref) os += xr_tag(C, me);
        snprintf(buf, sizeof buf, "\tZS:Z:%c%c\n", "+-"[me.chr & 1], "+-"[chain]); os += buf;
    }
}

// s_OutHitUnpair (pairs.cpp:418-485)
static void out_unpair(std::string &os, const Opts &O, const Context &C, const Rd &r, int side, u32 chain_a, u32 chain_b, int ma, u32 na,
                       const bsl_hit &ha, int mb, const bsl_hit &hb) {
    char buf[4096]; int flag = 1 | (0x40 * (side + 1)); u32 rev = chain_a ^ (ha.chr & 1);
    if (ma <= 0) {
        if (ma < 0) flag |= 0x204; if (ma == 0) flag |= 0x4;
        if (mb <= 0) { flag |= 0x8; snprintf(buf, sizeof buf, "%s\t%d\t*\t0\t0\t*\t*\t0\t0\t%s\t%s\n", r.name.c_str(), flag, r.seq.c_str(), r.qual.c_str()); }
        else { if (chain_b ^ (hb.chr & 1)) flag |= 0x20;
            snprintf(buf, sizeof buf, "%s\t%d\t*\t0\t0\t*\t%s\t%u\t0\t%s\t%s\n", r.name.c_str(), flag, C.names[hb.chr >> 1].c_str(), hb.loc + 1, r.seq.c_str(), r.qual.c_str()); }
        os += buf; return;
    }
    if (ma > 1) flag |= 0x100; if (rev) flag |= 0x10;
    std::string seq = rev ? revcomp(r.seq) : r.seq, q = rev ? std::string(r.qual.rbegin(), r.qual.rend()) : r.qual;
    if (mb <= 0) { flag |= 0x8;
        snprintf(buf, sizeof buf, "%s\t%d\t%s\t%u\t255\t%s\t*\t0\t0\t%s\t%s\tNM:i:%d", r.name.c_str(), flag, C.names[ha.chr >> 1].c_str(), ha.loc + 1, cigar(ha).c_str(), seq.c_str(), q.c_str(), (int)na); }
    else { if (chain_b ^ (hb.chr & 1)) flag |= 0x20;
        snprintf(buf, sizeof buf, "%s\t%d\t%s\t%u\t255\t%s\t%s\t%u\t0\t%s\t%s\tNM:i:%d", r.name.c_str(), flag, C.names[ha.chr >> 1].c_str(), ha.loc + 1, cigar(ha).c_str(),
                 C.names[hb.chr >> 1].c_str(), hb.loc + 1, seq.c_str(), q.c_str(), (int)na); }
    os += buf; if (O.outref) os += xr_tag(C, ha);
    snprintf(buf, sizeof buf, "\tZS:Z:%c%c\n", "+-"[ha.chr & 1], "+-"[chain_a]); os += buf;
}

// FixPairReadName (pairs.cpp:487-507)
static void fix_names(std::string &a, std::string &b) {
    if (a == b) return; int d = -1; size_t i, n = std::min(a.size(), b.size());
    for (i = 0; i < n; i++) { if (a[i] != b[i]) break; else if (isdigit((unsigned char)a[i])) d = (int)i; }
    if (i > 0) { if (d < 0) d = (int)i - 1; a.erase(d + 1); b.erase(d + 1); }
    else { fprintf(stderr, "Error: Paired reads name not match:\n%s\n%s\n", a.c_str(), b.c_str()); exit(1); }
}

static int mate_count(const bsl_hit &h) { return h.status == BSL_ST_FILTERED ? -1 : (h.status == BSL_ST_UNMAPPED ? 0 : (int)h.n_hits); }

} // namespace

int main(int argc, char **argv) {
    Opts O; bsl_params &P = O.P; memset(&P, 0, sizeof P);
    P.seed_size = 16; P.index_interval = 4; P.max_snp_num = 110; P.gap = 0; P.max_num_hits = 100; P.min_insert = 28; P.max_insert = 1000;
    P.chains = 0; P.report_repeat_hits = 1; P.randseed = 0; P.max_ns = 5; P.min_read_size = 16; P.max_kmer_ratio = 5e-7f;
    O.cmdline = argv[0]; for (int i = 1; i < argc; i++) O.cmdline += std::string(" ") + argv[i];
    for (int i = 1; i < argc; i++) {                                            // main.cpp:272-364 (subset: the hot-path flags)
        std::string f = argv[i]; if (f.size() < 2 || f[0] != '-') { fprintf(stderr, "unknown option: %s\n", argv[i]); return i; }
        auto val = [&]() -> std::string { if (f.size() > 2 && f[2] == '=') return f.substr(3); return std::string(argv[++i]); };
        switch (f[1]) {
        case 'a': O.a = val(); break; case 'b': O.b = val(); break; case 'd': O.d = val(); break; case 'o': O.o = val(); break; case 'M': O.M = val(); break;
        case 's': P.seed_size = atoi(val().c_str()); P.min_read_size = P.seed_size + P.index_interval - 1; break;       // param.cpp:108-115
        case 'I': P.index_interval = atoi(val().c_str()); break;
        case 'v': { double t = atof(val().c_str());
            if (t < 1.0) { P.max_snp_num = (int)(t * 100 + 0.5) + 100; if (P.max_snp_num == 100) P.max_snp_num = 0; }
            else { P.max_snp_num = (int)(t + 0.5); if (P.max_snp_num > MAXSNPS) P.max_snp_num = MAXSNPS; } break; }
        case 'g': P.gap = std::min(atoi(val().c_str()), 3); break;
        case 'w': P.max_num_hits = atoi(val().c_str()); break; case 'm': P.min_insert = atoi(val().c_str()); break; case 'x': P.max_insert = atoi(val().c_str()); break;
        case 'n': P.chains = atoi(val().c_str()); break; case 'r': P.report_repeat_hits = atoi(val().c_str()); break; case 'S': P.randseed = atoi(val().c_str()); break;
        case 'f': P.max_ns = atoi(val().c_str()); break; case 'k': P.max_kmer_ratio = (float)atof(val().c_str()); break; case 'p': val(); break; case 'V': val(); break;
        case 'L': O.max_readlen = atoi(val().c_str()); break;
        case 'B': O.read_start = std::max(atoi(val().c_str()), 1); break; case 'E': O.read_end = atoi(val().c_str()); break;
        case 'R': O.outref = true; break; case 'H': O.header = false; break; case 'u': O.unmap = true; break;
        default: fprintf(stderr, "unknown option: %s\n", argv[i]); return i;
        }
    }
    if (O.M.size() < 2) { fprintf(stderr, "\n-M option is required\n"); return 1; }
    P.from_base = O.M[0]; strncpy(P.to_bases, O.M.c_str() + 2, 6);
    orc_ctx *ctx; if (orc_ctx_create(&ctx, &P)) return 1; Context &C = ctx->C;
    orc_keep_state(ctx, 1);                                                     // one read (pair) per call below: the aligner objects must outlive the calls
    { Rule chk; if (!make_rule(O.M, chk)) { fprintf(stderr, "%s\n", chk.err.c_str()); return 1; } }
    // FASTA: name = first token after '>', sequence tokens concatenated (refbase.cpp:17-61)
    std::ifstream fd(O.d.c_str()); if (!fd) { fprintf(stderr, "\nfailed to open reference file (check -d option): %s\n", O.d.c_str()); return 1; }
    std::vector<u8> cat; std::vector<u64> off; std::vector<u32> len; std::string line;
    while (std::getline(fd, line)) {
        if (!line.empty() && line[0] == '>') { std::istringstream ss(line.substr(1)); std::string nm; ss >> nm; C.names.push_back(nm); off.push_back(cat.size()); len.push_back(0); }
        else { std::istringstream ss(line); std::string tok; while (ss >> tok) { cat.insert(cat.end(), tok.begin(), tok.end()); len.back() += (u32)tok.size(); } }
    }
    orc_index_build(ctx, cat.data(), off.data(), len.data(), (u32)len.size());
    FILE *fout = O.o.empty() ? stdout : fopen(O.o.c_str(), "wb");
    if (!fout) { fprintf(stderr, "\nfailed to open output file (check -o option): %s\n", O.o.c_str()); return 1; }
    std::string os;
    if (O.header) { os = "@HD\tVN:1.0\n"; char b[1024];                        // main.cpp:516-526
        for (size_t i = 0; i < len.size(); i++) { snprintf(b, sizeof b, "@SQ\tSN:%s\tLN:%u\n", C.names[i].c_str(), len[i]); os += b; }
        os += "@PG\tID:BASAL\tVN:1.8.1\tCL:\"" + O.cmdline + "\"\n"; }
    std::ifstream fa(O.a.c_str()), fb; if (!fa) { fprintf(stderr, "\nfailed to open read file (check -a option): %s\n", O.a.c_str()); return 1; }
    bool fq = fa.peek() == '@'; bool pe = !O.b.empty(); if (pe) fb.open(O.b.c_str());
    Rd ra, rb; u32 idx = 0; u64 n_al = 0, n_un = 0, n_mu = 0;
    for (u32 skip = 1; skip < O.read_start; skip++) { next_read(fa, fq, ra, O.max_readlen); if (pe) next_read(fb, fq, rb, O.max_readlen); idx++; }
    while (idx < O.read_end && next_read(fa, fq, ra, O.max_readlen)) {
        if (pe && !next_read(fb, fq, rb, O.max_readlen)) break;
        u64 offs[2]; bsl_batch ba; memset(&ba, 0, sizeof ba); ba.n = 1; ba.first_index = idx; ba.offsets = offs; offs[0] = 0;
        if (!pe) {
            ba.readset = 0; ba.bases = (const u8 *)ra.seq.data(); offs[1] = ra.seq.size();
            std::vector<bsl_hit> all(2 * 1001 * 16); u64 nall = 0; bsl_hit h;
            orc_align_se(ctx, &ba, &h, all.data(), all.size(), &nall);
            if (h.status == BSL_ST_FILTERED) out_single(os, O, C, ra, 0, h, -1);
            else if (h.status == BSL_ST_UNMAPPED) out_single(os, O, C, ra, 0, h, 0);
            else if (h.status == BSL_ST_UNIQUE) { n_al++; n_un++; out_single(os, O, C, ra, 0, h, 1); }
            else { n_mu++;                                                   // align.cpp:597-611
                if (P.report_repeat_hits == 1) { n_al++; out_single(os, O, C, ra, 0, h, (int)h.n_hits); }
                else if (P.report_repeat_hits == 2) { n_al++; for (u64 k = 0; k < nall; k++) out_single(os, O, C, ra, 0, all[k], (int)h.n_hits); }
                else out_single(os, O, C, ra, 0, h, 0); }
        } else {
            u64 offb[2] = {0, rb.seq.size()}; bsl_batch bb = ba; ba.readset = 1; bb.readset = 2;
            ba.bases = (const u8 *)ra.seq.data(); offs[1] = ra.seq.size(); bb.bases = (const u8 *)rb.seq.data(); bb.offsets = offb;
            bsl_hit ha, hb; bsl_pair pr; std::vector<bsl_hit> alla(2 * 1001 * 16 * 2), allb(alla.size()); u64 nall = 0;
            orc_align_pe(ctx, &ba, &bb, &ha, &hb, &pr, alla.data(), allb.data(), alla.size(), &nall);
            fix_names(ra.name, rb.name);
            bool reported = false;
            if (ha.status == BSL_ST_PAIRED) { reported = true;
                if (pr.n_pairs > 1 && P.report_repeat_hits == 2) { for (u64 k = 0; k < nall; k++) { bsl_pair q = pr; q.insert = alla[k].all_first; q.chain = alla[k].read_chain;
                        out_pair(os, O, C, ra, rb, alla[k], allb[k], q, (int)pr.n_pairs); } }
                else out_pair(os, O, C, ra, rb, ha, hb, pr, (int)pr.n_pairs); }
            if (!reported) {                                                 // StringAlignUnpair (pairs.cpp:232-305), -r 0/1
                int ma = mate_count(ha), mb = mate_count(hb);
                int ma1 = (ma > 1 && P.report_repeat_hits == 0) ? 0 : ma, mb1 = (mb > 1 && P.report_repeat_hits == 0) ? 0 : mb;
                u32 ca = ha.read_chain, cb = hb.read_chain;
                if (ma <= 0) { if (O.unmap) out_unpair(os, O, C, ra, 0, 0, cb, ma, 0, ha, mb1, hb); }
                else if (ma == 1 || P.report_repeat_hits == 1) out_unpair(os, O, C, ra, 0, ca, cb, ma, ha.nm, ha, mb1, hb);
                else if (P.report_repeat_hits == 2) { for (int k = 0; k < ma; k++) { const bsl_hit &x = alla[ha.all_first + k]; out_unpair(os, O, C, ra, 0, x.read_chain, cb, ma, ha.nm, x, mb1, hb); } }   // pairs.cpp:270-274
                else if (P.report_repeat_hits == 0 && O.unmap) out_unpair(os, O, C, ra, 0, 0, cb, 0, 0, ha, mb1, hb);
                if (mb <= 0) { if (O.unmap) out_unpair(os, O, C, rb, 1, 0, ca, mb, 0, hb, ma1, ha); }
                else if (mb == 1 || P.report_repeat_hits == 1) out_unpair(os, O, C, rb, 1, cb, ca, mb, hb.nm, hb, ma1, ha);
                else if (P.report_repeat_hits == 2) { for (int k = 0; k < mb; k++) { const bsl_hit &x = allb[hb.all_first + k]; out_unpair(os, O, C, rb, 1, x.read_chain, cb, mb, hb.nm, x, ma1, ha); } }   // pairs.cpp:293-297 (chain_b is cb there, not ca)
                else if (P.report_repeat_hits == 0 && O.unmap) out_unpair(os, O, C, rb, 1, 0, ca, 0, 0, hb, ma1, ha);
            }
        }
        idx++;
        if (os.size() > (1u << 20)) { fwrite(os.data(), 1, os.size(), fout); os.clear(); }
    }
    fwrite(os.data(), 1, os.size(), fout); if (fout != stdout) fclose(fout);
    fprintf(stderr, "[oracle] reads %u aligned %llu unique %llu multi %llu lookups %llu candidates %llu max_kmer_num %u\n", idx, (unsigned long long)n_al,
            (unsigned long long)n_un, (unsigned long long)n_mu, (unsigned long long)C.st.seed_lookups, (unsigned long long)C.st.candidates, C.maxk);
    orc_ctx_destroy(ctx);
    return 0;
}
#endif
