// rule.cpp — conversion-rule encoder: the product's counterpart of Param::SetAlign
// (/root/reference/param.cpp:163-263).  Produces the five byte->code tables every kernel uses.
#include <cctype>
#include <cstdio>
#include <cstring>

#include "common.cuh"

static int nt_index(int c) {
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; }
    return -1;
}

// Returns 0, or BSL_EINVAL with the reference's message in err.
int bsl_make_rule(const char from0, const char *to, RuleTables *rt, char *err, size_t errlen) {
    static const char kNt[4] = {'A', 'C', 'G', 'T'};
    const char from = (char)toupper((unsigned char)from0);
    const int fi = nt_index(from);
    if (fi < 0) { snprintf(err, errlen, "invalid -M, ref base %c not in A/C/G/T", from0); return BSL_EINVAL; }   // param.cpp:172-174
    bool is_to[4] = {false, false, false, false}; bool has_del = false; int n_to = 0;
    for (const char *p = to; p && *p; ++p) {
        const char t = (char)toupper((unsigned char)*p);
        if (t == from) { snprintf(err, errlen, "invalid -M, read base %c should not be equal to ref base %c", *p, from); return BSL_EINVAL; }   // :187-189
        if (t == '-') { if (!has_del) { has_del = true; n_to++; } continue; }
        const int ti = nt_index(t);
        if (ti < 0) { snprintf(err, errlen, "invalid -M, read base %c not in A/C/G/T/-", *p); return BSL_EINVAL; }           // :190-192
        if (!is_to[ti]) { is_to[ti] = true; n_to++; }                                                                     // duplicates ignored, :193-196
    }
    memset(rt, 0, sizeof *rt);
    for (int i = 0; i < 4; i++) { rt->reg[(u8)kNt[i]] = 3; rt->reg[(u8)tolower(kNt[i])] = 3; }                           // param.cpp:130-139
    memcpy(rt->conv, rt->reg, 256); memcpy(rt->rconv, rt->reg, 256);
    for (int i = 0; i < 4; i++) if (is_to[i]) {                                                                          // param.cpp:202-215
        rt->conv[(u8)kNt[i]] = rt->conv[(u8)tolower(kNt[i])] = 1;
        rt->rconv[(u8)kNt[3 - i]] = rt->rconv[(u8)tolower(kNt[3 - i])] = 1;
    }
    if (has_del) rt->conv[(u8)'-'] = 1;
    // 2-bit codes: from-base 01; lone non-'-' convert-to base 11; the others 00,10(,11) in ACGT order (param.cpp:216-233)
    rt->single = (n_to == 1 && !has_del) ? 1 : 0;
    int code[4] = {-1, -1, -1, -1};
    code[fi] = 1;
    if (rt->single) for (int i = 0; i < 4; i++) if (is_to[i]) code[i] = 3;
    const int spare[3] = {0, 2, 3};
    for (int i = 0, j = 0; i < 4; i++) if (code[i] < 0) code[i] = spare[j++];
    for (int i = 0; i < 4; i++) {                                                                                        // param.cpp:238-260
        rt->code[(u8)kNt[i]] = rt->code[(u8)tolower(kNt[i])] = (u8)code[i];
        rt->rcode[(u8)kNt[i]] = rt->rcode[(u8)tolower(kNt[i])] = (u8)code[3 - i];
        rt->letter[code[i]] = kNt[i]; rt->letter[code[i] + 4] = (char)tolower(kNt[i]);
    }
    return 0;
}
