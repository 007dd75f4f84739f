// Host unit test of cbc_b200/csrc/ac_core.h: the closed-form renormalisation must equal the
// reference's bit-at-a-time loop (src/Arithmetic_stream.c:313-344 encoder, :431-453 decoder),
// restated here bit-serially, over random and adversarial symbol sequences.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "ac_core.h"

struct Serial { uint32_t l = 0, u = CBCG_AC_TOP, t = 0; int scale3 = 0; std::vector<int> bits; size_t rp = 0; };
static int cond(const Serial &a, int *e3) {
    uint32_t ml = a.l >> 25, mu = a.u >> 25; *e3 = 0;
    if (ml == mu) return 1;
    *e3 = ((a.l >> 24) == 1u && (a.u >> 24) == 2u);
    return 0;
}
static void enc(Serial &a, uint32_t lo, uint32_t hi, uint32_t n) {
    uint64_t range = (uint64_t)a.u - a.l + 1;
    a.u = a.l + (uint32_t)((range * hi) / n) - 1; a.l = a.l + (uint32_t)((range * lo) / n);
    int e3, e12 = cond(a, &e3);
    while (e12 || e3) {
        if (e12) { uint32_t msb = a.l >> 25; a.bits.push_back(msb); a.l = (a.l & CBCG_AC_LOWMASK) << 1; a.u = ((a.u & CBCG_AC_LOWMASK) << 1) + 1;
                   while (a.scale3 > 0) { a.bits.push_back(!msb); a.scale3--; } }
        else { a.scale3++; a.u = (((a.u << 1) & CBCG_AC_LOWMASK) | CBCG_AC_MSB) + 1; a.l = (a.l << 1) & CBCG_AC_LOWMASK; }
        e12 = cond(a, &e3);
    }
}
static int rb(Serial &a, const std::vector<int> &s) { int v = a.rp < s.size() ? s[a.rp] : 0; a.rp++; return v; }
static void dec(Serial &a, uint32_t lo, uint32_t hi, uint32_t n, const std::vector<int> &s) {
    uint64_t range = (uint64_t)a.u - a.l + 1;
    a.u = a.l + (uint32_t)((range * hi) / n) - 1; a.l = a.l + (uint32_t)((range * lo) / n);
    int e3, e12 = cond(a, &e3);
    while (e12 || e3) {
        if (e12) { a.l = (a.l & CBCG_AC_LOWMASK) << 1; a.u = ((a.u & CBCG_AC_LOWMASK) << 1) + 1; a.t = ((a.t & CBCG_AC_LOWMASK) << 1) + rb(a, s); }
        else { a.l = (a.l << 1) & CBCG_AC_LOWMASK; a.u = (((a.u << 1) & CBCG_AC_LOWMASK) | CBCG_AC_MSB) + 1; a.t = (((a.t & CBCG_AC_LOWMASK) << 1) ^ CBCG_AC_MSB) + rb(a, s); }
        e12 = cond(a, &e3);
    }
}

static uint64_t rs = 88172645463325252ull;
static uint32_t rnd() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); }

int main() {
    for (int trial = 0; trial < 300; trial++) {
        Serial S; AcInterval a{0, CBCG_AC_TOP}; int scale3 = 0; std::vector<int> bits;
        std::vector<uint32_t> los, his, ns;
        int nsym = 2000;
        for (int i = 0; i < nsym; i++) {
            uint32_t n = (trial % 3 == 0) ? (1u << 20) - 1 - rnd() % 100 : 2 + rnd() % ((1u << 20) - 3);
            uint32_t lo, hi;
            int kind = rnd() % 8;
            if (kind == 0) { lo = 0; hi = n; }                         // probability 1
            else if (kind == 1) { lo = rnd() % n; hi = lo + 1; }       // tiny interval
            else if (kind == 2) { lo = n / 2 - 1 > n ? 0 : (n / 2 ? n / 2 - 1 : 0); hi = lo + 1 + (rnd() % 2); if (hi > n) hi = n; }  // straddles the middle: E3 runs
            else { lo = rnd() % n; hi = lo + 1 + rnd() % (n - lo); }
            los.push_back(lo); his.push_back(hi); ns.push_back(n);
            enc(S, lo, hi, n);
            ac_narrow(a, lo, hi, n);
            uint32_t k, b, m; AcInterval nx; ac_renorm_shape(a, k, b, m, nx);
            if (k) { int b0 = (b >> (k - 1)) & 1; bits.push_back(b0); while (scale3 > 0) { bits.push_back(!b0); scale3--; }
                     for (int j = (int)k - 2; j >= 0; j--) bits.push_back((b >> j) & 1); }
            scale3 += (int)m; a = nx;
            if (a.l != S.l || a.u != S.u || scale3 != S.scale3 || bits.size() != S.bits.size()) { printf("enc mismatch trial %d sym %d\n", trial, i); return 1; }
        }
        if (bits != S.bits) { printf("bit mismatch trial %d\n", trial); return 1; }
        // flush like encoder_last_step so the decoder has real bits
        std::vector<int> stream = bits; { int msb = a.l >> 25; stream.push_back(msb); while (scale3 > 0) { stream.push_back(!msb); scale3--; } for (int j = 24; j >= 0; j--) stream.push_back((a.l >> j) & 1); }
        Serial D; for (int i = 0; i < 26; i++) D.t = (D.t << 1) | rb(D, stream);
        AcInterval d{0, CBCG_AC_TOP}; uint32_t t = D.t; size_t rp = 26;
        for (int i = 0; i < nsym; i++) {
            uint32_t target = ac_target(d, t, ns[i]);
            { uint64_t range = (uint64_t)D.u - D.l + 1, gap = (uint64_t)D.t - D.l + 1; uint32_t ts = (uint32_t)((gap * ns[i] - 1) / range);
              if (ts != target || !(los[i] <= target && target < his[i])) { printf("target mismatch trial %d sym %d\n", trial, i); return 1; } }
            dec(D, los[i], his[i], ns[i], stream);
            ac_narrow(d, los[i], his[i], ns[i]);
            uint32_t k, b, m; AcInterval nx; ac_renorm_shape(d, k, b, m, nx);
            uint32_t in = 0; for (uint32_t j = 0; j < k + m; j++) { int v = rp < stream.size() ? stream[rp] : 0; rp++; in = (in << 1) | v; }
            t = ac_tag_shift(t, k, m, in); d = nx;
            if (d.l != D.l || d.u != D.u || t != D.t || rp != D.rp) { printf("dec mismatch trial %d sym %d (k=%u m=%u)\n", trial, i, k, m); return 1; }
        }
    }
    printf("ac_core ok\n");
    return 0;
}
