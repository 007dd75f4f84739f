/*
 * container.h -- framing of the blocked container ("CBCB" v4): header, chromosome names, varint block index.
 * The reference stream has no framing at all (src/compression.c:128-155); this is our design ("Container").
 * Plain host C++, shared by the C ABI (api.cu) and the CPU harness of the substream coder (tests/native).
 *
 *   u32 magic, version, max_read_len, read_len_header; u64 n_reads; u32 n_blocks, n_chr, block_reads, mode
 *   per chromosome: u32 name length, name bytes, zero padding to 4
 *   u32 index_bytes; per block LEB128 varints:
 *       zigzag(n_reads - previous n_reads) << 2 | chromosome changed << 1 | generation changed
 *       [chromosome ordinal]  [generation increment - 1]
 *       zigzag(second difference of base_pos)  zigzag(difference of n_edits)
 *       per substream (four or one: CBCG_BLOCK_NSUB(mode, generation)) zigzag(difference of its byte count against the
 *       previous block's same substream)
 *   payload: per block its substreams (A | B | C | D, or the one stream) back to back
 */
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include "internal.h"

static inline void put32(std::vector<uint8_t> &v, uint32_t x) { for (int i = 0; i < 4; i++) v.push_back((uint8_t)(x >> (8 * i))); }
static inline void put64(std::vector<uint8_t> &v, uint64_t x) { for (int i = 0; i < 8; i++) v.push_back((uint8_t)(x >> (8 * i))); }
static inline uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

struct IndexState { int64_t n_reads, chr, gen, base, d1, edits, sub[CBCG_N_SUB]; };
static inline IndexState index_state(uint32_t block_reads) { IndexState s; memset(&s, 0, sizeof s); s.n_reads = (int64_t)block_reads; return s; }
static inline void put_varint(std::vector<uint8_t> &v, uint64_t x) {
    do { uint8_t c = (uint8_t)(x & 0x7f); x >>= 7; if (x) c |= 0x80; v.push_back(c); } while (x);
}
static inline uint64_t zz(int64_t v) { return ((uint64_t)v << 1) ^ (uint64_t)(v >> 63); }
static inline int64_t unzz(uint64_t v) { return (int64_t)(v >> 1) ^ -(int64_t)(v & 1); }
static inline void index_put(std::vector<uint8_t> &out, IndexState &st, const BlockDesc &b, uint32_t mode) {
    const uint32_t n_sub = CBCG_BLOCK_NSUB(mode, b.gen);
    const bool chr_ch = (int64_t)b.chr != st.chr, gen_ch = (int64_t)b.gen != st.gen;
    put_varint(out, (zz((int64_t)b.n_reads - st.n_reads) << 2) | (chr_ch ? 2u : 0u) | (gen_ch ? 1u : 0u));
    if (chr_ch) { put_varint(out, b.chr); st.base = 0; st.d1 = 0; }
    if (gen_ch) put_varint(out, (uint64_t)((int64_t)b.gen - st.gen - 1));
    const int64_t d1 = (int64_t)b.base_pos - st.base;
    put_varint(out, zz(d1 - st.d1));
    put_varint(out, zz((int64_t)b.n_edits - st.edits));
    for (uint32_t k = 0; k < n_sub; k++) { put_varint(out, zz((int64_t)b.sub_bytes[k] - st.sub[k])); st.sub[k] = b.sub_bytes[k]; }
    st.n_reads = b.n_reads; st.chr = b.chr; st.gen = b.gen; st.base = b.base_pos; st.d1 = d1; st.edits = b.n_edits;
}
static inline bool get_varint(const uint8_t *p, uint64_t end, uint64_t &o, uint64_t &v) {
    uint64_t r = 0; int sh = 0;
    for (;;) {
        if (o >= end || sh > 63) return false;
        const uint8_t c = p[o++];
        r |= (uint64_t)(c & 0x7f) << sh; sh += 7;
        if (!(c & 0x80)) break;
    }
    v = r; return true;
}
static inline bool index_get(const uint8_t *p, uint64_t end, uint64_t &o, IndexState &st, BlockDesc &b, uint32_t mode) {
    uint64_t v;
    if (!get_varint(p, end, o, v)) return false;
    st.n_reads += unzz(v >> 2);
    if (v & 2) { uint64_t c; if (!get_varint(p, end, o, c)) return false; st.chr = (int64_t)c; st.base = 0; st.d1 = 0; }
    if (v & 1) { uint64_t gi; if (!get_varint(p, end, o, gi) || gi > 255) return false; st.gen += (int64_t)gi + 1; }
    if (!get_varint(p, end, o, v)) return false;
    st.d1 += unzz(v); st.base += st.d1;
    if (!get_varint(p, end, o, v)) return false;
    st.edits += unzz(v);
    int64_t total = 0;
    const uint32_t n_sub = CBCG_BLOCK_NSUB(mode, st.gen);
    for (uint32_t k = 0; k < n_sub; k++) {
        if (!get_varint(p, end, o, v)) return false;
        st.sub[k] += unzz(v);
        if (st.sub[k] < 0 || st.sub[k] > 0x3fffffffll) return false;
        total += st.sub[k];
    }
    const int64_t lim = 0xffffffffll;
    if (st.n_reads < 0 || st.n_reads > lim || st.chr < 0 || st.chr > lim || st.gen < 0 || st.gen > 255 || st.base < 0 || st.base > lim ||
        st.edits < 0 || st.edits > lim || total > lim) return false;
    memset(&b, 0, sizeof b);
    b.n_reads = (uint32_t)st.n_reads; b.chr = (uint32_t)st.chr; b.gen = (uint32_t)st.gen; b.base_pos = (uint32_t)st.base;
    b.n_edits = (uint32_t)st.edits; b.payload_bytes = (uint32_t)total;
    for (uint32_t k = 0; k < n_sub; k++) b.sub_bytes[k] = (uint32_t)st.sub[k];
    return true;
}

/* header + names + index of the blocks hb[0 .. nb) */
static inline void container_head(std::vector<uint8_t> &h, uint32_t max_len, uint32_t L, uint64_t n_reads, uint64_t nb, const std::vector<std::string> &names,
                                  uint32_t block_reads, uint32_t mode, const BlockDesc *hb) {
    h.clear();
    put32(h, CBCG_MAGIC); put32(h, CBCG_VERSION); put32(h, max_len); put32(h, L);
    put64(h, n_reads); put32(h, (uint32_t)nb); put32(h, (uint32_t)names.size()); put32(h, block_reads); put32(h, mode);
    for (const std::string &s : names) {
        put32(h, (uint32_t)s.size());
        h.insert(h.end(), s.begin(), s.end());
        for (size_t q = s.size(); q & 3; q++) h.push_back(0);
    }
    std::vector<uint8_t> ix;
    IndexState st = index_state(block_reads);
    for (uint64_t k = 0; k < nb; k++) index_put(ix, st, hb[k], mode);
    put32(h, (uint32_t)ix.size());
    h.insert(h.end(), ix.begin(), ix.end());
}
