/*
 * k2_coder.cu -- K2: adaptive context models + arithmetic coder, one warp per block (hot path part 2).
 *
 * Replaces, bit for bit:
 *   stream_model.c   update_model :31-51, send_value_to_as :53-76, read_value_from_as :78-117
 *   Arithmetic_stream.c  arithmetic_encoder_step :274-345, encoder_last_step :348-364,
 *                        arithmetic_get_symbol_range :373-381, arithmetic_decoder_step :389-454,
 *                        the MSB-first bit packer :155-194
 *   read_compression.c   compress_read :15-44, compress_pos(_alpha) :75-159, compress_flag :50-70,
 *                        compress_match/snps/indels/var/chars :164-260, the emission half of
 *                        compress_edits :557-600, compute_delta_to_first_snp :703-718
 *   read_decompression.c decompress_read :59-86 and the decoding half of reconstruct_read :339-458
 *   id_compression.c     compress_rname / decompress_rname :39-94 (single-block mode)
 *   qv_codebook.c        compress_int / decompress_int :14-95 (stream header, single-block mode)
 * with the initial model states of sam_models.c (:56-411, :562-586, :611-620, :734-770).
 *
 * B200 design. Each block of position-ordered reads is an independent coder instance run by ONE
 * WARP; K2_WARPS blocks share a CTA. All 32 lanes carry the coder interval redundantly (uniform
 * registers, no divergence); the lanes cooperate on what is a serial scan in the reference:
 *   - cumulative counts: strided partial sums + __reduce_add_sync (encode), 32-wide inclusive
 *     shuffle scan + ballot (decode search);
 *   - the 65 536-symbol FLAG model and the 255-symbol length models are kept SPARSE (all-ones
 *     initial state is closed under the reference's halve-and-increment rescale), so a FLAG step is
 *     one pass over <= 128 touched values in shared memory instead of a 65 536-entry scan;
 *   - the first 32 slots of the growing POS alphabet live one per lane in registers;
 *   - var rows (65 535 contexts x L counts) are created on first touch in a per-block arena in HBM,
 *     found through a 32-wide probed hash table (one coalesced 256 B load per lookup);
 *   - small dense models (snps, indels, chars, match, length byte 0) live in shared memory;
 *   - the SNP-site memory snpInRef[] (300 MB global array in the reference) is a 1024-bit ring held
 *     one word per lane: reads are position-sorted, so only [pos-1, pos+L) is ever consulted;
 *   - renormalisation is closed-form (ac_core.h): no bit-at-a-time loop.
 * The coder is latency-bound integer work, not HBM-bound: throughput comes from the number of
 * resident warps (blocks), see DESIGN.md.
 */
#include "common.cuh"
#include "internal.h"
#include "ac_core.h"

#define K2_WARPS    4u
#define K2_THREADS  (K2_WARPS * 32u)
#define FLAG_CAP    128u
#define VAR_DIRECT_MIN_EDITS 32768u       /* blocks with more edits index var rows directly by context */

enum { MODE_ENC = 0, MODE_DEC = 1, MODE_LIST = 2 };

/* ------------------------------------------------------------------------------------------------
 * per-block workspace layout in HBM (u32 units unless noted) */
struct WsLayout {
    uint64_t pos_cnt, pos_val;     /* pos_cap each */
    uint64_t pos_alpha;            /* 4 x 257 */
    uint64_t var_hash;             /* u64 x hash_cap, or init bitmap (2048 u32) in direct mode */
    uint64_t var_rows;             /* rows x Lp */
    uint64_t codebook;             /* legacy: 4 x 257 */
    uint64_t rname;                /* legacy: 256 x 257 */
    uint64_t total;                /* bytes */
    uint32_t pos_cap, hash_cap, rows_cap, Lp, direct;
};

__host__ __device__ inline uint32_t pow2_ceil(uint32_t x) { uint32_t p = 32; while (p < x) p <<= 1; return p; }

__host__ __device__ inline WsLayout ws_layout(uint32_t L, uint64_t n_reads, uint64_t n_edits, int legacy, int primed) {
    WsLayout w;
    w.Lp = (L + 1u + 31u) & ~31u;
    w.pos_cap = (uint32_t)(n_reads + 34u) + (primed ? CBCG_SNAP_POS_MAX : 0u);
    w.direct = (!primed && (legacy || n_edits >= VAR_DIRECT_MIN_EDITS)) ? 1u : 0u;
    w.rows_cap = w.direct ? CBCG_VAR_CONTEXTS : (uint32_t)n_edits;
    w.hash_cap = w.direct ? 1024u : pow2_ceil((uint32_t)(2u * n_edits + 2u));   /* u64 slots */
    uint64_t o = 0;                                                            /* in bytes, 16-aligned pieces */
    w.pos_cnt = o;   o += ((uint64_t)w.pos_cap * 4u + 15u) & ~15ull;
    w.pos_val = o;   o += ((uint64_t)w.pos_cap * 4u + 15u) & ~15ull;
    w.pos_alpha = o; o += 4u * 257u * 4u + 16u;
    w.var_hash = o;  o += (uint64_t)w.hash_cap * 8u;
    w.var_rows = o;  o += (uint64_t)w.rows_cap * w.Lp * 4u;
    w.codebook = o;  if (legacy) o += 4u * 257u * 4u + 16u;
    w.rname = o;     if (legacy) o += 256u * 257u * 4u + 16u;
    w.total = (o + 255u) & ~255ull;
    return w;
}

__host__ __device__ inline uint64_t payload_cap_bytes(uint64_t n_reads, uint64_t n_edits, int legacy) {
    /* <= 16 symbols per read + 2 per edit (+ header / names), <= 20 bits each (count >= 1, n <= 2^20) */
    uint64_t syms = 16u * n_reads + 2u * n_edits + (legacy ? 136u + 4096u : 0u) + 8u;
    return ((syms * 20u) / 8u + 64u + 15u) & ~15ull;
}
__host__ __device__ inline uint64_t symlist_cap(uint64_t n_reads, uint64_t n_edits, int legacy) {
    return 12u * n_reads + 2u * n_edits + (legacy ? 136u + 2048u : 0u) + 8u;
}

uint64_t coder_ws_bytes_bound(uint32_t L, uint64_t n_reads, uint64_t n_edits, uint64_t n_blocks, int legacy, int primed) {
    if (legacy) return ws_layout(L, n_reads, n_edits, 1, 0).total + 256;
    const uint64_t Lp = (L + 1u + 31u) & ~31u;
    /* per block: pos 2 x (n+34) x 4 (+32), pos_alpha 4128, hash <= max(32, 4 n_edits + 4) x 8, rows n_edits x Lp x 4;
       a block in direct mode replaces hash + rows by 8 KB + 65535 rows: bounded by its own n_edits >= 32768 rows. */
    uint64_t b = n_blocks * (2u * (34u * 4u + 16u) + 4u * 257u * 4u + 16u + 32u * 8u + 8192u + 512u);
    b += n_reads * 8u + n_edits * 32u + n_edits * Lp * 4u;
    if (primed) b += n_blocks * (uint64_t)CBCG_SNAP_POS_MAX * 8u;
    return 2u * b + 4096u;                                   /* direct blocks use <= 2 x their hashed size */
}
uint64_t coder_payload_bound(uint64_t n_reads, uint64_t n_edits, uint64_t n_blocks, int legacy) {
    return payload_cap_bytes(n_reads, n_edits, legacy) + n_blocks * 96u;
}

/* ------------------------------------------------------------------------------------------------
 * shared-memory models of one warp. Dense model = counts[card] followed by n at [card]. */
struct WarpModels {
    uint32_t snps[256];            /* card L, n at [L]                (initialize_stream_model_snps :243) */
    uint32_t indels[256];          /* card L                          (:277) */
    uint32_t rlen0[256];           /* card 255: length byte 0         (initialize_stream_model_id(.,4,255) :583) */
    uint32_t chars[6][8];          /* card 5                          (:350-411) */
    uint32_t match[4][4];          /* card 2                          (:204) */
    uint32_t same_ref[4];          /* card 2                          (:617) */
    uint32_t rlenk[3][2];          /* length bytes 1..3, always symbol 0: (count[0], n) */
    uint32_t flag_key[FLAG_CAP];   /* FLAG model (:96-130), sparse: touched values, ascending */
    uint32_t flag_cnt[FLAG_CAP];
    uint32_t flag_used, flag_n;
    uint16_t cumdel[256];          /* decoder: cumulative deletion offsets of the current read */
};

/* ------------------------------------------------------------------------------------------------
 * generation snapshot (gen_mode 1): the state every block of the next generation starts from. */
struct SnapLayout { uint64_t small, pos_hdr, pos_val, pos_cnt, pos_alpha, bitmap, flag_prev, flag_acc, var, total; uint32_t Lp; };
__host__ __device__ inline SnapLayout snap_layout(uint32_t L) {
    SnapLayout s; uint64_t o = 0;
    s.Lp = (L + 1u + 31u) & ~31u;
    s.small = o;     o += (sizeof(WarpModels) + 15u) & ~15ull;
    s.pos_hdr = o;   o += 16u;                                         /* card, n */
    s.pos_val = o;   o += (uint64_t)(CBCG_SNAP_POS_MAX + 32u) * 4u;
    s.pos_cnt = o;   o += (uint64_t)(CBCG_SNAP_POS_MAX + 32u) * 4u;
    s.pos_alpha = o; o += 4u * 257u * 4u + 16u;
    s.bitmap = o;    o += 2048u * 4u;
    s.flag_prev = o; o += 65536u * 4u;                                 /* merge scratch: dense FLAG counts */
    s.flag_acc = o;  o += 65536u * 4u;
    s.var = o;       o += (uint64_t)CBCG_VAR_CONTEXTS * s.Lp * 4u;
    s.total = (o + 255u) & ~255ull;
    return s;
}
uint64_t snapshot_bytes(uint32_t L) { return snap_layout(L).total; }
uint64_t fin_stride_bytes(void) { return (sizeof(WarpModels) + 15u) & ~15ull; }

__device__ __forceinline__ uint64_t fin_stride_dev() { return (sizeof(WarpModels) + 15u) & ~15ull; }

struct SnapView {
    const uint32_t *small, *pos_hdr, *pos_val, *pos_cnt, *pos_alpha, *bitmap, *var;
    uint32_t Lp;
    __host__ __device__ SnapView() {}
    __host__ __device__ SnapView(const uint8_t *base, uint32_t L) {
        const SnapLayout l = snap_layout(L);
        small = (const uint32_t *)(base + l.small); pos_hdr = (const uint32_t *)(base + l.pos_hdr);
        pos_val = (const uint32_t *)(base + l.pos_val); pos_cnt = (const uint32_t *)(base + l.pos_cnt);
        pos_alpha = (const uint32_t *)(base + l.pos_alpha); bitmap = (const uint32_t *)(base + l.bitmap);
        var = (const uint32_t *)(base + l.var); Lp = l.Lp;
    }
};

/* ------------------------------------------------------------------------------------------------ */
template <int MODE>
struct Coder {
    /* --- arithmetic coder state (uniform across the warp) */
    AcInterval a;
    uint32_t t;
    int32_t scale3;
    /* encoder bit packer */
    uint64_t acc; uint32_t nacc;
    uint8_t *out; uint32_t out_pos, out_cap;
    /* decoder bit reader */
    const uint8_t *in; uint32_t in_len; uint64_t in_bit;
    /* symbol list */
    cbcg_symbol *list; uint32_t list_n, list_cap;

    uint32_t lane;
    int err;
    uint32_t n_symbols;

    /* --- models */
    WarpModels *M;
    uint32_t L, Lp;
    /* pos: slots 0..31 one per lane, the rest in HBM */
    uint32_t pos_rv, pos_rc, pos_card, pos_n, pos_cap;
    uint32_t *pos_gval, *pos_gcnt;
    uint32_t *pos_alpha; bool pa_init;
    /* var */
    uint64_t *var_hash; uint32_t hash_mask; uint32_t *var_rows; uint32_t n_rows, rows_cap; bool var_direct;
    uint32_t *var_bitmap;
    /* legacy-only */
    uint32_t *codebook, *rname;
    /* primed blocks */
    bool primed, lean; SnapView snap;
    /* SNP-site ring: word (p >> 5) & 31 lives in lane; covers [ring_base, ring_base + 1024) */
    uint32_t ring; uint32_t ring_word;     /* ring_word = ring_base >> 5 */

    /* ============================================================ bit I/O */
    __device__ __forceinline__ void put_bits(uint32_t v, uint32_t k) {       /* k <= 32 */
        if (k == 0) return;
        acc = (acc << k) | (uint64_t)v;
        nacc += k;
        if (nacc >= 32u) {
            uint32_t w = (uint32_t)(acc >> (nacc - 32u));
            if (out_pos + 4u <= out_cap) { if (lane == 0) *reinterpret_cast<uint32_t *>(out + out_pos) = __byte_perm(w, 0u, 0x0123); }
            else err = CBCG_ERR_CAPACITY;
            out_pos += 4u;
            nacc -= 32u;
            acc &= (1ull << nacc) - 1ull;
        }
    }
    __device__ __forceinline__ void put_run(uint32_t bit, uint32_t count) {
        const uint32_t pat = bit ? 0xffffffffu : 0u;
        while (count >= 32u) { put_bits(pat, 32u); count -= 32u; }
        if (count) put_bits(pat >> (32u - count), count);
    }
    /* stream_finish_byte (:189-194): the byte in progress always goes out, even an empty one */
    __device__ __forceinline__ void finish_bits(bool always_last = true) {
        uint32_t full = nacc >> 3, rem = nacc & 7u;
        for (uint32_t i = 0; i < full; i++) {
            uint32_t byte = (uint32_t)(acc >> (nacc - 8u * (i + 1u))) & 0xffu;
            if (out_pos < out_cap) { if (lane == 0) out[out_pos] = (uint8_t)byte; } else err = CBCG_ERR_CAPACITY;
            out_pos++;
        }
        if (rem || always_last) {
            uint32_t last = rem ? (((uint32_t)acc & ((1u << rem) - 1u)) << (8u - rem)) : 0u;
            if (out_pos < out_cap) { if (lane == 0) out[out_pos] = (uint8_t)last; } else err = CBCG_ERR_CAPACITY;
            out_pos++;
        }
        nacc = 0; acc = 0;
    }
    /* next k bits of the input, first bit most significant; zeros past the end (the reference's
       zero-filled buffer, src/Arithmetic_stream.c:30,117) */
    __device__ __forceinline__ uint32_t get_bits(uint32_t k) {               /* k <= 32 */
        if (k == 0) return 0u;
        const uint64_t byte0 = in_bit >> 3;
        uint64_t w = 0;
#pragma unroll
        for (uint32_t i = 0; i < 5; i++) {
            uint64_t p = byte0 + i;
            uint64_t v = (p < in_len) ? (uint64_t)in[p] : 0ull;
            w = (w << 8) | v;
        }
        const uint32_t sh = 40u - (uint32_t)(in_bit & 7u) - k;
        in_bit += k;
        return (uint32_t)((w >> sh) & ((1ull << k) - 1ull));
    }

    /* ============================================================ arithmetic coder */
    __device__ __forceinline__ void ac_init() {
        a.l = 0; a.u = CBCG_AC_TOP; scale3 = 0; t = 0; acc = 0; nacc = 0; out_pos = 0; in_bit = 0;
        if (MODE == MODE_DEC) t = get_bits(CBCG_AC_BITS);                    /* :262 */
    }
    __device__ __forceinline__ void ac_encode(uint32_t lo, uint32_t cnt, uint32_t n) {
        if (cnt == 0u || n == 0u) { err = CBCG_ERR_INPUT; return; }          /* reference: assert :71 / :293 */
        ac_narrow(a, lo, lo + cnt, n);
        uint32_t k, bits, m; AcInterval nx;
        ac_renorm_shape(a, k, bits, m, nx);
        if (k) {
            const uint32_t b0 = (bits >> (k - 1u)) & 1u;
            put_bits(b0, 1u);
            if (scale3 > 0) { put_run(b0 ^ 1u, (uint32_t)scale3); scale3 = 0; }
            if (k > 1u) put_bits(bits & ((1u << (k - 1u)) - 1u), k - 1u);
        }
        scale3 += (int32_t)m;
        a = nx;
    }
    __device__ __forceinline__ void ac_decode_step(uint32_t lo, uint32_t cnt, uint32_t n) {
        ac_narrow(a, lo, lo + cnt, n);
        uint32_t k, bits, m; AcInterval nx;
        ac_renorm_shape(a, k, bits, m, nx);
        t = ac_tag_shift(t, k, m, get_bits(k + m));
        a = nx;
    }
    /* encoder_last_step (:348-364) */
    __device__ __forceinline__ void ac_flush() {
        const uint32_t msb = a.l >> (CBCG_AC_BITS - 1u);
        put_bits(msb, 1u);
        if (scale3 > 0) { put_run(msb ^ 1u, (uint32_t)scale3); scale3 = 0; }
        put_bits(a.l & CBCG_AC_LOWMASK, CBCG_AC_BITS - 1u);
        finish_bits();
    }

    /* Blocked containers: after renormalisation l < 2^25 <= u, so the value 2^25 -- "1", the pending E3 bits as
       "0", zeros ever after -- lies in [l, u]; the decoder reads zeros past the end of a block. */
    __device__ __forceinline__ void ac_flush_short() {
        put_bits(1u, 1u);
        if (scale3 > 0) { put_run(0u, (uint32_t)scale3); scale3 = 0; }
        finish_bits(false);
    }

    /* one coder step given the symbol's interval; decode: caller found (lo, cnt) from the target */
    __device__ __forceinline__ void code_interval(uint32_t lo, uint32_t cnt, uint32_t n) {
        if (MODE == MODE_ENC) ac_encode(lo, cnt, n); else ac_decode_step(lo, cnt, n);
        n_symbols++;
    }

    /* ============================================================ dense models (counts[card], n at [card]) */
    __device__ __forceinline__ void dense_update(uint32_t *m, uint32_t card, uint32_t step, uint32_t x) {
        uint32_t n = m[card] + step;
        __syncwarp();
        if (lane == 0) { m[x] += step; m[card] = n; }
        __syncwarp();
        if (n >= CBCG_RESCALE) {                                             /* update_model :38-49 */
            uint32_t s = 0;
            for (uint32_t i = lane; i < card; i += 32u) { uint32_t c = (m[i] >> 1) + 1u; m[i] = c; s += c; }
            s = warp_sum(s);
            __syncwarp();
            if (lane == 0) m[card] = s;
            __syncwarp();
        }
    }
    __device__ __forceinline__ uint32_t sym_dense(uint32_t *m, uint32_t card, uint32_t step, uint32_t x) {
        if (err) return 0u;
        const uint32_t n = m[card];
        uint32_t lo, cnt;
        if (MODE == MODE_ENC) {
            if (x >= card) { err = CBCG_ERR_INPUT; return 0u; }              /* reference: assert :62 */
            uint32_t s = 0;
            for (uint32_t i = lane; i < x; i += 32u) s += m[i];
            lo = warp_sum(s); cnt = m[x];
        } else {
            const uint32_t target = ac_target(a, t, n);
            uint32_t carry = 0; bool found = false;
            lo = 0; cnt = 0; x = 0;
            for (uint32_t base = 0; base < card; base += 32u) {
                const uint32_t i = base + lane;
                const uint32_t c = (i < card) ? m[i] : 0u;
                const uint32_t incl = warp_incl_scan(c) + carry;
                const uint32_t hit = __ballot_sync(FULL_MASK, i < card && incl > target);
                if (hit) {
                    const uint32_t h = (uint32_t)__ffs(hit) - 1u;
                    cnt = __shfl_sync(FULL_MASK, c, h);
                    lo = __shfl_sync(FULL_MASK, incl, h) - cnt;
                    x = base + h; found = true;
                    break;
                }
                carry = __shfl_sync(FULL_MASK, incl, 31);
            }
            if (!found) { err = CBCG_ERR_CORRUPT; return 0u; }
        }
        code_interval(lo, cnt, n);
        if (err) return 0u;
        dense_update(m, card, step, x);
        return x;
    }

    __device__ __forceinline__ void dense_init_ones(uint32_t *m, uint32_t card) {
        for (uint32_t i = lane; i < card; i += 32u) m[i] = 1u;
        if (lane == 0) m[card] = card;
    }

    /* ============================================================ length bytes 1..3 (always symbol 0) */
    __device__ __forceinline__ uint32_t sym_rlenk(uint32_t k, uint32_t x) {
        if (err) return 0u;
        uint32_t c0 = M->rlenk[k][0], n = M->rlenk[k][1];
        if (MODE == MODE_ENC) { if (x != 0u) { err = CBCG_ERR_INPUT; return 0u; } }
        else { if (ac_target(a, t, n) >= c0) { err = CBCG_ERR_CORRUPT; return 0u; } }
        code_interval(0u, c0, n);
        c0 += 10u; n += 10u;
        if (n >= CBCG_RESCALE) { c0 = (c0 >> 1) + 1u; n = c0 + 254u; }
        __syncwarp();
        if (lane == 0) { M->rlenk[k][0] = c0; M->rlenk[k][1] = n; }
        __syncwarp();
        return 0u;
    }

    /* ============================================================ FLAG (65 536 symbols, sparse) */
    __device__ __forceinline__ uint32_t sym_flag(uint32_t x) {
        if (err) return 0u;
        const uint32_t used = M->flag_used, n = M->flag_n;
        uint32_t lo, cnt; int found_idx = -1;
        if (MODE == MODE_ENC) {
            if (x > 0xffffu) { err = CBCG_ERR_INPUT; return 0u; }
            uint32_t extra = 0; cnt = 1u;
            for (uint32_t base = 0; base < used; base += 32u) {
                const uint32_t i = base + lane;
                const uint32_t k = (i < used) ? M->flag_key[i] : 0xffffffffu;
                const uint32_t c = (i < used) ? M->flag_cnt[i] : 1u;
                if (k < x) extra += c - 1u;
                const uint32_t hit = __ballot_sync(FULL_MASK, k == x);
                if (hit) { const uint32_t h = (uint32_t)__ffs(hit) - 1u; cnt = __shfl_sync(FULL_MASK, c, h); found_idx = (int)(base + h); }
            }
            lo = x + warp_sum(extra);
        } else {
            const uint32_t target = ac_target(a, t, n);
            uint32_t carry = 0; bool done = false;
            lo = target; cnt = 1u; x = 0;
            for (uint32_t base = 0; base < used && !done; base += 32u) {
                const uint32_t i = base + lane;
                const bool valid = i < used;
                const uint32_t k = valid ? M->flag_key[i] : 0xffffffffu;
                const uint32_t c = valid ? M->flag_cnt[i] : 1u;
                const uint32_t e = c - 1u;
                const uint32_t incl = warp_incl_scan(e) + carry;
                const uint32_t E = incl - e;                       /* extras of all touched values below this one */
                const uint32_t A = k + E;                          /* cumulative count at the start of value k */
                const uint32_t below = __ballot_sync(FULL_MASK, valid && (A + c <= target));
                const uint32_t nb = (uint32_t)__popc(below);
                if (nb == 32u) { carry = __shfl_sync(FULL_MASK, incl, 31); continue; }
                const uint32_t Ec = __shfl_sync(FULL_MASK, E, nb);
                const uint32_t Ac = __shfl_sync(FULL_MASK, A, nb);
                const uint32_t cc = __shfl_sync(FULL_MASK, c, nb);
                const uint32_t kc = __shfl_sync(FULL_MASK, k, nb);
                const bool vc = (base + nb) < used;
                if (vc && Ac <= target) { x = kc; lo = Ac; cnt = cc; found_idx = (int)(base + nb); }
                else { x = target - Ec; lo = target; cnt = 1u; }
                carry = Ec; done = true;
            }
            if (!done) x = target - carry;                         /* every touched value lies below */
            if (x > 0xffffu) { err = CBCG_ERR_CORRUPT; return 0u; }
        }
        code_interval(lo, cnt, n);
        if (err) return 0u;
        /* update_model with step 8 */
        __syncwarp();
        uint32_t nused = used;
        if (found_idx >= 0) { if (lane == 0) M->flag_cnt[found_idx] += 8u; }
        else {
            if (used >= FLAG_CAP) { err = CBCG_ERR_LIMIT; return 0u; }
            uint32_t p = 0;                                        /* insertion point: touched values below x */
            for (uint32_t base = 0; base < used; base += 32u) {
                const uint32_t i = base + lane;
                p += (uint32_t)__popc(__ballot_sync(FULL_MASK, i < used && M->flag_key[i] < x));
            }
            if (used > p) {
                for (int base = (int)((used - 1u) & ~31u); base >= 0; base -= 32) {
                    const uint32_t i = (uint32_t)base + lane;
                    const bool mv = (i >= p && i < used);
                    uint32_t k = 0, c = 0;
                    if (mv) { k = M->flag_key[i]; c = M->flag_cnt[i]; }
                    __syncwarp();
                    if (mv) { M->flag_key[i + 1u] = k; M->flag_cnt[i + 1u] = c; }
                    __syncwarp();
                    if ((uint32_t)base <= p) break;
                }
            }
            if (lane == 0) { M->flag_key[p] = x; M->flag_cnt[p] = 1u + 8u; M->flag_used = used + 1u; }
            nused = used + 1u;
        }
        uint32_t nn = n + 8u;
        __syncwarp();
        if (nn >= CBCG_RESCALE) {
            uint32_t s = 0;
            for (uint32_t i = lane; i < nused; i += 32u) { uint32_t c = (M->flag_cnt[i] >> 1) + 1u; M->flag_cnt[i] = c; s += c; }
            nn = warp_sum(s) + (65536u - nused);
        }
        if (lane == 0) M->flag_n = nn;
        __syncwarp();
        return x;
    }

    /* ============================================================ POS (growing alphabet) */
    __device__ __forceinline__ void pos_load(uint32_t base, uint32_t &v, uint32_t &c) {
        const uint32_t i = base + lane;
        if (base == 0u) { v = pos_rv; c = (i < pos_card) ? pos_rc : 0u; }
        else if (i < pos_card) { v = pos_gval[i]; c = pos_gcnt[i]; }
        else { v = 0u; c = 0u; }
    }
    __device__ __forceinline__ void pos_update(uint32_t slot) {                 /* update_model, step 10 */
        if (slot < 32u) { if (lane == slot) pos_rc += 10u; }
        else if (lane == 0) pos_gcnt[slot] += 10u;
        pos_n += 10u;
        __syncwarp();
        if (pos_n >= CBCG_RESCALE) {
            uint32_t s = 0;
            if (lane < pos_card) { pos_rc = (pos_rc >> 1) + 1u; s += pos_rc; }
            for (uint32_t i = 32u + lane; i < pos_card; i += 32u) { uint32_t c = (pos_gcnt[i] >> 1) + 1u; pos_gcnt[i] = c; s += c; }
            pos_n = warp_sum(s);
            __syncwarp();
        }
    }
    __device__ __forceinline__ void pos_append(uint32_t x) {                    /* new symbol, count 0, then updated (:147-153) */
        const uint32_t slot = pos_card;
        if (slot >= pos_cap) { err = CBCG_ERR_INTERNAL; return; }
        if (slot < 32u) { if (lane == slot) { pos_rv = x; pos_rc = 0u; } }
        else if (lane == 0) { pos_gval[slot] = x; pos_gcnt[slot] = 0u; }
        pos_card = slot + 1u;
        __syncwarp();
        pos_update(slot);
    }
    __device__ __forceinline__ void pa_ensure() {
        if (!pa_init) {
            if (primed) { for (uint32_t i = lane; i < 4u * 257u; i += 32u) pos_alpha[i] = snap.pos_alpha[i]; }
            else for (uint32_t k = 0; k < 4u; k++) dense_init_ones(pos_alpha + k * 257u, 256u);
            pa_init = true;
            __syncwarp();
        }
    }
    /* compress_pos / decompress_pos: x = pos - prevPos + 1 */
    __device__ __forceinline__ uint32_t sym_pos(uint32_t x) {
        if (err) return 0u;
        uint32_t lo = 0, cnt = 0, slot = 0;
        if (MODE == MODE_ENC) {
            bool found = false;
            for (uint32_t base = 0; base < pos_card; base += 32u) {
                uint32_t v, c; pos_load(base, v, c);
                const uint32_t i = base + lane;
                const uint32_t hit = __ballot_sync(FULL_MASK, i >= 1u && i < pos_card && v == x);
                if (hit) {
                    const uint32_t h = (uint32_t)__ffs(hit) - 1u;
                    lo += warp_sum(lane < h ? c : 0u);
                    cnt = __shfl_sync(FULL_MASK, c, h);
                    slot = base + h; found = true;
                    break;
                }
                lo += warp_sum(c);
            }
            if (!found) { slot = 0; lo = 0; cnt = __shfl_sync(FULL_MASK, pos_rc, 0); }
        } else {
            const uint32_t target = ac_target(a, t, pos_n);
            uint32_t carry = 0; bool found = false;
            for (uint32_t base = 0; base < pos_card; base += 32u) {
                uint32_t v, c; pos_load(base, v, c);
                const uint32_t incl = warp_incl_scan(c) + carry;
                const uint32_t hit = __ballot_sync(FULL_MASK, (base + lane) < pos_card && incl > target);
                if (hit) {
                    const uint32_t h = (uint32_t)__ffs(hit) - 1u;
                    cnt = __shfl_sync(FULL_MASK, c, h);
                    lo = __shfl_sync(FULL_MASK, incl, h) - cnt;
                    slot = base + h; x = __shfl_sync(FULL_MASK, v, h); found = true;
                    break;
                }
                carry = __shfl_sync(FULL_MASK, incl, 31);
            }
            if (!found) { err = CBCG_ERR_CORRUPT; return 0u; }
        }
        code_interval(lo, cnt, pos_n);
        if (err) return 0u;
        pos_update(slot);
        if (slot != 0u) return x;
        /* escape: the value goes out as 4 bytes, MSB first (compress_pos_alpha :75-108) */
        pa_ensure();
        uint32_t y = 0;
        for (uint32_t k = 0; k < 4u; k++) {
            uint32_t byte = sym_dense(pos_alpha + k * 257u, 256u, 10u, (x >> (24u - 8u * k)) & 0xffu);
            y |= byte << (24u - 8u * k);
        }
        if (MODE == MODE_DEC) x = y;
        if (err) return 0u;
        pos_append(x);
        return x;
    }

    /* ============================================================ var rows */
    __device__ __forceinline__ uint32_t *var_row(uint32_t ctx) {
        if (ctx >= CBCG_VAR_CONTEXTS) { err = (MODE == MODE_ENC) ? CBCG_ERR_INPUT : CBCG_ERR_CORRUPT; return nullptr; }
        if (var_direct) {
            uint32_t *row = var_rows + (uint64_t)ctx * Lp;
            const uint32_t w = var_bitmap[ctx >> 5];
            if (!((w >> (ctx & 31u)) & 1u)) {
                dense_init_ones(row, L);
                __syncwarp();
                if (lane == 0) var_bitmap[ctx >> 5] = w | (1u << (ctx & 31u));
                __syncwarp();
            }
            return row;
        }
        const uint32_t key = ctx + 1u;
        uint32_t h = (ctx * 0x9E3779B1u) >> 7;
        for (uint32_t probes = 0; probes <= hash_mask; probes += 32u, h += 32u) {
            const uint32_t idx = (h + lane) & hash_mask;
            const uint64_t s = var_hash[idx];
            const uint32_t k = (uint32_t)(s >> 32);
            const uint32_t mm = __ballot_sync(FULL_MASK, k == key);
            const uint32_t ee = __ballot_sync(FULL_MASK, k == 0u);
            if (mm && (!ee || __ffs(mm) < __ffs(ee))) {
                const uint32_t r = __shfl_sync(FULL_MASK, (uint32_t)s, __ffs(mm) - 1);
                return var_rows + (uint64_t)r * Lp;
            }
            if (ee) {
                const uint32_t el = (uint32_t)__ffs(ee) - 1u;
                if (n_rows >= rows_cap) { err = CBCG_ERR_INTERNAL; return nullptr; }
                const uint32_t r = n_rows++;
                uint32_t *row = var_rows + (uint64_t)r * Lp;
                if (lane == el) var_hash[idx] = ((uint64_t)key << 32) | r;
                if (primed && ((snap.bitmap[ctx >> 5] >> (ctx & 31u)) & 1u)) {       /* copy on first touch */
                    const uint32_t *src = snap.var + (uint64_t)ctx * Lp;
                    for (uint32_t i = lane; i <= L; i += 32u) row[i] = src[i];
                } else dense_init_ones(row, L);
                __syncwarp();
                return row;
            }
        }
        err = CBCG_ERR_INTERNAL;
        return nullptr;
    }
    __device__ __forceinline__ uint32_t sym_var(uint32_t ctx, uint32_t x) {
        if (err) return 0u;
        uint32_t *row = var_row(ctx);
        if (!row) return 0u;
        return sym_dense(row, L, 10u, x);
    }

    /* ============================================================ SNP-site ring (snpInRef) */
    __device__ __forceinline__ void ring_reset() { ring = 0u; ring_word = 0u; }
    __device__ __forceinline__ void ring_advance(uint32_t pos) {               /* window must start at or below pos - 1 */
        const uint32_t nw = (pos - 1u) >> 5;
        if (nw > ring_word) {
            const uint32_t adv = nw - ring_word;
            if (adv >= 32u || ((lane - ring_word) & 31u) < adv) ring = 0u;     /* words that rotate out and back in */
            ring_word = nw;
        }
    }
    __device__ __forceinline__ void ring_set(uint32_t p) {                      /* snpInRef[p] = 1 */
        if ((p >> 5) - ring_word < 32u && lane == ((p >> 5) & 31u)) ring |= 1u << (p & 31u);
    }
    /* compute_delta_to_first_snp (:703-718): distance from position s to the first marked site in
       [s, e), else `none` */
    __device__ __forceinline__ uint32_t ring_first(uint32_t s, uint32_t e, uint32_t none) {
        uint32_t best = 0xffffffffu;
        if (e > s) {
            const uint32_t sw = s >> 5;
            const uint32_t w = sw + ((lane - sw) & 31u);                        /* this lane's word at or after s */
            if (w - ring_word < 32u && (w << 5) < e) {
                uint32_t bitsw = ring;
                if (w == sw) bitsw &= 0xffffffffu << (s & 31u);
                if (((w + 1u) << 5) > e) bitsw &= (e & 31u) ? ((1u << (e & 31u)) - 1u) : 0xffffffffu;
                if (bitsw) best = (w << 5) + (uint32_t)__ffs(bitsw) - 1u;
            }
        }
        best = warp_min(best);
        return best == 0xffffffffu ? none : best - s;
    }

    /* ============================================================ symbol dispatch */
    __device__ __forceinline__ void list_put(uint32_t stream, uint32_t ctx, uint32_t x) {
        if (list_n < list_cap) { if (lane == 0) { list[list_n].key = CBCG_SYM_KEY(stream, ctx); list[list_n].value = x; } }
        else err = CBCG_ERR_CAPACITY;
        list_n++;
        n_symbols++;
    }
    __device__ __forceinline__ uint32_t sym(uint32_t stream, uint32_t ctx, uint32_t x) {
        if (MODE == MODE_LIST) { list_put(stream, ctx, x); return x; }
        switch (stream) {
            case CBCG_S_CODEBOOK:  return sym_dense(codebook + ctx * 257u, 256u, 1u, x);
            case CBCG_S_SAME_REF:  return sym_dense(M->same_ref, 2u, 10u, x);
            case CBCG_S_RNAME:     return sym_dense(rname + ctx * 257u, 256u, 10u, x);
            case CBCG_S_RLENGTH:   return ctx == 0u ? sym_dense(M->rlen0, 255u, 10u, x) : sym_rlenk(ctx - 1u, x);
            case CBCG_S_FLAG:      return sym_flag(x);
            case CBCG_S_MATCH:     return sym_dense(M->match[ctx], 2u, 1u, x);
            case CBCG_S_SNPS:      return sym_dense(M->snps, L, 10u, x);
            case CBCG_S_INDELS:    return sym_dense(M->indels, L, 16u, x);
            case CBCG_S_VAR:       return sym_var(ctx, x);
            case CBCG_S_CHARS:     return sym_dense(M->chars[ctx], 5u, 8u, x);
            default: err = CBCG_ERR_INTERNAL; return 0u;
        }
    }
    __device__ __forceinline__ uint32_t sym_posx(uint32_t x) {
        if (MODE == MODE_LIST) { list_put(CBCG_S_POS_X, 0u, x); return x; }
        return sym_pos(x);
    }

    /* ============================================================ model initial states */
    __device__ __forceinline__ void init_L_models() {          /* the models whose alphabet is the header read length */
        dense_init_ones(M->snps, L);
        dense_init_ones(M->indels, L);
        __syncwarp();
    }
    __device__ __forceinline__ void init_from_snapshot() {
        uint32_t *dst = reinterpret_cast<uint32_t *>(M);
        for (uint32_t i = lane; i < (uint32_t)(sizeof(WarpModels) / 4u); i += 32u) dst[i] = snap.small[i];
        pos_card = snap.pos_hdr[0]; pos_n = snap.pos_hdr[1];
        pos_rv = (lane < pos_card) ? snap.pos_val[lane] : 0u;
        pos_rc = (lane < pos_card) ? snap.pos_cnt[lane] : 0u;
        for (uint32_t i = 32u + lane; i < pos_card; i += 32u) { pos_gval[i] = snap.pos_val[i]; pos_gcnt[i] = snap.pos_cnt[i]; }
        pa_init = false;
        n_rows = 0;
        for (uint32_t i = lane; i <= hash_mask; i += 32u) var_hash[i] = 0ull;
        __syncwarp();
    }
    __device__ __forceinline__ void init_models(bool legacy) {
        if (MODE == MODE_LIST) return;
        dense_init_ones(M->rlen0, 255u);
        if (lane < 6u) {                                        /* initialize_stream_model_chars :350-411 */
            uint32_t n = 0;
            for (uint32_t i = 0; i < 4u; i++) { uint32_t c = (i == lane) ? 0u : 8u; M->chars[lane][i] = c; n += c; }
            M->chars[lane][4] = 1u; n += 1u;
            if (lane < 4u) {
                const uint32_t f0 = (lane == 0u || lane == 3u) ? 1u : 0u, f1 = (lane == 0u || lane == 3u) ? 2u : 3u;
                M->chars[lane][f0] += 8u; M->chars[lane][f1] += 8u; n += 16u;
            }
            M->chars[lane][5] = n;
        }
        if (lane < 4u) { M->match[lane][0] = 1u; M->match[lane][1] = 1u; M->match[lane][2] = 2u; }
        if (lane == 0) {
            M->same_ref[0] = 1u; M->same_ref[1] = 1u; M->same_ref[2] = 2u;
            for (uint32_t k = 0; k < 3u; k++) { M->rlenk[k][0] = 1u; M->rlenk[k][1] = 255u; }
            M->flag_used = 0u; M->flag_n = 65536u;
        }
        pos_rv = 0u; pos_rc = (lane == 0u) ? 1u : 0u; pos_card = 1u; pos_n = 1u;   /* escape only (:132-162) */
        pa_init = false;
        n_rows = 0;
        if (var_direct) { for (uint32_t i = lane; i < 2048u; i += 32u) var_bitmap[i] = 0u; }
        else { for (uint32_t i = lane; i <= hash_mask; i += 32u) var_hash[i] = 0ull; }
        if (legacy) {
            for (uint32_t k = 0; k < 4u; k++) dense_init_ones(codebook + k * 257u, 256u);
            for (uint32_t k = 0; k < 256u; k++) dense_init_ones(rname + k * 257u, 256u);
        }
        __syncwarp();
    }
};

/* ------------------------------------------------------------------------------------------------
 * read-level driver: compress_read / decompress_read and the emission / decoding half of
 * compress_edits / reconstruct_read. State kept by the reference in statics: prev_pos
 * (src/read_compression.c:115), prev_m (:167). */
template <int MODE>
struct ReadState { uint32_t prev_pos, prev_m; };

template <int MODE>
__device__ __forceinline__ void code_read(Coder<MODE> &C, ReadState<MODE> &st, cbcg_read_rec &rec,
                                          const uint16_t *e_in, uint16_t *e_out, uint32_t &n_edits_out,
                                          uint32_t edits_room, const uint8_t *ref, uint64_t ref_len) {
    const uint32_t lane = C.lane;
    /* length: byte 0 carries it, bytes 1..3 are always 0 (:29-33) */
    uint32_t len = C.sym(CBCG_S_RLENGTH, 0u, rec.len & 0xffu);
    if (!C.lean || MODE == MODE_LIST) for (uint32_t k = 1; k < 4u; k++) len |= C.sym(CBCG_S_RLENGTH, k, 0u) << (8u * k);
    if (C.err) return;
    if (MODE != MODE_DEC) len = rec.len;
    /* position (:113-159) */
    uint32_t x;
    if (MODE == MODE_DEC) { x = C.sym_posx(0u); if (!C.err && x == 0u) C.err = CBCG_ERR_CORRUPT; }
    else {
        if (rec.pos == 0u || rec.len == 0u || rec.len > CBCG_MAX_READ_LEN) { C.err = CBCG_ERR_INPUT; return; }
        if (rec.pos < st.prev_pos || rec.pos - st.prev_pos + 1u > CBCG_MAX_POS_X) { C.err = CBCG_ERR_INPUT; return; }
        x = C.sym_posx(rec.pos - st.prev_pos + 1u);
    }
    if (C.err) return;
    const uint32_t pos = (MODE == MODE_DEC) ? st.prev_pos + x - 1u : rec.pos;
    if (MODE == MODE_DEC && (pos == 0u || len == 0u || len > CBCG_MAX_READ_LEN)) { C.err = CBCG_ERR_CORRUPT; return; }
    st.prev_pos = pos;
    C.ring_advance(pos);
    const uint32_t flag = C.sym(CBCG_S_FLAG, 0u, rec.flag);
    const uint32_t strand = (flag >> 4) & 1u;                                   /* :57-60 */
    const uint32_t match = C.sym(CBCG_S_MATCH, ((uint32_t)(x == 1u) << 1) | st.prev_m, rec.match);
    if (C.err) return;
    st.prev_m = match;
    if (MODE == MODE_DEC) {
        rec.pos = pos; rec.flag = (uint16_t)flag; rec.len = (uint16_t)len; rec.match = (uint8_t)match;
        rec.n_snps = rec.n_dels = rec.n_ins = 0;
    }
    n_edits_out = 0;
    if (match) return;

    uint32_t ns = rec.n_snps, nd = rec.n_dels, ni = rec.n_ins;
    if (MODE == MODE_DEC) {
        ns = C.sym(CBCG_S_SNPS, 0u, 0u); nd = 0; ni = 0;
        if (ns == 0u) { ns = C.sym(CBCG_S_INDELS, 0u, 0u); nd = C.sym(CBCG_S_INDELS, 0u, 0u); ni = C.sym(CBCG_S_INDELS, 0u, 0u); }
        if (C.err) return;
        if (ni > len || ns > 255u || nd > 255u || ni > 255u) { C.err = CBCG_ERR_CORRUPT; return; }
        if (ns + nd + ni > edits_room) { C.err = CBCG_ERR_CAPACITY; return; }
        rec.n_snps = (uint8_t)ns; rec.n_dels = (uint8_t)nd; rec.n_ins = (uint8_t)ni;
    } else {
        if ((nd | ni) == 0u) C.sym(CBCG_S_SNPS, 0u, ns);
        else { C.sym(CBCG_S_SNPS, 0u, 0u); C.sym(CBCG_S_INDELS, 0u, ns); C.sym(CBCG_S_INDELS, 0u, nd); C.sym(CBCG_S_INDELS, 0u, ni); }
    }
    uint32_t ne = 0;
    /* deletions (:568-572) */
    uint32_t prev = 0;
    for (uint32_t k = 0; k < nd && !C.err; k++) {
        const uint32_t d_in = (MODE == MODE_DEC) ? 0u : CBCG_EDIT_DELTA(e_in[k]);
        const uint32_t d = C.sym(CBCG_S_VAR, (prev << 1) | strand, d_in);
        prev += d;
        if (MODE == MODE_DEC) {
            if (lane == 0) { e_out[ne] = CBCG_EDIT(d, 0, 0); C.M->cumdel[k] = (uint16_t)min(prev, 0xffffu); }
        }
        ne++;
    }
    if (MODE == MODE_DEC) __syncwarp();
    /* SNPs (:573-593) */
    prev = 0;
    for (uint32_t k = 0; k < ns && !C.err; k++) {
        const uint32_t ed = (MODE == MODE_DEC) ? 0u : e_in[nd + k];
        const uint32_t delta = C.ring_first(pos - 1u + prev, (prev < len) ? pos - 1u + len : pos - 1u + prev, len + 2u);
        const uint32_t p = C.sym(CBCG_S_VAR, (((delta << CBCG_BITS_DELTA) + prev) << 1) | strand, CBCG_EDIT_DELTA(ed));
        if (C.err) break;
        const uint32_t idx = prev + p;                                          /* index in the insertion-free read */
        prev += p + 1u;
        C.ring_set(pos + prev - 2u);                                            /* :589 */
        uint32_t refb;
        if (MODE == MODE_DEC) {
            uint32_t skipped = 0;                                               /* deletions at or before idx (:426-437) */
            for (uint32_t q = lane; q < nd; q += 32u) skipped += (C.M->cumdel[q] <= idx);
            skipped = warp_sum(skipped);
            const uint64_t ri = (uint64_t)pos - 1u + idx + skipped;
            refb = base_code(ri < ref_len ? (uint32_t)ref[ri] : 0u);
        } else refb = CBCG_EDIT_REFB(ed);
        const uint32_t tgt = C.sym(CBCG_S_CHARS, refb, CBCG_EDIT_TARGET(ed));
        if (MODE == MODE_DEC && lane == 0) e_out[ne] = CBCG_EDIT(p, tgt, refb);
        ne++;
    }
    /* insertions (:594-600) */
    prev = 0;
    for (uint32_t k = 0; k < ni && !C.err; k++) {
        const uint32_t ed = (MODE == MODE_DEC) ? 0u : e_in[nd + ns + k];
        const uint32_t p = C.sym(CBCG_S_VAR, (prev << 1) | strand, CBCG_EDIT_DELTA(ed));
        prev += p;
        const uint32_t tgt = C.sym(CBCG_S_CHARS, CBCG_BP_O, CBCG_EDIT_TARGET(ed));
        if (MODE == MODE_DEC && lane == 0) e_out[ne] = CBCG_EDIT(p, tgt, CBCG_BP_O);
        ne++;
    }
    n_edits_out = ne;
}

/* compress_int (src/qv_codebook.c:14-52): 4 bytes MSB first through codebook[0..3] */
template <int MODE>
__device__ __forceinline__ uint32_t code_int(Coder<MODE> &C, uint32_t v) {
    uint32_t r = 0;
    for (uint32_t k = 0; k < 4u; k++) r |= C.sym(CBCG_S_CODEBOOK, k, (v >> (24u - 8u * k)) & 0xffu) << (24u - 8u * k);
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(K2_THREADS)
k2_coder_kernel(CoderParams P) {
    __shared__ WarpModels smodels[K2_WARPS];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t bl = blockIdx.x * K2_WARPS + warp;
    if (bl >= P.n_blocks) return;
    const uint32_t b = P.block_begin + bl;
    BlockDesc &B = P.blocks[b];
    const bool legacy = P.legacy != 0;
    const bool primed = P.primed != 0 && MODE != MODE_LIST;

    Coder<MODE> C;
    C.lane = lane; C.err = 0; C.n_symbols = 0; C.M = &smodels[warp];
    C.primed = primed; C.lean = P.lean != 0;
    if (primed) C.snap = SnapView(P.snap, P.L);
    C.L = P.L; C.Lp = (P.L + 1u + 31u) & ~31u;
    /* workspace */
    const uint64_t ws_edits = (MODE == MODE_DEC && legacy) ? 0xffffffffull : B.n_edits;
    const uint64_t ws_reads = B.n_reads;
    uint32_t decode_L = P.L;
    {
        const WsLayout w = ws_layout(P.L ? P.L : 252u, ws_reads, ws_edits, legacy, primed);
        uint8_t *base = P.ws + B.ws_off;
        C.pos_gcnt = reinterpret_cast<uint32_t *>(base + w.pos_cnt);
        C.pos_gval = reinterpret_cast<uint32_t *>(base + w.pos_val);
        C.pos_alpha = reinterpret_cast<uint32_t *>(base + w.pos_alpha);
        C.var_hash = reinterpret_cast<uint64_t *>(base + w.var_hash);
        C.var_bitmap = reinterpret_cast<uint32_t *>(base + w.var_hash);
        C.var_rows = reinterpret_cast<uint32_t *>(base + w.var_rows);
        C.codebook = reinterpret_cast<uint32_t *>(base + w.codebook);
        C.rname = reinterpret_cast<uint32_t *>(base + w.rname);
        C.pos_cap = w.pos_cap; C.hash_mask = w.hash_cap - 1u; C.rows_cap = w.rows_cap; C.var_direct = w.direct != 0u;
        C.Lp = w.Lp;
    }
    C.out = P.payload + B.payload_off; C.out_cap = (uint32_t)payload_cap_bytes(B.n_reads, B.n_edits, legacy);
    C.in = P.payload + B.payload_off; C.in_len = B.payload_bytes;
    C.list = P.symbols + B.sym_off; C.list_n = 0; C.list_cap = (uint32_t)symlist_cap(B.n_reads, B.n_edits, legacy);
    if (primed) C.init_from_snapshot(); else C.init_models(legacy);
    C.ring_reset();
    if (MODE != MODE_LIST) C.ac_init();

    ReadState<MODE> st; st.prev_pos = B.base_pos; st.prev_m = 0u;
    uint32_t prev_char = 0u;
    uint32_t cur_chr = legacy ? 0xffffffffu : B.chr;
    const uint8_t *ref = nullptr; uint64_t ref_len = 0;
    if (!legacy && B.chr < P.genome.n_chr) { ref = P.genome.bases + P.genome.chr_off[B.chr]; ref_len = P.genome.chr_len[B.chr]; }

    if (legacy) {
        /* stream header: read length, 32 WELL words, LOSSLESS (src/sam_file_allocation.c:363-404,
           src/compression.c:139) */
        uint32_t Lh = code_int(C, P.L);
        for (uint32_t i = 0; i < CBCG_WELL_WORDS; i++) code_int(C, CBCG_WELL_DEBUG);
        uint32_t lossy = code_int(C, CBCG_LOSSLESS);
        if (MODE == MODE_DEC) {
            if (C.err || Lh == 0u || Lh > CBCG_MAX_READ_LEN || lossy != CBCG_LOSSLESS) C.err = C.err ? C.err : CBCG_ERR_FORMAT;
            else { C.L = Lh; decode_L = Lh; }
        }
    } else if (MODE == MODE_DEC && B.chr >= P.genome.n_chr) C.err = CBCG_ERR_NO_REFERENCE;
    if (MODE != MODE_LIST && !C.err && !primed) C.init_L_models();

    const uint64_t r0 = B.first_read;
    uint64_t e_cursor = B.edit_base;
    uint32_t n_done = 0;
    const uint32_t n_reads = B.n_reads;                 /* legacy decode: capacity, the end marker stops the loop */
    const uint64_t edits_cap_abs = B.edit_base + B.n_edits;

    for (uint32_t i = 0; !C.err; i++) {
        if (!(legacy && MODE == MODE_DEC) && i >= n_reads) break;
        const uint64_t r = r0 + i;
        __align__(16) cbcg_read_rec rec;
        uint32_t chr = cur_chr;
        if (MODE != MODE_DEC) {
            *reinterpret_cast<uint4 *>(&rec) = reinterpret_cast<const uint4 *>(P.recs)[r];
            chr = P.chr[r];
        } else { rec.pos = 0; rec.flag = 0; rec.len = 0; rec.edit_off = 0; rec.match = 0; rec.n_snps = rec.n_dels = rec.n_ins = 0; }

        /* compress_rname / decompress_rname (src/id_compression.c:39-94) */
        if (legacy) {
            bool change;
            if (MODE == MODE_DEC) {
                change = C.sym(CBCG_S_SAME_REF, 0u, 0u) != 0u;
                if (C.err) break;
                if (change) {
                    bool end = false; uint32_t ch;
                    while (!C.err && (ch = C.sym(CBCG_S_RNAME, prev_char, 0u)) != 0u) {
                        if (ch == '\n') { end = true; break; }
                        prev_char = ch;
                    }
                    if (C.err || end) break;
                    chr = cur_chr + 1u;                  /* records are taken in FASTA order (src/compression.c:91-101) */
                    if (chr >= P.genome.n_chr) { C.err = CBCG_ERR_NO_REFERENCE; break; }
                }
                if (cur_chr == 0xffffffffu && !change) { C.err = CBCG_ERR_CORRUPT; break; }
                if (i >= n_reads) { C.err = CBCG_ERR_CAPACITY; break; }
            } else {
                change = (chr != cur_chr);
                if (chr >= P.genome.n_chr) { C.err = CBCG_ERR_NO_REFERENCE; break; }
                if (change) {
                    C.sym(CBCG_S_SAME_REF, 0u, 1u);
                    const uint8_t *name = P.chr_names + (uint64_t)chr * MAX_NAME;
                    for (uint32_t q = 0; q < MAX_NAME && name[q]; q++) { C.sym(CBCG_S_RNAME, prev_char, name[q]); prev_char = name[q]; }
                    C.sym(CBCG_S_RNAME, prev_char, 0u);
                } else C.sym(CBCG_S_SAME_REF, 0u, 0u);
            }
            if (change) {                                /* src/compression.c:58-64 */
                cur_chr = chr;
                st.prev_pos = 0u;
                C.ring_reset();
                ref = P.genome.bases + P.genome.chr_off[chr]; ref_len = P.genome.chr_len[chr];
            }
        } else {
            if (MODE != MODE_DEC && chr != cur_chr) { C.err = CBCG_ERR_INTERNAL; break; }   /* blocks never span chromosomes */
            if (!C.lean || MODE == MODE_LIST) {
                const uint32_t same = C.sym(CBCG_S_SAME_REF, 0u, 0u);
                if (MODE == MODE_DEC && same != 0u) { C.err = CBCG_ERR_CORRUPT; break; }
            }
        }
        if (C.err) break;

        uint32_t ne = 0;
        const uint64_t room64 = edits_cap_abs - e_cursor;
        const uint32_t room = room64 > 0xffffffffull ? 0xffffffffu : (uint32_t)room64;
        code_read<MODE>(C, st, rec, P.edits + (MODE == MODE_DEC ? 0 : rec.edit_off), P.edits + e_cursor, ne, room, ref, ref_len);
        if (C.err) break;
        if (MODE == MODE_DEC) {
            rec.edit_off = (uint32_t)e_cursor;
            if (lane == 0) { reinterpret_cast<uint4 *>(P.recs)[r] = *reinterpret_cast<uint4 *>(&rec); P.chr[r] = cur_chr; }
            e_cursor += ne;
        }
        n_done++;
    }

    if (MODE == MODE_ENC && !C.err) {
        if (legacy) {                                    /* end-of-stream marker: name "\n" (src/compression.c:152) */
            C.sym(CBCG_S_SAME_REF, 0u, 1u);
            C.sym(CBCG_S_RNAME, prev_char, '\n'); prev_char = '\n';
            C.sym(CBCG_S_RNAME, prev_char, 0u);
        }
        if (!C.err) { if (P.short_flush && !legacy) C.ac_flush_short(); else C.ac_flush(); }
    }
    if (MODE == MODE_LIST && legacy && !C.err) {
        C.sym(CBCG_S_SAME_REF, 0u, 1u);
        C.sym(CBCG_S_RNAME, prev_char, '\n');
        C.sym(CBCG_S_RNAME, '\n', 0u);
    }
    if (C.err) dev_set_error(P.err, C.err, ((uint64_t)b << 20) | (n_done & 0xfffffu));
    if (primed && P.fin && !C.err) {                     /* final state for the generation merge */
        __syncwarp();
        uint32_t *dst = reinterpret_cast<uint32_t *>(P.fin + (uint64_t)bl * fin_stride_dev());
        const uint32_t *src = reinterpret_cast<const uint32_t *>(C.M);
        for (uint32_t i = lane; i < (uint32_t)(sizeof(WarpModels) / 4u); i += 32u) dst[i] = src[i];
        if (lane < C.pos_card) { C.pos_gval[lane] = C.pos_rv; C.pos_gcnt[lane] = C.pos_rc; }
        if (lane == 0) { B.pos_card = C.pos_card; B.n_rows = C.n_rows; B.pa_touched = C.pa_init ? 1u : 0u; }
    }
    if (lane == 0) {
        B.n_symbols = (MODE == MODE_LIST) ? C.list_n : C.n_symbols;
        if (MODE == MODE_ENC) B.payload_bytes = C.out_pos;
        if (MODE == MODE_DEC) { B.n_reads = n_done; B.n_edits = (uint32_t)(e_cursor - B.edit_base); B.gen = decode_L; }
    }
}

int launch_coder(const CoderParams &p, cudaStream_t st) {
    if (p.n_blocks == 0) return 0;
    const unsigned grid = (p.n_blocks + K2_WARPS - 1) / K2_WARPS;
    if (p.mode == MODE_ENC) k2_coder_kernel<MODE_ENC><<<grid, K2_THREADS, 0, st>>>(p);
    else if (p.mode == MODE_DEC) k2_coder_kernel<MODE_DEC><<<grid, K2_THREADS, 0, st>>>(p);
    else k2_coder_kernel<MODE_LIST><<<grid, K2_THREADS, 0, st>>>(p);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

/* ------------------------------------------------------------------------------------------------
 * plan: per-block sizes -> offsets (one CTA; blocks in chunks of 1024 with a running carry).
 * totals[0] = workspace bytes, [1] = payload bytes, [2] = symbol-list entries, [3] = reads, [4] = edits. */
#define PLAN_THREADS 1024u
__global__ void __launch_bounds__(PLAN_THREADS)
k2_plan_kernel(CoderParams P, uint32_t n_reads_total, uint64_t n_edits_total, uint64_t ws_cap, uint64_t payload_cap,
               uint64_t *totals) {
    __shared__ uint64_t wsum[5][32];
    __shared__ uint64_t carry[5];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid < 5) carry[tid] = 0;
    __syncthreads();
    const bool dec = (P.mode == MODE_DEC);
    for (uint32_t base = 0; base < P.n_blocks; base += PLAN_THREADS) {
        const uint32_t b = base + tid;
        uint64_t v[5] = { 0, 0, 0, 0, 0 };
        BlockDesc d;
        if (b < P.n_blocks) {
            d = P.blocks[b];
            if (!dec) {                                  /* edits of the block from the records' offsets */
                const uint64_t lo = P.recs[d.first_read].edit_off;
                const uint64_t end = (uint64_t)d.first_read + d.n_reads;
                const uint64_t hi = (end < n_reads_total) ? P.recs[end].edit_off : n_edits_total;
                d.n_edits = (uint32_t)(hi - lo);
                d.edit_base = lo;
                if (!P.legacy) d.base_pos = P.recs[d.first_read].pos;
            }
            const uint64_t ws_edits = (dec && P.legacy) ? 0xffffffffull : d.n_edits;
            v[0] = ws_layout(P.L ? P.L : 252u, d.n_reads, ws_edits, P.legacy, P.primed && P.mode != MODE_LIST).total;
            v[1] = dec ? d.payload_bytes : payload_cap_bytes(d.n_reads, d.n_edits, P.legacy);
            v[2] = (P.mode == MODE_LIST) ? symlist_cap(d.n_reads, d.n_edits, P.legacy) : 0;
            v[3] = d.n_reads; v[4] = d.n_edits;
        }
        uint64_t incl[5];
#pragma unroll
        for (int q = 0; q < 5; q++) {
            uint64_t x = v[q];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint64_t y = __shfl_up_sync(FULL_MASK, x, o); if (lane >= (uint32_t)o) x += y; }
            incl[q] = x;
            if (lane == 31) wsum[q][warp] = x;
        }
        __syncthreads();
        uint64_t off[5];
#pragma unroll
        for (int q = 0; q < 5; q++) {
            uint64_t wb = 0;
            for (uint32_t k = 0; k < warp; k++) wb += wsum[q][k];
            off[q] = carry[q] + wb + incl[q] - v[q];
        }
        if (b < P.n_blocks) {
            d.ws_off = off[0];
            if (!dec) d.payload_off = off[1];
            else { d.payload_off = off[1]; d.first_read = (uint32_t)off[3]; d.edit_base = off[4]; }
            d.sym_off = off[2];
            P.blocks[b] = d;
        }
        __syncthreads();
        if (tid == PLAN_THREADS - 1) {
#pragma unroll
            for (int q = 0; q < 5; q++) carry[q] = off[q] + v[q];
        }
        __syncthreads();
    }
    if (tid == 0) {
        for (int q = 0; q < 5; q++) totals[q] = carry[q];
        if (carry[0] > ws_cap || (!dec && P.mode == MODE_ENC && carry[1] > payload_cap)) dev_set_error(P.err, CBCG_ERR_INTERNAL, 0xabcdefull);
    }
}

int launch_plan(const CoderParams &p, uint32_t n_reads_total, uint64_t n_edits_total, uint64_t ws_cap,
                uint64_t payload_cap, uint64_t *totals, cudaStream_t st) {
    k2_plan_kernel<<<1, PLAN_THREADS, 0, st>>>(p, n_reads_total, n_edits_total, ws_cap, payload_cap, totals);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

/* ------------------------------------------------------------------------------------------------
 * gather: block payloads (scratch regions) -> one contiguous payload; out_off[b] = its offset. */
__global__ void __launch_bounds__(PLAN_THREADS)
k2_payload_scan_kernel(const BlockDesc *blocks, uint32_t n_blocks, uint64_t *out_off) {
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_blocks; base += PLAN_THREADS) {
        const uint32_t b = base + tid;
        const uint64_t v = (b < n_blocks) ? blocks[b].payload_bytes : 0;
        uint64_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint64_t y = __shfl_up_sync(FULL_MASK, x, o); if (lane >= (uint32_t)o) x += y; }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        uint64_t wb = 0;
        for (uint32_t k = 0; k < warp; k++) wb += wsum[k];
        const uint64_t off = carry + wb + x - v;
        if (b < n_blocks) out_off[b] = off;
        __syncthreads();
        if (tid == PLAN_THREADS - 1) carry = off + v;
        __syncthreads();
    }
    if (tid == 0) out_off[n_blocks] = carry;
}

__global__ void k2_gather_kernel(const BlockDesc *blocks, uint32_t n_blocks, const uint8_t *scratch, uint8_t *out,
                                 const uint64_t *out_off) {
    for (uint32_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const uint8_t *src = scratch + blocks[b].payload_off;
        uint8_t *dst = out + out_off[b];
        const uint32_t n = blocks[b].payload_bytes;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    }
}

int launch_gather(const BlockDesc *blocks, uint32_t n_blocks, const uint8_t *scratch, uint8_t *out,
                  uint64_t *out_off, cudaStream_t st) {
    if (n_blocks == 0) return 0;
    k2_payload_scan_kernel<<<1, PLAN_THREADS, 0, st>>>(blocks, n_blocks, out_off);
    unsigned grid = n_blocks < 148u * 8u ? n_blocks : 148u * 8u;
    k2_gather_kernel<<<grid, 128, 0, st>>>(blocks, n_blocks, scratch, out, out_off);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

/* ================================================================================================
 * Generation snapshots (gen_mode 1). S_g = S_{g-1} + sum over the blocks b of generation g of
 * (final state of b - S_{g-1}), count by count (wrapping 32-bit sums, read back as signed), clamped to
 * >= 1 (>= 0 where the snapshot held 0) and rescaled like update_model (src/stream_model.c:38-49).
 * The decoder runs the same kernels on the blocks it has decoded. `next` arrives as a byte copy of `prev`. */

__global__ void __launch_bounds__(32) snapshot_init_kernel(uint8_t *snap, uint32_t L) {
    __shared__ WarpModels M;
    const SnapLayout l = snap_layout(L);
    const uint32_t lane = threadIdx.x;
    Coder<MODE_ENC> C;                                     /* borrow the initial-state code of the block coder */
    C.lane = lane; C.M = &M; C.L = L; C.var_direct = false; C.hash_mask = 0; C.err = 0;
    __shared__ uint64_t dummy_hash[32];                     /* init_models clears the block's hash table */
    C.var_hash = dummy_hash; C.hash_mask = 31u;
    C.init_models(false);
    C.init_L_models();
    for (uint32_t i = lane; i < 256u; i += 32u) M.cumdel[i] = 0;
    __syncwarp();
    uint32_t *small = reinterpret_cast<uint32_t *>(snap + l.small);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(&M);
    for (uint32_t i = lane; i < (uint32_t)(sizeof(WarpModels) / 4u); i += 32u) small[i] = src[i];
    uint32_t *hdr = reinterpret_cast<uint32_t *>(snap + l.pos_hdr);
    uint32_t *pv = reinterpret_cast<uint32_t *>(snap + l.pos_val), *pc = reinterpret_cast<uint32_t *>(snap + l.pos_cnt);
    if (lane == 0) { hdr[0] = 1u; hdr[1] = 1u; hdr[2] = 0u; hdr[3] = 0u; pv[0] = 0u; pc[0] = 1u; }   /* escape only (sam_models.c:132-162) */
    uint32_t *pa = reinterpret_cast<uint32_t *>(snap + l.pos_alpha);
    for (uint32_t k = 0; k < 4u; k++) { for (uint32_t i = lane; i < 256u; i += 32u) pa[k * 257u + i] = 1u; if (lane == 0) pa[k * 257u + 256u] = 256u; }
    uint32_t *bm = reinterpret_cast<uint32_t *>(snap + l.bitmap);
    for (uint32_t i = lane; i < 2048u; i += 32u) bm[i] = 0u;
}

int launch_snapshot_init(uint8_t *snap, uint32_t L, cudaStream_t st) {
    snapshot_init_kernel<<<1, 32, 0, st>>>(snap, L);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

struct MergeParams {
    const BlockDesc *blocks; uint32_t block_begin, n_blocks, L;
    const uint8_t *prev; uint8_t *next; const uint8_t *fin; const uint8_t *ws;
    unsigned long long *err;
};

/* One dense model by one warp. fin(b) -> the block's counts, or NULL when the block never touched the model. */
template <class Fin>
__device__ __forceinline__ void merge_dense(uint32_t lane, const uint32_t *prev_m, uint32_t *next_m, uint32_t card,
                                            uint32_t implicit_ones, uint32_t n_blocks, Fin fin) {
    uint32_t nsum = 0;
    for (uint32_t i0 = 0; i0 < card; i0 += 32u) {
        const uint32_t i = i0 + lane;
        uint32_t p = 0, acc = 0;
        if (i < card) { p = prev_m[i]; acc = p; }
        for (uint32_t b = 0; b < n_blocks; b++) {
            const uint32_t *f = fin(b);
            if (f && i < card) acc += f[i] - p;
        }
        if (i < card) {
            int32_t v = (int32_t)acc; const int32_t fl = p == 0u ? 0 : 1;
            if (v < fl) v = fl;
            next_m[i] = (uint32_t)v; nsum += (uint32_t)v;
        }
    }
    __syncwarp();
    uint32_t n = warp_sum(nsum) + implicit_ones;
    while (n >= CBCG_RESCALE) {
        uint32_t s = 0;
        for (uint32_t i = lane; i < card; i += 32u) { const uint32_t c = (next_m[i] >> 1) + 1u; next_m[i] = c; s += c; }
        n = warp_sum(s) + implicit_ones;
    }
    if (lane == 0) next_m[card] = n;
    __syncwarp();
}

#define MERGE_SMALL_WARPS 21u
__global__ void __launch_bounds__(MERGE_SMALL_WARPS * 32u) merge_small_kernel(MergeParams P) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const SnapLayout l = snap_layout(P.L);
    const uint32_t *ps = reinterpret_cast<const uint32_t *>(P.prev + l.small);
    uint32_t *ns = reinterpret_cast<uint32_t *>(P.next + l.small);
    const uint64_t stride = fin_stride_dev();
    uint32_t off = 0, card = 0, ones = 0;
    if (warp == 0)       { off = offsetof(WarpModels, snps) / 4u;   card = P.L; }
    else if (warp == 1)  { off = offsetof(WarpModels, indels) / 4u; card = P.L; }
    else if (warp == 2)  { off = offsetof(WarpModels, rlen0) / 4u;  card = 255u; }
    else if (warp < 9)   { off = offsetof(WarpModels, chars) / 4u + (warp - 3u) * 8u; card = 5u; }
    else if (warp < 13)  { off = offsetof(WarpModels, match) / 4u + (warp - 9u) * 4u; card = 2u; }
    else if (warp == 13) { off = offsetof(WarpModels, same_ref) / 4u; card = 2u; }
    else if (warp < 17)  { off = offsetof(WarpModels, rlenk) / 4u + (warp - 14u) * 2u; card = 1u; ones = 254u; }
    if (warp < 17) {
        const uint8_t *fin = P.fin;
        merge_dense(lane, ps + off, ns + off, card, ones, P.n_blocks,
                    [=](uint32_t b) { return reinterpret_cast<const uint32_t *>(fin + (uint64_t)b * stride) + off; });
    } else {                                              /* pos_alpha[k]: lives in each block's workspace, if instantiated */
        const uint32_t k = warp - 17u;
        const uint32_t *pa_prev = reinterpret_cast<const uint32_t *>(P.prev + l.pos_alpha) + k * 257u;
        uint32_t *pa_next = reinterpret_cast<uint32_t *>(P.next + l.pos_alpha) + k * 257u;
        const BlockDesc *blocks = P.blocks + P.block_begin; const uint8_t *ws = P.ws; const uint32_t L = P.L;
        merge_dense(lane, pa_prev, pa_next, 256u, 0u, P.n_blocks, [=](uint32_t b) -> const uint32_t * {
            const BlockDesc &B = blocks[b];
            if (!B.pa_touched) return nullptr;
            const WsLayout w = ws_layout(L, B.n_reads, B.n_edits, 0, 1);
            return reinterpret_cast<const uint32_t *>(ws + B.ws_off + w.pos_alpha) + k * 257u;
        });
    }
}

/* FLAG: dense 65 536-entry accumulation in global scratch, then back to the sorted sparse form. */
__global__ void __launch_bounds__(1024) merge_flag_kernel(MergeParams P) {
    __shared__ uint32_t red[32];
    __shared__ uint32_t scan[1024];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const SnapLayout l = snap_layout(P.L);
    const WarpModels *pm = reinterpret_cast<const WarpModels *>(P.prev + l.small);
    WarpModels *nm = reinterpret_cast<WarpModels *>(P.next + l.small);
    uint32_t *dprev = reinterpret_cast<uint32_t *>(P.next + l.flag_prev), *dacc = reinterpret_cast<uint32_t *>(P.next + l.flag_acc);
    for (uint32_t i = tid; i < 65536u; i += 1024u) { dprev[i] = 1u; dacc[i] = 1u; }
    __syncthreads();
    if (tid < pm->flag_used) { const uint32_t k = pm->flag_key[tid], c = pm->flag_cnt[tid]; dprev[k] = c; dacc[k] = c; }
    __syncthreads();
    const uint64_t stride = fin_stride_dev();
    for (uint32_t b = 0; b < P.n_blocks; b++) {
        const WarpModels *fm = reinterpret_cast<const WarpModels *>(P.fin + (uint64_t)b * stride);
        if (tid < fm->flag_used) { const uint32_t k = fm->flag_key[tid] & 0xffffu; dacc[k] += fm->flag_cnt[tid] - dprev[k]; }
        __syncthreads();
    }
    /* clamp, total, rescale */
    uint32_t n;
    {
        uint32_t s = 0;
        for (uint32_t i = tid; i < 65536u; i += 1024u) { int32_t v = (int32_t)dacc[i]; if (v < 1) v = 1; dacc[i] = (uint32_t)v; s += (uint32_t)v; }
        s = warp_sum(s); if (lane == 0) red[warp] = s; __syncthreads();
        n = 0; for (uint32_t k = 0; k < 32u; k++) n += red[k];
        __syncthreads();
    }
    while (n >= CBCG_RESCALE) {
        uint32_t s = 0;
        for (uint32_t i = tid; i < 65536u; i += 1024u) { const uint32_t c = (dacc[i] >> 1) + 1u; dacc[i] = c; s += c; }
        s = warp_sum(s); if (lane == 0) red[warp] = s; __syncthreads();
        n = 0; for (uint32_t k = 0; k < 32u; k++) n += red[k];
        __syncthreads();
    }
    /* ordered compaction of the values whose count is not 1: thread t owns values [64 t, 64 t + 64) */
    uint32_t mine = 0;
    for (uint32_t i = 0; i < 64u; i++) mine += dacc[tid * 64u + i] != 1u;
    scan[tid] = mine; __syncthreads();
    for (uint32_t o = 1; o < 1024u; o <<= 1) { uint32_t v = tid >= o ? scan[tid - o] : 0u; __syncthreads(); scan[tid] += v; __syncthreads(); }
    const uint32_t total = scan[1023];
    uint32_t at = scan[tid] - mine;
    if (total > FLAG_CAP) { if (tid == 0) dev_set_error(P.err, CBCG_ERR_LIMIT, total); }
    else for (uint32_t i = 0; i < 64u; i++) { const uint32_t c = dacc[tid * 64u + i]; if (c != 1u) { nm->flag_key[at] = tid * 64u + i; nm->flag_cnt[at] = c; at++; } }
    if (tid == 0) { nm->flag_used = total > FLAG_CAP ? 0u : total; nm->flag_n = n; }
}

/* POS: slots below the snapshot's alphabet size are the same value in every block; new values are
 * appended in block order, then order of appearance (one warp, blocks in sequence). */
__global__ void __launch_bounds__(32) merge_pos_kernel(MergeParams P) {
    const uint32_t lane = threadIdx.x;
    const SnapLayout l = snap_layout(P.L);
    const uint32_t *phdr = reinterpret_cast<const uint32_t *>(P.prev + l.pos_hdr);
    const uint32_t *pcnt = reinterpret_cast<const uint32_t *>(P.prev + l.pos_cnt);
    uint32_t *nhdr = reinterpret_cast<uint32_t *>(P.next + l.pos_hdr);
    uint32_t *nval = reinterpret_cast<uint32_t *>(P.next + l.pos_val), *ncnt = reinterpret_cast<uint32_t *>(P.next + l.pos_cnt);
    const BlockDesc *blocks = P.blocks + P.block_begin;
    const uint32_t pc = phdr[0];
    uint32_t an = pc;
    for (uint32_t s0 = 0; s0 < pc; s0 += 32u) {
        const uint32_t s = s0 + lane;
        if (s < pc) {
            const uint32_t p = pcnt[s]; uint32_t acc = p;
            for (uint32_t b = 0; b < P.n_blocks; b++) {
                const BlockDesc &B = blocks[b];
                const WsLayout w = ws_layout(P.L, B.n_reads, B.n_edits, 0, 1);
                acc += reinterpret_cast<const uint32_t *>(P.ws + B.ws_off + w.pos_cnt)[s] - p;
            }
            ncnt[s] = acc;
        }
    }
    __syncwarp();
    for (uint32_t b = 0; b < P.n_blocks; b++) {
        const BlockDesc &B = blocks[b];
        const WsLayout w = ws_layout(P.L, B.n_reads, B.n_edits, 0, 1);
        const uint32_t *bval = reinterpret_cast<const uint32_t *>(P.ws + B.ws_off + w.pos_val);
        const uint32_t *bcnt = reinterpret_cast<const uint32_t *>(P.ws + B.ws_off + w.pos_cnt);
        for (uint32_t s = pc; s < B.pos_card; s++) {
            const uint32_t x = bval[s], c = bcnt[s];
            int found = -1;
            for (uint32_t q0 = pc; q0 < an && found < 0; q0 += 32u) {
                const uint32_t q = q0 + lane;
                const uint32_t hit = __ballot_sync(FULL_MASK, q < an && nval[q] == x);
                if (hit) found = (int)(q0 + (uint32_t)__ffs(hit) - 1u);
            }
            if (found >= 0) { if (lane == 0) ncnt[found] += c; }
            else if (an < CBCG_SNAP_POS_MAX) { if (lane == 0) { nval[an] = x; ncnt[an] = c; } an++; }
            __syncwarp();
        }
    }
    uint32_t s = 0;
    for (uint32_t i = lane; i < an; i += 32u) { int32_t v = (int32_t)ncnt[i]; if (v < 1) v = 1; ncnt[i] = (uint32_t)v; s += (uint32_t)v; }
    uint32_t n = warp_sum(s);
    while (n >= CBCG_RESCALE) {
        s = 0;
        for (uint32_t i = lane; i < an; i += 32u) { const uint32_t c = (ncnt[i] >> 1) + 1u; ncnt[i] = c; s += c; }
        n = warp_sum(s);
    }
    if (lane == 0) { nhdr[0] = an; nhdr[1] = n; }
}

/* var rows, phase 1: rows new to the snapshot are created (all ones) by whichever block flips their bit. */
__global__ void __launch_bounds__(128) merge_var_mark_kernel(MergeParams P) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t b = blockIdx.x * 4u + warp;
    if (b >= P.n_blocks) return;
    const SnapLayout l = snap_layout(P.L);
    uint32_t *bm = reinterpret_cast<uint32_t *>(P.next + l.bitmap);
    uint32_t *var = reinterpret_cast<uint32_t *>(P.next + l.var);
    const BlockDesc &B = P.blocks[P.block_begin + b];
    const WsLayout w = ws_layout(P.L, B.n_reads, B.n_edits, 0, 1);
    const uint64_t *hash = reinterpret_cast<const uint64_t *>(P.ws + B.ws_off + w.var_hash);
    for (uint32_t h0 = 0; h0 < w.hash_cap; h0 += 32u) {
        const uint64_t sl = hash[h0 + lane];
        const uint32_t key = (uint32_t)(sl >> 32);
        bool won = false; uint32_t ctx = 0;
        if (key) { ctx = key - 1u; const uint32_t bit = 1u << (ctx & 31u); won = !(atomicOr(&bm[ctx >> 5], bit) & bit); }
        uint32_t wm = __ballot_sync(FULL_MASK, won);
        while (wm) {
            const uint32_t src = (uint32_t)__ffs(wm) - 1u; wm &= wm - 1u;
            const uint32_t c = __shfl_sync(FULL_MASK, ctx, src);
            uint32_t *row = var + (uint64_t)c * l.Lp;
            for (uint32_t i = lane; i < P.L; i += 32u) row[i] = 1u;
            if (lane == 0) row[P.L] = P.L;
        }
    }
}
/* phase 2: add every block's row deltas. */
__global__ void __launch_bounds__(128) merge_var_add_kernel(MergeParams P) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t b = blockIdx.x * 4u + warp;
    if (b >= P.n_blocks) return;
    const SnapLayout l = snap_layout(P.L);
    const uint32_t *pbm = reinterpret_cast<const uint32_t *>(P.prev + l.bitmap);
    const uint32_t *pvar = reinterpret_cast<const uint32_t *>(P.prev + l.var);
    uint32_t *var = reinterpret_cast<uint32_t *>(P.next + l.var);
    const BlockDesc &B = P.blocks[P.block_begin + b];
    const WsLayout w = ws_layout(P.L, B.n_reads, B.n_edits, 0, 1);
    const uint64_t *hash = reinterpret_cast<const uint64_t *>(P.ws + B.ws_off + w.var_hash);
    const uint32_t *rows = reinterpret_cast<const uint32_t *>(P.ws + B.ws_off + w.var_rows);
    for (uint32_t h0 = 0; h0 < w.hash_cap; h0 += 32u) {
        const uint64_t sl = hash[h0 + lane];
        uint32_t km = __ballot_sync(FULL_MASK, (uint32_t)(sl >> 32) != 0u);
        while (km) {
            const uint32_t src = (uint32_t)__ffs(km) - 1u; km &= km - 1u;
            const uint64_t e = __shfl_sync(FULL_MASK, sl, src);
            const uint32_t ctx = (uint32_t)(e >> 32) - 1u, r = (uint32_t)e;
            const bool in_prev = (pbm[ctx >> 5] >> (ctx & 31u)) & 1u;
            const uint32_t *row = rows + (uint64_t)r * w.Lp, *prow = pvar + (uint64_t)ctx * l.Lp;
            uint32_t *nrow = var + (uint64_t)ctx * l.Lp;
            for (uint32_t i = lane; i < P.L; i += 32u) {
                const uint32_t d = row[i] - (in_prev ? prow[i] : 1u);
                if (d) atomicAdd(&nrow[i], d);
            }
        }
    }
}
/* phase 3: clamp, total, rescale every row of the new snapshot (idempotent on rows no block touched). */
__global__ void __launch_bounds__(256) merge_var_finish_kernel(MergeParams P) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t ctx = blockIdx.x * 8u + warp;
    if (ctx >= CBCG_VAR_CONTEXTS) return;
    const SnapLayout l = snap_layout(P.L);
    const uint32_t *bm = reinterpret_cast<const uint32_t *>(P.next + l.bitmap);
    if (!((bm[ctx >> 5] >> (ctx & 31u)) & 1u)) return;
    uint32_t *row = reinterpret_cast<uint32_t *>(P.next + l.var) + (uint64_t)ctx * l.Lp;
    uint32_t s = 0;
    for (uint32_t i = lane; i < P.L; i += 32u) { int32_t v = (int32_t)row[i]; if (v < 1) v = 1; row[i] = (uint32_t)v; s += (uint32_t)v; }
    uint32_t n = warp_sum(s);
    while (n >= CBCG_RESCALE) {
        s = 0;
        for (uint32_t i = lane; i < P.L; i += 32u) { const uint32_t c = (row[i] >> 1) + 1u; row[i] = c; s += c; }
        n = warp_sum(s);
    }
    if (lane == 0) row[P.L] = n;
}

int launch_merge(const BlockDesc *blocks, uint32_t block_begin, uint32_t n_blocks, uint32_t L, const uint8_t *prev,
                 uint8_t *next, const uint8_t *fin, const uint8_t *ws, unsigned long long *err, cudaStream_t st) {
    if (cudaMemcpyAsync(next, prev, snapshot_bytes(L), cudaMemcpyDeviceToDevice, st) != cudaSuccess) return -1;
    MergeParams P = { blocks, block_begin, n_blocks, L, prev, next, fin, ws, err };
    merge_small_kernel<<<1, MERGE_SMALL_WARPS * 32u, 0, st>>>(P);
    merge_flag_kernel<<<1, 1024, 0, st>>>(P);
    merge_pos_kernel<<<1, 32, 0, st>>>(P);
    merge_var_mark_kernel<<<(n_blocks + 3u) / 4u, 128, 0, st>>>(P);
    merge_var_add_kernel<<<(n_blocks + 3u) / 4u, 128, 0, st>>>(P);
    merge_var_finish_kernel<<<(CBCG_VAR_CONTEXTS + 7u) / 8u, 256, 0, st>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
