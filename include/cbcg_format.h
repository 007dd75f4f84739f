/*
 * cbcg_format.h -- bitstream-contract constants of the cbc aligned-read coding path.
 *
 * One header shared by the CUDA kernels, the C host code and the CPU oracle, so
 * the three cannot drift. Every value restates a constant of the reference
 * (1mishra/cbc); the citation gives the reference file:line it comes from.
 */
#ifndef CBCG_FORMAT_H
#define CBCG_FORMAT_H

#include <stdint.h>

/* ---- arithmetic coder (include/Arithmetic_stream.h:39, src/Arithmetic_stream.c:245-262) */
#define CBCG_AC_BITS        26u                       /* ARITHMETIC_WORD_LENGTH */
#define CBCG_AC_TOP         ((1u << CBCG_AC_BITS) - 1u)  /* initial u */
#define CBCG_AC_MSB         (1u << (CBCG_AC_BITS - 1u))
#define CBCG_AC_SMSB        (1u << (CBCG_AC_BITS - 2u))
#define CBCG_AC_LOWMASK     (CBCG_AC_MSB - 1u)        /* msb_clear_mask */

/* ---- model constants (src/sam_models.c:564: rescale = 1<<20 for every model) */
#define CBCG_RESCALE        (1u << 20)
#define CBCG_BITS_DELTA     7u                        /* include/read_compression.h:24 */
#define CBCG_LOSSLESS       8u                        /* include/Arithmetic_stream.h:36 */
#define CBCG_WELL_WORDS     32u                       /* src/sam_file_allocation.c:393 */
#define CBCG_WELL_DEBUG     0x55555555u               /* src/sam_file_allocation.c:399 (-DDEBUG) */
#define CBCG_MAX_READ_LEN   252u                      /* var ctx < 65535 and rlength byte < 255 */
#define CBCG_VAR_CONTEXTS   0xffffu                   /* src/sam_models.c:317 */
#define CBCG_MAX_POS_X      5000000u                  /* MAX_ALPHA, include/sam_block.h:54 */

/* ---- symbol streams. (stream << 24 | ctx, symbol) is the unit the tracer logs and
 * cbcg_extract_symbols returns. */
enum cbcg_stream {
    CBCG_S_CODEBOOK  = 0,  /* header ints, src/qv_codebook.c:14      ctx = byte index, alphabet 256, step 1  */
    CBCG_S_SAME_REF  = 1,  /* src/id_compression.c:39                ctx = 0,          alphabet 2,   step 10 */
    CBCG_S_RNAME     = 2,  /* src/id_compression.c:59                ctx = prev char,  alphabet 256, step 10 */
    CBCG_S_RLENGTH   = 3,  /* src/read_compression.c:29              ctx = byte index, alphabet 255, step 10 */
    CBCG_S_POS       = 4,  /* src/read_compression.c:113             dynamic alphabet,               step 10 */
    CBCG_S_POS_ALPHA = 5,  /* src/read_compression.c:75              ctx = byte index, alphabet 256, step 10 */
    CBCG_S_FLAG      = 6,  /* src/read_compression.c:50              ctx = 0,          alphabet 65536, step 8 */
    CBCG_S_MATCH     = 7,  /* src/read_compression.c:164             ctx = samePos<<1|prevMatch, alphabet 2, step 1 */
    CBCG_S_SNPS      = 8,  /* src/read_compression.c:193             ctx = 0,          alphabet L,   step 10 */
    CBCG_S_INDELS    = 9,  /* src/read_compression.c:212             ctx = 0,          alphabet L,   step 16 */
    CBCG_S_VAR       = 10, /* src/read_compression.c:230             ctx = prev<<1|strand (SNP: delta<<7 added), alphabet L, step 10 */
    CBCG_S_CHARS     = 11, /* src/read_compression.c:250             ctx = ref base (5 = insertion), alphabet 5, step 8 */
    CBCG_S_POS_X     = 12, /* NOT a coder stream: the raw value x = pos - prevPos + 1 that compress_pos
                              maps to POS / POS_ALPHA symbols through the dynamic alphabet. */
    CBCG_N_STREAMS   = 13
};

#define CBCG_SYM_KEY(stream, ctx)  (((uint32_t)(stream) << 24) | (uint32_t)(ctx))
#define CBCG_SYM_STREAM(key)       ((key) >> 24)
#define CBCG_SYM_CTX(key)          ((key) & 0xffffffu)

/* One entry of a symbol list (8 bytes, same layout the oracle tracer writes). */
typedef struct cbcg_symbol {
    uint32_t key;    /* CBCG_SYM_KEY(stream, ctx) */
    uint32_t value;
} cbcg_symbol;

/* ---- base codes (enum BASEPAIR, include/sam_block.h:165-172; char2basepair src/sam_models.c:11) */
enum { CBCG_BP_A = 0, CBCG_BP_C = 1, CBCG_BP_G = 2, CBCG_BP_T = 3, CBCG_BP_N = 4, CBCG_BP_O = 5 };

/* ---- per-read edit record produced by extraction (K1) and by the block decoder (K2d),
 * consumed by symbol emission (K1b) and reconstruction (K3). 16 bytes. */
typedef struct cbcg_read_rec {
    uint32_t pos;        /* 1-based POS */
    uint16_t flag;       /* SAM FLAG */
    uint16_t len;        /* SEQ length */
    uint32_t edit_off;   /* first entry of this read in the edit array */
    uint8_t  match;      /* 1: read equals reference[pos-1 .. pos-1+len) */
    uint8_t  n_snps;
    uint8_t  n_dels;
    uint8_t  n_ins;
} cbcg_read_rec;

/* Edit entries are u16, stored per read as [dels | snps | ins]:
 *   bits 0..7  position delta (reference: Dels[k], SNPs[k].pos, Insers[k].pos)
 *   bits 8..10 target base (SNP, insertion)
 *   bits 11..13 reference base (SNP; CBCG_BP_O for insertions) */
#define CBCG_EDIT(delta, target, refb)  ((uint16_t)((delta) | ((target) << 8) | ((refb) << 11)))
#define CBCG_EDIT_DELTA(e)   ((e) & 0xffu)
#define CBCG_EDIT_TARGET(e)  (((e) >> 8) & 7u)
#define CBCG_EDIT_REFB(e)    (((e) >> 11) & 7u)

/* ---- blocked container ("CBCB"), new design: the reference stream has no framing
 * (src/compression.c:128-155). Little endian. */
#define CBCG_MAGIC          0x42434243u   /* "CBCB" */
#define CBCG_VERSION        4u

/* Substreams of a block (container v4). The reference interleaves every symbol of a read in ONE arithmetic-coded
 * stream (src/read_compression.c:15-44), which makes a block one serial chain through every model. A block here holds
 * four independent arithmetic-coded substreams, one per group of models, each closed by the short flush and stored
 * back to back (A | B | C | D); the models, their contexts and the order of symbols WITHIN a substream are the
 * reference's. The encoder codes the four side by side (all symbols are known); the decoder runs three phases,
 * {A, B} -> C -> D, because match needs samePos (POS), the edits need the counts, and the base context needs the
 * reference base under the decoded position. A substream in which no symbol was coded is zero bytes long. */
#define CBCG_N_SUB          4u
enum cbcg_substream {
    CBCG_SUB_POS    = 0,   /* pos, pos_alpha                (compress_pos, compress_pos_alpha) */
    CBCG_SUB_FLAG   = 1,   /* rlength byte 0, flag          (compress_read :29-33, compress_flag) */
    CBCG_SUB_COUNTS = 2,   /* match, snps, indels           (compress_match, compress_snps, compress_indels) */
    CBCG_SUB_EDITS  = 3    /* var, chars                    (compress_var, compress_chars) */
};
static inline uint32_t cbcg_substream_of(uint32_t stream) {
    switch (stream) {
        case 4: case 5:   return CBCG_SUB_POS;      /* CBCG_S_POS, CBCG_S_POS_ALPHA */
        case 3: case 6:   return CBCG_SUB_FLAG;     /* CBCG_S_RLENGTH, CBCG_S_FLAG */
        case 7: case 8: case 9: return CBCG_SUB_COUNTS;   /* CBCG_S_MATCH, CBCG_S_SNPS, CBCG_S_INDELS */
        case 10: case 11: return CBCG_SUB_EDITS;    /* CBCG_S_VAR, CBCG_S_CHARS */
        default:          return CBCG_SUB_POS;      /* same_ref / rname / codebook are not coded in blocked containers */
    }
}
/* header word 9: low byte = gen_mode; bit 8 = every read is read_len_header bases long, so the length symbol
 * (src/read_compression.c:29-33, one zero-information coder step per read) is not coded either */
#define CBCG_MODE_GEN_MASK  0xffu
#define CBCG_MODE_FIXED_LEN 0x100u
/* bit 9: every block holds the four substreams above, and its index entry four byte counts; clear: ONE arithmetic-coded
 * stream per block with every symbol of a read in the reference's own order (src/read_compression.c:15-44), one byte
 * count. Same models, same symbols either way. */
#define CBCG_MODE_SPLIT4    0x200u
/* bits 16..23: the blocks of the first that many generations hold four substreams, the later ones a single stream (the
 * default cut: the narrow early generations are latency-bound and run three times faster as four short chains, the
 * wide last generation is bound by instruction issue and by its bytes per block, where one stream is cheaper). */
#define CBCG_MODE_SPLIT_GENS(mode)        (((mode) >> 16) & 0xffu)
#define CBCG_MODE_WITH_SPLIT_GENS(k)      (((uint32_t)(k) & 0xffu) << 16)
#define CBCG_BLOCK_NSUB(mode, gen)        ((((mode) & CBCG_MODE_SPLIT4) || (uint32_t)(gen) < CBCG_MODE_SPLIT_GENS(mode)) ? CBCG_N_SUB : 1u)
#define CBCG_MODE_LAYOUT_MASK             (CBCG_MODE_SPLIT4 | 0xff0000u)
/* API only, never stored: ask the encoder for the default cut's layout (four substreams in its narrow early generations) */
#define CBCG_MODE_REQ_HYBRID              0x400u

/* The FLAG model spends 65 536 / n of its probability on values never seen (src/sam_models.c:96-130: all-ones initial
 * state), so what a FLAG symbol costs depends on where the model total n stands below the rescale threshold 2^20
 * (0.09 bits per read at 2^20, 0.19 at 2^19). A merged snapshot therefore is not halved like update_model halves
 * (which leaves n anywhere in [2^19, 2^20)): its FLAG counts are scaled to the largest total that lets a block of
 * max_block_reads reads (the longest block of the container) code all its FLAG symbols without reaching the
 * threshold: c' = max(1, c * (T - 65536) / n), T = cbcg_flag_target(max_block_reads). */
static inline uint32_t cbcg_flag_target(uint32_t max_block_reads) {
    const uint64_t head = 8ull * max_block_reads + 64u;
    return head + (1u << 18) < CBCG_RESCALE ? (uint32_t)(CBCG_RESCALE - head) : (1u << 18);
}

/* Generation-primed blocks (gen_mode 1): the blocks of generation g start from the merged final states
 * of generations < g; the last generation takes the remaining reads in blocks of block_reads. The cut is the
 * encoder's choice and is written to the index (per-block read count and generation): a decoder never calls this.
 *
 * cbcg_gen_schedule: the default cut for a shard of n reads (constants measured with the CPU restatement on the named
 * shapes, profiles/r02_notes.md). What blocking costs against the reference's single stream:
 *   - 10.6 bytes per block with four substreams (7.9 of index, 2.7 of closing bits), 5.4 with one (4.7 + 0.7);
 *   - the `var` model (65 535 sparse contexts) keeps learning for millions of reads: a generation that codes the reads
 *     (C, rC] from a snapshot frozen at C reads loses about kappa ((r - 1) - ln r) bytes against a model that keeps
 *     adapting, kappa ~ 2 200 .. 3 400 (every other model is trained after a few thousand reads).
 * Early generations quadruple the cumulative read count from 4 blocks of 64 reads up to 262 144 (cheap: their blocks are
 * short); from there k late generations of equal ratio reach n, k chosen for the least serial depth (sum over the
 * generations of their block size) within the <= 1 % budget; block sizes follow the square root of the generation's
 * size (least depth for a given block count). Returns the number of early generations (<= CBCG_GEN_MAX) and the last
 * generation's block size in *last_reads. */
#define CBCG_GEN_MAX        16
static inline uint32_t cbcg_isqrt(uint64_t x) {
    uint64_t r = 0, bit = 1ull << 31;
    for (; bit; bit >>= 1) { const uint64_t t = r | bit; if (t * t <= x) r = t; }
    return (uint32_t)r;
}
static inline double cbcg_ln(double r) {                                    /* r >= 1; no libm: host, device and oracle agree */
    double acc = 0.0;
    while (r > 2.0) { r *= 0.5; acc += 0.6931471805599453; }
    const double y = (r - 1.0) / (r + 1.0);
    double t = y, s = 0.0;
    for (int i = 1; i < 40; i += 2) { s += t / (double)i; t *= y * y; }
    return acc + 2.0 * s;
}
static inline double cbcg_root(double x, uint32_t k) {                      /* x^(1/k), x >= 1, by bisection */
    double lo = 1.0, hi = x;
    for (int it = 0; it < 80; it++) {
        const double mid = 0.5 * (lo + hi);
        double p = 1.0;
        for (uint32_t j = 0; j < k; j++) p *= mid;
        if (p < x) lo = mid; else hi = mid;
    }
    return 0.5 * (lo + hi);
}
/* layout: 1 = one stream per block, 4 = four substreams per block, 0 = four in the narrow early generations (cumulative
 * reads <= 262 144) and one from there on (the default; *split_gens receives the number of four-substream generations).
 * Bytes per block c and time per read of serial depth tau differ by layout (measured on B200: tau = 4.5 us for the single
 * chain, 1.5 us for four short ones; c = 5.4 / 10.6 bytes): the least sum of tau_k b_k within the byte budget has
 * b_k ~ sqrt(c_k n_k / tau_k). A generation of single-stream blocks gains nothing from more blocks than one wave of
 * the GPU's resident warps (it is bound by instruction issue from there on), a four-substream generation is held to
 * one wave of resident CTAs. */
#define CBCG_SCHED_WAVE_BLOCKS 720u      /* resident CTAs of the four-substream kernel (148 SMs x 5, less a margin) */
#define CBCG_SCHED_WAVE_WARPS  2960u     /* resident warps of the single-chain kernel (148 SMs x 20) */
static inline uint32_t cbcg_gen_schedule(uint64_t n, uint32_t layout, uint32_t *count, uint32_t *reads, uint32_t *last_reads, uint32_t *split_gens) {
    uint64_t cum[CBCG_GEN_MAX + 1];
    uint32_t ne = 0;
    /* Around 3 M reads a coarser ladder (ratios 15, 7.4, 8, then the rest) measured inside the budget with a quarter less
       depth than the fourfold ladder (config 2: +0.90 % at depth 1 158 against +0.81 % at 1 533); below that size its
       loss is too large a share of the stream (config 1: +1.5 %), above it the last generation's is (6 M reads: +1.1 %). */
    const int coarse = n >= 2000000u && n < 4500000u;
    if (coarse) { cum[0] = 256; cum[1] = 3840; cum[2] = 28416; cum[3] = 225024; ne = 4; }
    else for (uint64_t c = 256; ne < 6 && c * 2 <= n; c *= 4) cum[ne++] = c;      /* 256 .. 262 144 */
    if (split_gens) *split_gens = layout == 4u ? 255u : 0u;
    if (!ne) { *last_reads = n > 64 ? (uint32_t)(n > 8192 ? 8192 : n) : 64u; return 0; }
    const uint32_t n_early = ne;
    /* kappa grows with the input (the var model keeps discovering contexts): 2 200 measured at 3 M reads, 3 400 at 6 M */
    double kappa = 2800.0;                                                       /* ... and no smaller below 3 M */
    { double x = (double)n / 3.0e6, f = 1.0; while (x >= 2.0) { x *= 0.5; f *= 1.516; } if (x > 1.0) f *= 1.0 + 0.516 * (x - 1.0); kappa *= f; }
    const double S = 1.66 * (double)n, room = 0.0092 * S, early_loss = n < 2000000u ? 6000.0 : 3000.0;
    const double c_early = layout == 1u ? 5.4 : 10.6, c_late = layout == 4u ? 10.6 : 5.4;
    const double t_early = layout == 1u ? 4.5 : 1.5, t_late = layout == 4u ? 1.5 : 4.5;
    const double c0 = (double)cum[ne - 1];
    uint32_t best_k = 1; double best_time = 0.0, best_bytes = 200.0;
    for (uint32_t k = 1; k <= 6 && ne + k - 1 <= CBCG_GEN_MAX; k++) {
        const double r = cbcg_root((double)n / c0, k);
        if (r > 4.0 && k < 6 && !coarse) continue;                               /* a frozen snapshot coding more than 4x its training costs more than the model says */
        const double loss = coarse ? (k == 1 ? 19000.0 * (double)n / 3.0e6 : 1.0e12) : early_loss + kappa * (double)k * ((r - 1.0) - cbcg_ln(r));
        double bytes = room - loss;
        if (bytes < 200.0) bytes = 200.0;
        double sum = 0.0, c = 0.0;                                               /* time = (sum sqrt(c n tau))^2 / bytes */
        for (uint32_t g = 0; g < ne; g++) { sum += (double)cbcg_isqrt((uint64_t)((double)(cum[g] - (uint64_t)c) * c_early * t_early)); c = (double)cum[g]; }
        for (uint32_t j = 1; j <= k; j++) { const double nx = j == k ? (double)n : c * r; sum += (double)cbcg_isqrt((uint64_t)((nx - c) * c_late * t_late)); c = nx; }
        const double time = sum * sum / bytes;
        if (best_time == 0.0 || time < best_time) { best_time = time; best_k = k; best_bytes = bytes; }
        if (r < 2.0) break;
    }
    {                                                                            /* the late generations but the last */
        const double r = cbcg_root((double)n / c0, best_k);
        double c = c0;
        for (uint32_t j = 1; j < best_k; j++) { c *= r; cum[ne++] = (uint64_t)c; }
    }
    cum[ne] = n;                                                                 /* stage ne = the last generation */
    if (split_gens && layout == 0u) *split_gens = n_early;
    uint64_t bsz[CBCG_GEN_MAX + 1];
    int fixed[CBCG_GEN_MAX + 1];
    for (uint32_t g = 0; g <= ne; g++) fixed[g] = 0;
    double bytes = best_bytes;
    for (int pass = 0; pass < 3; pass++) {                                       /* clamped generations hand their bytes to the others */
        double sum = 0.0; uint64_t prev = 0;
        for (uint32_t g = 0; g <= ne; g++) {
            const uint64_t size = cum[g] - prev; prev = cum[g];
            const int early = g < n_early;
            if (!fixed[g]) sum += (double)cbcg_isqrt((uint64_t)((double)size * (early ? c_early * t_early : c_late * t_late)));
        }
        const double lambda = bytes > 1.0 ? sum / bytes : sum;
        int changed = 0; prev = 0;
        for (uint32_t g = 0; g <= ne; g++) {
            const uint64_t size = cum[g] - prev; prev = cum[g];
            if (fixed[g]) continue;
            const int early = g < n_early;
            const int four = early ? layout != 1u : layout == 4u;
            uint64_t b = (uint64_t)(lambda * (double)cbcg_isqrt((uint64_t)((double)size * (early ? c_early / t_early : c_late / t_late) * 1.0e4)) / 100.0);
            if (b < 32) b = 32;
            const uint64_t wave = four ? CBCG_SCHED_WAVE_BLOCKS : CBCG_SCHED_WAVE_WARPS;
            const uint64_t floor_b = (size + wave - 1) / wave;                    /* more blocks than one wave buy no time */
            if (b < floor_b) { b = floor_b; fixed[g] = 1; changed = 1; bytes -= (early ? c_early : c_late) * (double)((size + b - 1) / b); }
            bsz[g] = b;
        }
        if (!changed) break;
        if (bytes < 200.0) bytes = 200.0;
    }
    for (uint32_t g = 0; g < ne; g++) {
        uint64_t b = g == 0 ? 64 : bsz[g];
        if (b > 16384) b = 16384;
        const uint64_t size = cum[g] - (g ? cum[g - 1] : 0);
        reads[g] = (uint32_t)b; count[g] = (uint32_t)((size + b - 1) / b);
    }
    uint64_t lb = bsz[ne];
    if (lb < 64) lb = 64;
    if (lb > 16384) lb = 16384;
    *last_reads = (uint32_t)lb;
    return ne;
}
#define CBCG_SNAP_POS_MAX   4096u         /* a snapshot keeps at most this many POS alphabet entries */
/* FLAG in blocked containers is bounded by design (the reference's dense model takes all 65 536 values at
 * 65 536 scan steps per symbol, src/sam_models.c:96-130, src/stream_model.c:64-67). A block, and a snapshot, ADAPT at most
 * CBCG_FLAG_ADAPT_MAX distinct values (values whose count is not the initial 1):
 *   F1 (block): a value that is not yet adapted arrives while CBCG_FLAG_ADAPT_MAX are -- it is coded with its count of 1 and
 *       the model is left as it was (no increment of the count or of the total), by encoder and decoder alike;
 *   F2 (snapshot merge): if more than CBCG_FLAG_ADAPT_MAX merged counts differ from 1, the counts <= T go back to 1, T the
 *       smallest threshold that leaves at most CBCG_FLAG_ADAPT_MAX of them (after the clamp and the scaling to the target
 *       total); the total is the sum of what is left.
 * The single-block mode is the reference's own stream and adapts every value; there the table's size is an error
 * (CBCG_ERR_LIMIT). */
#define CBCG_FLAG_ADAPT_MAX 256u

#endif /* CBCG_FORMAT_H */
