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
#define CBCG_VERSION        3u
/* header word 9: low byte = gen_mode; bit 8 = every read is read_len_header bases long, so the length symbol
 * (src/read_compression.c:29-33, one zero-information coder step per read) is not coded either */
#define CBCG_MODE_GEN_MASK  0xffu
#define CBCG_MODE_FIXED_LEN 0x100u

/* Generation-primed blocks (gen_mode 1): generation i has CBCG_GEN_COUNTS[i] blocks of
 * CBCG_GEN_READS[i] reads, each starting from the merged final states of the generation before; the
 * last generation takes all remaining reads in blocks of block_reads. */
#define CBCG_GEN_LEVELS     4
#define CBCG_GEN_COUNTS     { 16u, 112u, 384u, 1536u }
#define CBCG_GEN_READS      { 16u, 32u, 64u, 128u }
#define CBCG_SNAP_POS_MAX   4096u         /* a snapshot keeps at most this many POS alphabet entries */

#endif /* CBCG_FORMAT_H */
