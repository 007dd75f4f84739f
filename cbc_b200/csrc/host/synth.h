/* synth.h -- synthetic genome / aligned-read generator (see synth.c). */
#ifndef CBC_SYNTH_H
#define CBC_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cbcs_params {
    uint64_t seed;
    uint64_t n_reads;
    uint32_t n_chr;
    uint32_t len_min, len_max;   /* SEQ length range (fixed length: equal) */
    double p_sub;                /* per aligned base substitution probability */
    double p_indel;              /* per aligned base: p/2 one-base insertion, p/2 deletion of 1..3 */
    double p_clip;               /* per end soft-clip probability, length 1..8 */
    double p_rev;                /* reverse-strand probability */
    double p_n;                  /* per aligned base probability of an 'N' in the read */
    uint32_t flag_mode;          /* 0: FLAG in {0,16}; 1: paired-like {99,147,83,163} */
    uint32_t avoid_b3;           /* 1: no deletion right after a mismatch in reads with a leading clip
                                    (the reference encoder aborts on it) */
} cbcs_params;

/* Caller-allocated SoA batch (same layout as cbcg_batch) with capacities. */
typedef struct cbcs_out {
    uint64_t n_reads, reads_cap;
    uint32_t *pos; uint16_t *flag; uint16_t *seq_len; uint32_t *chr;
    uint64_t *seq_off;   uint8_t *seq;   uint64_t seq_size,   seq_cap;     /* offsets: reads_cap+1 */
    uint64_t *cigar_off; uint8_t *cigar; uint64_t cigar_size, cigar_cap;
    uint64_t *md_off;    uint8_t *md;    uint64_t md_size,    md_cap;
} cbcs_out;

void cbcs_genome(uint64_t seed, uint32_t chr, uint8_t *bases, uint64_t len);
int cbcs_reads(const cbcs_params *p, const uint8_t *const *chr_bases, const uint64_t *chr_len, cbcs_out *o);
/* reads [r0, r1) of the whole position-sorted input: one region shard (chr_bases[c] may be NULL where untouched) */
int cbcs_reads_range(const cbcs_params *p, const uint8_t *const *chr_bases, const uint64_t *chr_len, uint64_t r0, uint64_t r1,
                     cbcs_out *o);
void cbcs_chr_counts(const cbcs_params *p, const uint64_t *chr_len, uint64_t *out);
int cbcs_write_fasta(const char *path, uint32_t n_chr, const char *const *names,
                     const uint8_t *const *chr_bases, const uint64_t *chr_len);
int cbcs_write_sam(const char *path, const cbcs_out *o, uint32_t n_chr, const char *const *names,
                   const uint64_t *chr_len, int with_header);

#ifdef __cplusplus
}
#endif
#endif
