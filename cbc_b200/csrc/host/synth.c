/*
 * synth.c -- deterministic synthetic workload generator.
 *
 * Produces a random ACGT genome and position-sorted aligned reads with SAM-spec CIGAR
 * and MD strings, directly in the SoA batch layout the C ABI consumes (include/cbcg.h),
 * and can print the same reads as SAM / FASTA text for the CPU reference. Every read
 * has its own RNG stream keyed by (seed, ordinal), so any sub-range can be regenerated.
 *
 * The SAM records obey the reference's input contract: MD is never the
 * last field (load_sam_line keeps the newline on it, src/sam_file_allocation.c:507-511),
 * lines are < 1024 bytes, mapped reads only, sorted by position.
 */
#include "synth.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t splitmix(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
typedef struct { uint64_t s[2]; } rng;
static void rng_seed(rng *r, uint64_t seed, uint64_t stream) {
    uint64_t x = seed * 0xd1342543de82ef95ULL + stream * 0x2545f4914f6cdd1dULL + 0x1234567ULL;
    r->s[0] = splitmix(&x); r->s[1] = splitmix(&x);
    if (!(r->s[0] | r->s[1])) r->s[0] = 1;
}
static uint64_t rng_next(rng *r) {         /* xoroshiro128+ */
    uint64_t a = r->s[0], b = r->s[1], out = a + b;
    b ^= a;
    r->s[0] = ((a << 24) | (a >> 40)) ^ b ^ (b << 16);
    r->s[1] = (b << 37) | (b >> 27);
    return out;
}
static double rng_u(rng *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static uint32_t rng_below(rng *r, uint32_t n) { return (uint32_t)(((rng_next(r) >> 32) * (uint64_t)n) >> 32); }

static const char BASES[4] = { 'A', 'C', 'G', 'T' };

void cbcs_genome(uint64_t seed, uint32_t chr, uint8_t *bases, uint64_t len) {
    rng r; rng_seed(&r, seed ^ 0x67656e6f6d65ULL, chr);
    uint64_t i = 0;
    while (i < len) {
        uint64_t v = rng_next(&r);
        for (int k = 0; k < 32 && i < len; k++, v >>= 2) bases[i++] = (uint8_t)BASES[v & 3];
    }
}

static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

static uint32_t put_num(uint8_t *dst, uint32_t v) {
    char tmp[12]; int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int i = 0; i < n; i++) dst[i] = (uint8_t)tmp[n - 1 - i];
    return (uint32_t)n;
}

/* Generates reads [lo, hi) of the n reads of one chromosome (all n positions are drawn and sorted, so a sub-range
 * is the same reads the whole range would hold there). Returns 0 or <0 (capacity). */
static int gen_chr(const cbcs_params *p, uint32_t chr, const uint8_t *ref, uint64_t ref_len,
                   uint64_t first_ordinal, uint64_t n, uint64_t lo, uint64_t hi, cbcs_out *o) {
    uint32_t lmax = p->len_max;
    if (ref_len < 4ull * lmax + 16) return -1;
    uint64_t span = ref_len - 2ull * lmax - 8;
    uint32_t *posv = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    if (!posv) return -2;
    rng pr; rng_seed(&pr, p->seed ^ 0x706f73ULL, chr);
    for (uint64_t i = 0; i < n; i++) posv[i] = 1 + (uint32_t)(rng_u(&pr) * (double)span);
    qsort(posv, n, sizeof(uint32_t), cmp_u32);

    uint8_t cig[2048], md[2048];
    if (hi > n) hi = n;
    for (uint64_t i = lo; i < hi; i++) {
        uint64_t ord = first_ordinal + i, r_idx = o->n_reads;
        rng r; rng_seed(&r, p->seed, ord);
        uint32_t len = p->len_min + (p->len_max > p->len_min ? rng_below(&r, p->len_max - p->len_min + 1) : 0);
        uint32_t clip_l = 0, clip_r = 0;
        if (p->p_clip > 0) {
            if (rng_u(&r) < p->p_clip) clip_l = 1 + rng_below(&r, 8);
            if (rng_u(&r) < p->p_clip) clip_r = 1 + rng_below(&r, 8);
            if (clip_l + clip_r + 24 > len) clip_l = clip_r = 0;
        }
        if (o->seq_size + len > o->seq_cap) { free(posv); return -3; }
        uint8_t *seq = o->seq + o->seq_size;
        uint32_t q = 0, nc = 0, nm = 0;
        uint64_t rp = (uint64_t)posv[i] - 1;
        for (; q < clip_l; q++) seq[q] = (uint8_t)BASES[rng_below(&r, 4)];
        if (clip_l) { nc += put_num(cig + nc, clip_l); cig[nc++] = 'S'; }
        uint32_t body_end = len - clip_r;
        int prev_indel = 1;          /* forbids an indel as the first aligned event */
        int prev_mismatch = 0;
        uint32_t run_m = 0, md_run = 0;
        int oops = 0;
        while (q < body_end) {
            double u = (p->p_indel > 0) ? rng_u(&r) : 1.0;
            int last = (q + 1 == body_end);
            if (!prev_indel && !last && u < p->p_indel * 0.5) {            /* insertion of one base */
                if (run_m) { nc += put_num(cig + nc, run_m); cig[nc++] = 'M'; run_m = 0; }
                nc += put_num(cig + nc, 1); cig[nc++] = 'I';
                seq[q++] = (uint8_t)BASES[rng_below(&r, 4)];
                prev_indel = 1; prev_mismatch = 0;
            } else if (!prev_indel && !last && u < p->p_indel && !(p->avoid_b3 && clip_l && prev_mismatch)) {
                uint32_t dl = 1 + rng_below(&r, 3);                        /* deletion of 1..3 bases */
                if (run_m) { nc += put_num(cig + nc, run_m); cig[nc++] = 'M'; run_m = 0; }
                nc += put_num(cig + nc, dl); cig[nc++] = 'D';
                nm += put_num(md + nm, md_run); md_run = 0;
                md[nm++] = '^';
                for (uint32_t k = 0; k < dl; k++) md[nm++] = ref[rp + k];
                rp += dl;
                prev_indel = 1; prev_mismatch = 0;
            } else {                                                       /* aligned base */
                uint8_t rb = ref[rp++], b = rb;
                double us = rng_u(&r);
                if (us < p->p_sub) {
                    uint32_t k = rng_below(&r, 3);
                    for (int c = 0, seen = 0; c < 4; c++) if ((uint8_t)BASES[c] != rb) { if ((uint32_t)seen == k) b = (uint8_t)BASES[c]; seen++; }
                } else if (us < p->p_sub + p->p_n) b = 'N';
                if (b != rb) { nm += put_num(md + nm, md_run); md_run = 0; md[nm++] = rb; prev_mismatch = 1; }
                else { md_run++; prev_mismatch = 0; }
                seq[q++] = b; run_m++;
                prev_indel = 0;
            }
            if (nc > 1900 || nm > 1900) { oops = 1; break; }
        }
        if (oops) { free(posv); return -4; }
        if (run_m) { nc += put_num(cig + nc, run_m); cig[nc++] = 'M'; }
        nm += put_num(md + nm, md_run);
        for (uint32_t k = 0; k < clip_r; k++) seq[q++] = (uint8_t)BASES[rng_below(&r, 4)];
        if (clip_r) { nc += put_num(cig + nc, clip_r); cig[nc++] = 'S'; }

        if (o->cigar_size + nc > o->cigar_cap || o->md_size + nm > o->md_cap || r_idx >= o->reads_cap) { free(posv); return -3; }
        memcpy(o->cigar + o->cigar_size, cig, nc);
        memcpy(o->md + o->md_size, md, nm);
        int rev = rng_u(&r) < p->p_rev;
        uint16_t flag;
        if (p->flag_mode == 1) { uint32_t mate = rng_below(&r, 2); flag = rev ? (mate ? 147 : 83) : (mate ? 163 : 99); }
        else flag = rev ? 16 : 0;
        o->pos[r_idx] = posv[i]; o->flag[r_idx] = flag; o->seq_len[r_idx] = (uint16_t)len; o->chr[r_idx] = chr;
        o->seq_off[r_idx] = o->seq_size; o->cigar_off[r_idx] = o->cigar_size; o->md_off[r_idx] = o->md_size;
        o->seq_size += len; o->cigar_size += nc; o->md_size += nm;
        o->n_reads = r_idx + 1;
        o->seq_off[r_idx + 1] = o->seq_size; o->cigar_off[r_idx + 1] = o->cigar_size; o->md_off[r_idx + 1] = o->md_size;
    }
    free(posv);
    return 0;
}

/* Reads [r0, r1) of the whole position-sorted input (ordinals over all chromosomes): what one region shard holds.
 * chr_bases[c] may be NULL for chromosomes the range does not touch. */
int cbcs_reads_range(const cbcs_params *p, const uint8_t *const *chr_bases, const uint64_t *chr_len, uint64_t r0, uint64_t r1,
                     cbcs_out *o) {
    if (!p || !o || p->n_chr == 0 || p->len_min == 0 || p->len_max < p->len_min || p->len_max > 252) return -1;
    o->n_reads = 0; o->seq_size = o->cigar_size = o->md_size = 0;
    if (o->reads_cap) { o->seq_off[0] = o->cigar_off[0] = o->md_off[0] = 0; }
    uint64_t total = 0;
    for (uint32_t c = 0; c < p->n_chr; c++) total += chr_len[c];
    uint64_t done = 0;
    for (uint32_t c = 0; c < p->n_chr; c++) {
        uint64_t n = (c + 1 == p->n_chr) ? p->n_reads - done
                                         : (uint64_t)((double)p->n_reads * (double)chr_len[c] / (double)total);
        if (done < r1 && done + n > r0) {
            if (!chr_bases[c]) return -5;
            const uint64_t lo = r0 > done ? r0 - done : 0, hi = r1 - done < n ? r1 - done : n;
            int rc = gen_chr(p, c, chr_bases[c], chr_len[c], done, n, lo, hi, o);
            if (rc) return rc;
        }
        done += n;
    }
    return 0;
}

/* Read counts per chromosome, as cbcs_reads / cbcs_reads_range deal them out (out[n_chr]). */
void cbcs_chr_counts(const cbcs_params *p, const uint64_t *chr_len, uint64_t *out) {
    uint64_t total = 0, done = 0;
    for (uint32_t c = 0; c < p->n_chr; c++) total += chr_len[c];
    for (uint32_t c = 0; c < p->n_chr; c++) {
        out[c] = (c + 1 == p->n_chr) ? p->n_reads - done : (uint64_t)((double)p->n_reads * (double)chr_len[c] / (double)total);
        done += out[c];
    }
}

int cbcs_reads(const cbcs_params *p, const uint8_t *const *chr_bases, const uint64_t *chr_len, cbcs_out *o) {
    if (!p) return -1;
    return cbcs_reads_range(p, chr_bases, chr_len, 0, p->n_reads, o);
}

int cbcs_write_fasta(const char *path, uint32_t n_chr, const char *const *names,
                     const uint8_t *const *chr_bases, const uint64_t *chr_len) {
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    for (uint32_t c = 0; c < n_chr; c++) {
        fprintf(f, ">%s\n", names[c]);
        for (uint64_t i = 0; i < chr_len[c]; i += 60) {
            uint64_t n = chr_len[c] - i < 60 ? chr_len[c] - i : 60;
            fwrite(chr_bases[c] + i, 1, n, f);
            fputc('\n', f);
        }
    }
    return fclose(f) ? -1 : 0;
}

int cbcs_write_sam(const char *path, const cbcs_out *o, uint32_t n_chr, const char *const *names,
                   const uint64_t *chr_len, int with_header) {
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    static char big[1 << 20];
    setvbuf(f, big, _IOFBF, sizeof big);
    if (with_header) {
        fprintf(f, "@HD\tVN:1.6\tSO:coordinate\n");
        for (uint32_t c = 0; c < n_chr; c++) fprintf(f, "@SQ\tSN:%s\tLN:%llu\n", names[c], (unsigned long long)chr_len[c]);
    }
    char qual[256]; memset(qual, 'I', sizeof qual);
    for (uint64_t r = 0; r < o->n_reads; r++) {
        fprintf(f, "r%llu\t%u\t%s\t%u\t60\t", (unsigned long long)r, o->flag[r], names[o->chr[r]], o->pos[r]);
        fwrite(o->cigar + o->cigar_off[r], 1, o->cigar_off[r + 1] - o->cigar_off[r], f);
        fputs("\t*\t0\t0\t", f);
        fwrite(o->seq + o->seq_off[r], 1, o->seq_len[r], f);
        fputc('\t', f);
        fwrite(qual, 1, o->seq_len[r], f);
        fputs("\tMD:Z:", f);
        fwrite(o->md + o->md_off[r], 1, o->md_off[r + 1] - o->md_off[r], f);
        fputs("\tAS:i:0\n", f);
    }
    return fclose(f) ? -1 : 0;
}
