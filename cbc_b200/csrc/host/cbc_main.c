/*
 * cbc_main.c -- the reference's command line over the B200 path.
 *
 *   cbc -c <sam> <out> <ref.fa>     compress   (README.md:59;  checked-in spelling `-c <ratio> ...`, src/main.c:114-123, also accepted)
 *   cbc -d <in>  <out> <ref.fa>     decompress (README.md:67;  checked-in spelling `-x`, src/main.c:136-139, also accepted)
 * options: -b N   reads per block (default: sized to the GPU, CBCG_BLOCK_AUTO; the container is the blocked "CBCB" format)
 *          -1     single-block mode: the reference's own stream, byte-identical to `program -c 1` built with -DDEBUG
 *          -l     variable-length reads: header read length = longest SEQ (src/main.c -l)
 *          -g N   CUDA device
 * Host C only: SAM/FASTA ingest (sam_ingest.c) and file I/O; the coding runs on the GPU through include/cbcg.h.
 * Prints the reference's progress lines (src/compression.c:157,166,206). Returns 0 on success (the reference's
 * main returns 1, src/main.c:370).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "cbcg.h"
#include "sam_ingest.h"

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

static int usage(void) {
    fprintf(stderr, "usage: cbc -c [-1] [-l] [-b reads_per_block] [-g device] <sam> <out> <ref.fa>\n"
                    "       cbc -d [-g device] <in> <out> <ref.fa>\n");
    return 2;
}

int main(int argc, char **argv) {
    int mode = 0, single = 0, var_length = 0, device = 0;
    uint32_t block_reads = CBCG_BLOCK_AUTO;                 /* sized to the GPU (cbcg.h) */
    const char *files[4]; int nfiles = 0;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (!strcmp(a, "-c")) mode = 'c';
        else if (!strcmp(a, "-d") || !strcmp(a, "-x")) mode = 'd';
        else if (!strcmp(a, "-1")) single = 1;
        else if (!strcmp(a, "-l")) var_length = 1;
        else if (!strcmp(a, "-b") && i + 1 < argc) block_reads = (uint32_t)strtoul(argv[++i], NULL, 10);
        else if (!strcmp(a, "-g") && i + 1 < argc) device = atoi(argv[++i]);
        else if (a[0] == '-' && a[1]) return usage();
        else if (nfiles < 4) files[nfiles++] = a;
        else return usage();
    }
    /* the checked-in reference wants a ratio before the file names (`-c 1 sam out ref`, 1 = lossless): accept and drop it */
    if (mode == 'c' && nfiles == 4) {
        char *e; (void)strtod(files[0], &e);
        if (*e != 0) return usage();
        files[0] = files[1]; files[1] = files[2]; files[2] = files[3]; nfiles = 3;
    }
    if (!mode || nfiles != 3) { fprintf(stderr, "Missing required filenames\n"); return usage(); }
    if (!single && block_reads == 0) block_reads = CBCG_BLOCK_AUTO;

    char err[256] = "";
    const double t0 = now();
    cbch_fasta fa;
    if (cbch_read_fasta(files[2], &fa, err, sizeof err)) { fprintf(stderr, "cbc: %s\n", err); return 1; }
    cbcg_ctx *ctx = NULL;
    int rc = cbcg_create(device, &ctx);
    if (rc) { fprintf(stderr, "cbc: %s\n", cbcg_strerror(rc)); return 1; }
    rc = cbcg_set_reference(ctx, fa.n, (const char *const *)fa.names, (const uint8_t *const *)fa.bases, fa.len);
    if (rc) { fprintf(stderr, "cbc: %s\n", cbcg_last_error(ctx)); return 1; }

    int status = 0;
    if (mode == 'c') {
        printf("Compressing...\n");
        cbch_batch hb;
        if (cbch_read_sam(files[0], &fa, var_length, &hb, err, sizeof err)) { fprintf(stderr, "cbc: %s\n", err); return 1; }
        const double t1 = now();
        cbcg_batch b; cbch_batch_view(&hb, &b);
        cbcg_encode_opts o = { hb.read_len_header ? hb.read_len_header : 1u, single ? 0u : block_reads, single ? 0u : 1u, 0u };
        uint64_t cap = cbcg_encode_bound(&b, &o), n = 0;
        uint8_t *out = (uint8_t *)malloc(cap ? cap : 1);
        if (!out) { fprintf(stderr, "cbc: out of memory\n"); return 1; }
        rc = cbcg_encode(ctx, &b, &o, out, cap, &n);
        if (rc == CBCG_ERR_CAPACITY && n > cap) { free(out); out = (uint8_t *)malloc(n); rc = out ? cbcg_fetch_container(ctx, out, n, &n) : CBCG_ERR_NOMEM; }
        const double t2 = now();
        if (rc) { fprintf(stderr, "cbc: %s\n", cbcg_last_error(ctx)); status = 1; }
        else {
            FILE *f = fopen(files[1], "wb");
            if (!f || fwrite(out, 1, n, f) != n || fclose(f)) { fprintf(stderr, "cbc: cannot write %s\n", files[1]); status = 1; }
            cbcg_stats st; cbcg_get_stats(ctx, &st);
            printf("Final Size: %llu\n", (unsigned long long)n);
            printf("Compression took %f\n", t2 - t1);
            printf("reads %llu (unmapped skipped %llu), blocks %llu, %.4f bits/base, ingest %.3f s, device %.3f ms\n",
                   (unsigned long long)hb.n_reads, (unsigned long long)hb.n_unmapped, (unsigned long long)st.n_blocks,
                   hb.seq_off[hb.n_reads] ? 8.0 * (double)n / (double)hb.seq_off[hb.n_reads] : 0.0, t1 - t0, st.ms_total);
        }
        free(out); cbch_free_batch(&hb);
    } else {
        printf("Decompressing...\n");
        FILE *f = fopen(files[0], "rb");
        if (!f) { fprintf(stderr, "cbc: cannot open %s\n", files[0]); return 1; }
        fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
        uint8_t *in = (uint8_t *)malloc(sz > 0 ? (size_t)sz : 1);
        if (!in || fread(in, 1, (size_t)sz, f) != (size_t)sz) { fprintf(stderr, "cbc: cannot read %s\n", files[0]); return 1; }
        fclose(f);
        const double t1 = now();
        uint64_t n_reads = 0, cap = 0, n = 0;
        const int legacy = cbcg_decoded_size(in, (uint64_t)sz, &n_reads, &cap) != CBCG_OK;   /* no "CBCB" header: a reference stream */
        if (legacy) cap = 1u << 20;
        uint8_t *text = (uint8_t *)malloc(cap ? cap : 1);
        rc = text ? cbcg_decode(ctx, in, (uint64_t)sz, legacy, text, cap, &n, &n_reads) : CBCG_ERR_NOMEM;
        if (rc == CBCG_ERR_CAPACITY && n > cap) { free(text); text = (uint8_t *)malloc(n); rc = text ? cbcg_fetch_decoded(ctx, text, n, &n) : CBCG_ERR_NOMEM; }
        const double t2 = now();
        if (rc) { fprintf(stderr, "cbc: %s\n", cbcg_last_error(ctx)); status = 1; }
        else {
            FILE *g = fopen(files[1], "wb");
            if (!g || fwrite(text, 1, n, g) != n || fclose(g)) { fprintf(stderr, "cbc: cannot write %s\n", files[1]); status = 1; }
            printf("Decompression took %f\n", t2 - t1);
            printf("reads %llu\n", (unsigned long long)n_reads);
        }
        free(text); free(in);
    }
    cbcg_destroy(ctx);
    cbch_free_fasta(&fa);
    printf("Total time elapsed: %f seconds.\n", now() - t0);
    return status;
}
