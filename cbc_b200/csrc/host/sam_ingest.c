/* sam_ingest.c -- see sam_ingest.h. Files are mapped whole and split with memchr: no per-line buffers, so the
 * reference's 1024-byte line limit (src/sam_file_allocation.c:444) and its "MD must not be the last field"
 * trap (:507-511) do not exist here. */
#define _GNU_SOURCE
#include "sam_ingest.h"

#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

static int fail(char *err, size_t n, int rc, const char *fmt, ...) {
    if (err && n) { va_list ap; va_start(ap, fmt); vsnprintf(err, n, fmt, ap); va_end(ap); }
    return rc;
}

typedef struct { const uint8_t *p; size_t n; int fd; } mapped;
static int map_file(const char *path, mapped *m) {
    m->p = NULL; m->n = 0;
    m->fd = open(path, O_RDONLY);
    if (m->fd < 0) return -1;
    struct stat st;
    if (fstat(m->fd, &st)) { close(m->fd); return -1; }
    m->n = (size_t)st.st_size;
    if (m->n == 0) return 0;
    void *p = mmap(NULL, m->n, PROT_READ, MAP_PRIVATE, m->fd, 0);
    if (p == MAP_FAILED) { close(m->fd); return -1; }
    madvise(p, m->n, MADV_SEQUENTIAL);
    m->p = (const uint8_t *)p;
    return 0;
}
static void unmap_file(mapped *m) { if (m->p) munmap((void *)m->p, m->n); if (m->fd >= 0) close(m->fd); m->p = NULL; }

/* ------------------------------------------------------------------ FASTA */
int cbch_read_fasta(const char *path, cbch_fasta *out, char *err, size_t errlen) {
    memset(out, 0, sizeof *out);
    mapped m;
    if (map_file(path, &m)) return fail(err, errlen, CBCH_ERR_IO, "cannot open %s", path);
    uint32_t cap = 0;
    const uint8_t *p = m.p, *end = m.p + m.n;
    int rc = CBCH_OK;
    while (p < end) {
        const uint8_t *nl = memchr(p, '\n', (size_t)(end - p));
        const uint8_t *le = nl ? nl : end;
        if (p < le && *p == '>') {
            if (out->n == cap) {
                cap = cap ? cap * 2 : 32;
                out->names = realloc(out->names, cap * sizeof *out->names);
                out->bases = realloc(out->bases, cap * sizeof *out->bases);
                out->len = realloc(out->len, cap * sizeof *out->len);
                if (!out->names || !out->bases || !out->len) { rc = CBCH_ERR_NOMEM; break; }
            }
            const uint8_t *q = p + 1;
            while (q < le && *q != ' ' && *q != '\t' && *q != '\r') q++;
            size_t nl_len = (size_t)(q - (p + 1));
            char *name = malloc(nl_len + 1);
            if (!name) { rc = CBCH_ERR_NOMEM; break; }
            memcpy(name, p + 1, nl_len); name[nl_len] = 0;
            /* record body: up to the next '>' at a line start; size bound = bytes until then */
            const uint8_t *body = nl ? nl + 1 : end, *scan = body, *next = end;
            while (scan < end) {
                if (*scan == '>') { next = scan; break; }
                const uint8_t *e2 = memchr(scan, '\n', (size_t)(end - scan));
                if (!e2) break;
                scan = e2 + 1;
            }
            uint8_t *bases = malloc((size_t)(next - body) + 1);
            if (!bases) { free(name); rc = CBCH_ERR_NOMEM; break; }
            uint64_t len = 0;
            for (const uint8_t *s = body; s < next;) {
                const uint8_t *e2 = memchr(s, '\n', (size_t)(next - s));
                const uint8_t *le2 = e2 ? e2 : next;
                size_t k = (size_t)(le2 - s);
                if (k && s[k - 1] == '\r') k--;
                memcpy(bases + len, s, k); len += k;
                s = e2 ? e2 + 1 : next;
            }
            out->names[out->n] = name; out->bases[out->n] = bases; out->len[out->n] = len; out->n++;
            p = next;
            continue;
        }
        p = nl ? nl + 1 : end;
    }
    unmap_file(&m);
    if (rc) { cbch_free_fasta(out); return fail(err, errlen, rc, "out of memory reading %s", path); }
    if (out->n == 0) return fail(err, errlen, CBCH_ERR_PARSE, "%s holds no FASTA record", path);
    return CBCH_OK;
}
void cbch_free_fasta(cbch_fasta *fa) {
    for (uint32_t i = 0; i < fa->n; i++) { free(fa->names[i]); free(fa->bases[i]); }
    free(fa->names); free(fa->bases); free(fa->len);
    memset(fa, 0, sizeof *fa);
}

/* ------------------------------------------------------------------ SAM */
static int grow(void **p, uint64_t *cap, uint64_t need, size_t elem) {
    if (need <= *cap) return 0;
    uint64_t c = *cap ? *cap : 1024;
    while (c < need) c += c / 2 + 1024;
    void *q = realloc(*p, (size_t)(c * elem));
    if (!q) return -1;
    *p = q; *cap = c;
    return 0;
}
static int reserve_reads(cbch_batch *b, uint64_t n) {
    if (n <= b->cap) return 0;
    uint64_t c = b->cap ? b->cap : 4096;
    while (c < n) c += c / 2;
#define RS(field, extra) do { void *q = realloc(b->field, (size_t)((c + extra) * sizeof *b->field)); if (!q) return -1; b->field = q; } while (0)
    RS(pos, 0); RS(flag, 0); RS(seq_len, 0); RS(chr, 0); RS(seq_off, 1); RS(cigar_off, 1); RS(md_off, 1);
#undef RS
    b->cap = c;
    return 0;
}
static int parse_u32(const uint8_t *s, const uint8_t *e, uint32_t *v) {
    if (s >= e) return -1;
    uint64_t x = 0;
    for (; s < e; s++) { if (*s < '0' || *s > '9') return -1; x = x * 10 + (uint64_t)(*s - '0'); if (x > 0xffffffffull) return -1; }
    *v = (uint32_t)x;
    return 0;
}

int cbch_read_sam(const char *path, const cbch_fasta *fa, int var_length, cbch_batch *b, char *err, size_t errlen) {
    memset(b, 0, sizeof *b);
    mapped m;
    if (map_file(path, &m)) return fail(err, errlen, CBCH_ERR_IO, "cannot open %s", path);
    const uint8_t *p = m.p, *end = m.p + m.n;
    int rc = CBCH_OK;
    uint32_t last_chr = 0; const uint8_t *last_name = NULL; size_t last_name_len = 0;
    uint64_t records = 0; uint32_t second_len = 0, first_len = 0;
    if (reserve_reads(b, 4096)) rc = CBCH_ERR_NOMEM;
    uint64_t so = 0, co = 0, mo = 0;
    while (!rc && p < end) {
        const uint8_t *nl = memchr(p, '\n', (size_t)(end - p));
        const uint8_t *le = nl ? nl : end;
        const uint8_t *line = p;
        p = nl ? nl + 1 : end;
        if (le > line && le[-1] == '\r') le--;
        b->n_lines++;
        if (le == line || *line == '@') continue;                 /* header (get_read_length skips them, :40-46) */
        /* the 11 mandatory fields */
        const uint8_t *f[12]; int nf = 0;
        const uint8_t *s = line;
        while (nf < 11) {
            const uint8_t *t = memchr(s, '\t', (size_t)(le - s));
            f[nf++] = s;
            if (!t) { s = le + 1; break; }
            s = t + 1;
        }
        if (nf < 11) { rc = fail(err, errlen, CBCH_ERR_PARSE, "line %llu: fewer than 11 fields", (unsigned long long)b->n_lines); break; }
        f[11] = s;                                                /* start of the optional fields (or le + 1) */
#define FEND(i) ((i) < 10 ? f[(i) + 1] - 1 : (f[11] > le ? le : f[11] - 1))
        uint32_t flag, pos;
        if (parse_u32(f[1], FEND(1), &flag) || flag > 0xffffu || parse_u32(f[3], FEND(3), &pos)) {
            rc = fail(err, errlen, CBCH_ERR_PARSE, "line %llu: bad FLAG or POS", (unsigned long long)b->n_lines); break;
        }
        const uint32_t seqlen = (uint32_t)(FEND(9) - f[9]);
        records++;
        if (records == 1) first_len = seqlen;
        if (records == 2) second_len = seqlen;
        if (flag & 4u) { b->n_unmapped++; continue; }             /* src/compression.c:50 */
        if (seqlen > 0xffffu) { rc = fail(err, errlen, CBCH_ERR_PARSE, "line %llu: SEQ too long", (unsigned long long)b->n_lines); break; }
        /* RNAME -> ordinal (consecutive records mostly share it) */
        const uint8_t *rn = f[2]; size_t rl = (size_t)(FEND(2) - f[2]);
        uint32_t chr = last_chr;
        if (!(last_name && rl == last_name_len && !memcmp(rn, last_name, rl))) {
            uint32_t c;
            for (c = 0; c < fa->n; c++) if (strlen(fa->names[c]) == rl && !memcmp(fa->names[c], rn, rl)) break;
            if (c == fa->n) { rc = fail(err, errlen, CBCH_ERR_RNAME, "line %llu: RNAME %.*s is not in the reference", (unsigned long long)b->n_lines, (int)rl, rn); break; }
            chr = c; last_chr = c; last_name = rn; last_name_len = rl;
        }
        /* MD:Z among the optional fields */
        const uint8_t *md = NULL; size_t mdl = 0;
        for (const uint8_t *o = f[11]; o < le;) {
            const uint8_t *t = memchr(o, '\t', (size_t)(le - o));
            const uint8_t *oe = t ? t : le;
            if (oe - o >= 5 && o[0] == 'M' && o[1] == 'D' && o[2] == ':' && o[3] == 'Z' && o[4] == ':') { md = o + 5; mdl = (size_t)(oe - md); break; }
            o = oe + 1;
        }
        if (!md) { rc = fail(err, errlen, CBCH_ERR_NO_MD, "line %llu: no MD:Z tag (README.md:25-29 requires it)", (unsigned long long)b->n_lines); break; }
        const size_t cgl = (size_t)(FEND(5) - f[5]);
        if (reserve_reads(b, b->n_reads + 1) || grow((void **)&b->seq, &b->seq_cap, so + seqlen + 64, 1) ||
            grow((void **)&b->cigar, &b->cigar_cap, co + cgl + 64, 1) || grow((void **)&b->md, &b->md_cap, mo + mdl + 64, 1)) { rc = CBCH_ERR_NOMEM; break; }
        const uint64_t r = b->n_reads++;
        b->pos[r] = pos; b->flag[r] = (uint16_t)flag; b->seq_len[r] = (uint16_t)seqlen; b->chr[r] = chr;
        b->seq_off[r] = so; memcpy(b->seq + so, f[9], seqlen); so += seqlen;
        b->cigar_off[r] = co; memcpy(b->cigar + co, f[5], cgl); co += cgl;
        b->md_off[r] = mo; memcpy(b->md + mo, md, mdl); mo += mdl;
        if (seqlen > b->max_len) b->max_len = seqlen;
    }
    unmap_file(&m);
    if (rc == CBCH_ERR_NOMEM) fail(err, errlen, rc, "out of memory reading %s", path);
    if (rc) { cbch_free_batch(b); return rc; }
    if (b->n_reads == 0) {                                    /* keep the arrays addressable */
        if (reserve_reads(b, 1) || grow((void **)&b->seq, &b->seq_cap, 64, 1) || grow((void **)&b->cigar, &b->cigar_cap, 64, 1) ||
            grow((void **)&b->md, &b->md_cap, 64, 1)) return fail(err, errlen, CBCH_ERR_NOMEM, "out of memory");
    }
    b->seq_off[b->n_reads] = so; b->cigar_off[b->n_reads] = co; b->md_off[b->n_reads] = mo;
    /* get_read_length (:47-53): fixed-length mode takes the SECOND record's SEQ length; -l takes the maximum */
    b->read_len_header = var_length ? b->max_len : (records >= 2 ? second_len : first_len);
    return CBCH_OK;
}

void cbch_free_batch(cbch_batch *b) {
    free(b->pos); free(b->flag); free(b->seq_len); free(b->chr); free(b->seq_off); free(b->cigar_off); free(b->md_off);
    free(b->seq); free(b->cigar); free(b->md);
    memset(b, 0, sizeof *b);
}

void cbch_batch_view(const cbch_batch *b, cbcg_batch *v) {
    v->n_reads = b->n_reads; v->pos = b->pos; v->flag = b->flag; v->seq_len = b->seq_len; v->chr = b->chr;
    v->seq_off = b->seq_off; v->seq = b->seq; v->cigar_off = b->cigar_off; v->cigar = b->cigar; v->md_off = b->md_off; v->md = b->md;
}
