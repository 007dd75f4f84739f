/*
 * test_k2_roles.cpp -- the substream coder of blocked containers (cbc_b200/csrc/k2_roles.cuh), compiled for the HOST
 * and held against the CPU oracle: the same scalar code the GPU threads run, here one (block, substream) at a time.
 * Input: a synthetic batch (csrc/host/synth.c); edit records from the oracle's extraction (K1 has its own GPU tests).
 *   1. encode every block with the four roles, frame the container exactly as api.cu does, compare with the oracle's
 *      container byte for byte (gen_mode 0: every block from the initial snapshot; gen_mode 1: generations with a host
 *      restatement of the merge kernels over the same memory layouts);
 *   2. decode that container with the roles in their three phases and compare records and edits with the input's.
 * TEST CODE: links the oracle; nothing here ships.
 *   usage: test_k2_roles <seed> <n_reads> <genome_len> <len_min> <len_max> <p_sub> <p_indel> <p_clip> <gen_mode> [block_reads]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>
typedef struct CUstream_st *cudaStream_t;
typedef struct CUevent_st *cudaEvent_t;
#include "k2_roles.cuh"
#include "container.h"
extern "C" {
#include "synth.h"
#include "cbc_oracle.h"
}

static uint64_t al(uint64_t x) { return (x + 255u) & ~255ull; }

/* ---- host restatement of launch_merge (k2_coder.cu) over the same layouts */
static void finish_dense(const uint32_t *prev, uint32_t *m, uint32_t card, uint32_t implicit_ones) {
    uint32_t n = implicit_ones;
    for (uint32_t i = 0; i < card; i++) { int32_t v = (int32_t)m[i]; const int32_t fl = (prev && prev[i] == 0u) ? 0 : 1; if (v < fl) v = fl; m[i] = (uint32_t)v; n += (uint32_t)v; }
    while (n >= CBCG_RESCALE) { n = implicit_ones; for (uint32_t i = 0; i < card; i++) { m[i] = (m[i] >> 1) + 1u; n += m[i]; } }
    m[card] = n;
}
static int host_merge(const BlockDesc *blocks, uint32_t b0, uint32_t nb, uint32_t L, const uint8_t *prev, uint8_t *next, const uint8_t *fin,
                      const uint8_t *ws, uint32_t flag_target) {
    const SnapLayout l = snap_layout(L);
    memcpy(next, prev, l.total);
    const WarpModels *pm = (const WarpModels *)(prev + l.small);
    WarpModels *nm = (WarpModels *)(next + l.small);
    std::vector<uint32_t> fprev(65536, 1u), facc;
    for (uint32_t j = 0; j < pm->flag_used; j++) fprev[pm->flag_key[j]] = pm->flag_cnt[j];
    facc = fprev;
    uint32_t *bm = (uint32_t *)(next + l.bitmap), *var = (uint32_t *)(next + l.var);
    const uint32_t *pbm = (const uint32_t *)(prev + l.bitmap), *pvar = (const uint32_t *)(prev + l.var);
    const uint32_t pc = ((const uint32_t *)(prev + l.pos_hdr))[0];
    const uint32_t *pcnt = (const uint32_t *)(prev + l.pos_cnt);
    uint32_t *ncnt = (uint32_t *)(next + l.pos_cnt), *nval = (uint32_t *)(next + l.pos_val);
    std::vector<uint32_t> newv, newc;
    const uint32_t w0 = offsetof(WarpModels, snps) / 4, w1 = offsetof(WarpModels, flag_key) / 4;
    for (uint32_t k = 0; k < nb; k++) {
        const BlockDesc &B = blocks[b0 + k];
        const WsLayout w = ws_layout(L, B.n_reads, B.n_edits, 0, 1);
        const uint8_t *wsb = ws + B.ws_off;
        const uint32_t *fs = (const uint32_t *)(fin + (uint64_t)(b0 + k) * fin_stride_dev());
        const uint32_t *ps = (const uint32_t *)pm; uint32_t *ns = (uint32_t *)nm;
        for (uint32_t i = w0; i < w1; i++) ns[i] += fs[i] - ps[i];
        const WarpModels *fm = (const WarpModels *)fs;
        for (uint32_t j = 0; j < fm->flag_used; j++) { const uint32_t key = fm->flag_key[j] & 0xffffu; facc[key] += fm->flag_cnt[j] - fprev[key]; }
        const uint32_t *bcnt = (const uint32_t *)(wsb + w.pos_cnt), *bval = (const uint32_t *)(wsb + w.pos_val);
        for (uint32_t s = 0; s < pc; s++) ncnt[s] += bcnt[s] - pcnt[s];
        for (uint32_t s = pc; s < B.pos_card; s++) {
            size_t q = 0;
            for (; q < newv.size(); q++) if (newv[q] == bval[s]) break;
            if (q < newv.size()) newc[q] += bcnt[s];
            else if (pc + newv.size() < CBCG_SNAP_POS_MAX) { newv.push_back(bval[s]); newc.push_back(bcnt[s]); }
        }
        if (B.pa_touched) {
            const uint32_t *pa_prev = (const uint32_t *)(prev + l.pos_alpha), *pa_blk = (const uint32_t *)(wsb + w.pos_alpha);
            uint32_t *pa_next = (uint32_t *)(next + l.pos_alpha);
            for (uint32_t i = 0; i < 4u * PA_STRIDE; i++) pa_next[i] += pa_blk[i] - pa_prev[i];
        }
        const uint64_t *hash = (const uint64_t *)(wsb + w.var_hash);
        const uint32_t *rows = (const uint32_t *)(wsb + w.var_rows);
        for (uint32_t h = 0; h < w.hash_cap; h++) {
            const uint64_t e = hash[h];
            if (!(uint32_t)(e >> 32)) continue;
            const uint32_t ctx = (uint32_t)(e >> 32) - 1u, r = (uint32_t)e;
            uint32_t *nrow = var + (uint64_t)ctx * l.Lp;
            const bool in_prev = (pbm[ctx >> 5] >> (ctx & 31u)) & 1u;
            if (!((bm[ctx >> 5] >> (ctx & 31u)) & 1u)) { bm[ctx >> 5] |= 1u << (ctx & 31u); for (uint32_t i = 0; i < L; i++) nrow[i] = 1u; nrow[L] = L; }
            if (r & VAR_DEFERRED) { nrow[r & 0xffffu] += 10u; continue; }
            const uint32_t *row = rows + (uint64_t)r * w.Lp, *prow = pvar + (uint64_t)ctx * l.Lp;
            for (uint32_t i = 0; i < L; i++) nrow[i] += row[i] - (in_prev ? prow[i] : 1u);
        }
    }
    const uint32_t *ps = (const uint32_t *)pm; uint32_t *ns = (uint32_t *)nm;
#define W(f) (offsetof(WarpModels, f) / 4)
    finish_dense(ps + W(snps), ns + W(snps), L, 0); finish_dense(ps + W(indels), ns + W(indels), L, 0);
    finish_dense(ps + W(rlen0), ns + W(rlen0), 255, 0);
    for (uint32_t r = 0; r < 6; r++) finish_dense(ps + W(chars) + r * 8, ns + W(chars) + r * 8, 5, 0);
    for (uint32_t r = 0; r < 4; r++) finish_dense(ps + W(match) + r * 4, ns + W(match) + r * 4, 2, 0);
    finish_dense(ps + W(same_ref), ns + W(same_ref), 2, 0);
    for (uint32_t r = 0; r < 3; r++) finish_dense(ps + W(rlenk) + r * 2, ns + W(rlenk) + r * 2, 1, 254);
    for (uint32_t k = 0; k < 4; k++) finish_dense((const uint32_t *)(prev + l.pos_alpha) + k * PA_STRIDE, (uint32_t *)(next + l.pos_alpha) + k * PA_STRIDE, 256, 0);
    {   /* POS */
        const uint32_t an = pc + (uint32_t)newv.size();
        for (size_t q = 0; q < newv.size(); q++) { nval[pc + q] = newv[q]; ncnt[pc + q] = newc[q]; }
        uint32_t n = 0;
        for (uint32_t i = 0; i < an; i++) { int32_t v = (int32_t)ncnt[i]; if (v < 1) v = 1; ncnt[i] = (uint32_t)v; n += (uint32_t)v; }
        while (n >= CBCG_RESCALE) { n = 0; for (uint32_t i = 0; i < an; i++) { ncnt[i] = (ncnt[i] >> 1) + 1u; n += ncnt[i]; } }
        uint32_t *hdr = (uint32_t *)(next + l.pos_hdr); hdr[0] = an; hdr[1] = n;
    }
    {   /* FLAG: clamp, scale to the target total (cbcg_flag_target), back to the sorted sparse form */
        uint64_t n = 0;
        for (uint32_t i = 0; i < 65536; i++) { int32_t v = (int32_t)facc[i]; if (v < 1) v = 1; facc[i] = (uint32_t)v; n += (uint32_t)v; }
        if (n > flag_target) {
            const uint64_t a = flag_target - 65536u; uint64_t s = 0;
            for (uint32_t i = 0; i < 65536; i++) { uint64_t c = (uint64_t)facc[i] * a / n; if (c < 1) c = 1; facc[i] = (uint32_t)c; s += c; }
            n = s;
        }
        uint32_t used = 0;
        for (uint32_t i = 0; i < 65536; i++) used += facc[i] != 1u;
        if (used > FLAG_CAP) {                                 /* rule F2 (cbcg_format.h): the counts <= T go back to 1 */
            uint32_t lo = 1u, hi = 0x7fffffffu;
            while (lo < hi) {
                const uint32_t mid = lo + (hi - lo) / 2u; uint32_t above = 0;
                for (uint32_t i = 0; i < 65536; i++) above += facc[i] > mid;
                if (above <= FLAG_CAP) hi = mid; else lo = mid + 1u;
            }
            n = 0;
            for (uint32_t i = 0; i < 65536; i++) { if (facc[i] <= lo) facc[i] = 1u; n += facc[i]; }
        }
        used = 0;
        for (uint32_t i = 0; i < 65536; i++) if (facc[i] != 1u) { if (used >= FLAG_CAP) return CBCG_ERR_LIMIT; nm->flag_key[used] = i; nm->flag_cnt[used] = facc[i]; used++; }
        nm->flag_used = used; nm->flag_n = (uint32_t)n;
    }
    for (uint32_t ctx = 0; ctx < CBCG_VAR_CONTEXTS; ctx++) {
        if (!((bm[ctx >> 5] >> (ctx & 31u)) & 1u)) continue;
        finish_dense(nullptr, var + (uint64_t)ctx * l.Lp, L, 0);
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 10) { fprintf(stderr, "usage: see the header\n"); return 2; }
    cbcs_params sp; memset(&sp, 0, sizeof sp);
    sp.seed = strtoull(argv[1], 0, 10); sp.n_reads = strtoull(argv[2], 0, 10); sp.n_chr = 1;
    const uint64_t glen = strtoull(argv[3], 0, 10);
    sp.len_min = atoi(argv[4]); sp.len_max = atoi(argv[5]); sp.p_sub = atof(argv[6]); sp.p_indel = atof(argv[7]); sp.p_clip = atof(argv[8]);
    sp.p_rev = 0.5; sp.avoid_b3 = 1; sp.flag_mode = sp.seed & 1;
    const uint32_t gen_mode = atoi(argv[9]);
    uint32_t block_reads = argc > 10 ? (uint32_t)strtoul(argv[10], 0, 10) : 0xffffffffu;
    const uint64_t n = sp.n_reads; const uint32_t L = sp.len_max;
    if (argc > 11) sp.n_chr = atoi(argv[11]);
    std::vector<std::vector<uint8_t>> chrs(sp.n_chr);
    std::vector<const uint8_t *> cptr(sp.n_chr); std::vector<uint64_t> clen(sp.n_chr); std::vector<std::string> names(sp.n_chr); std::vector<const char *> nptr(sp.n_chr);
    for (uint32_t c = 0; c < sp.n_chr; c++) { chrs[c].resize(glen / sp.n_chr); cbcs_genome(sp.seed, c, chrs[c].data(), chrs[c].size()); cptr[c] = chrs[c].data(); clen[c] = chrs[c].size(); names[c] = "chr" + std::to_string(c + 1); nptr[c] = names[c].c_str(); }
    cbcs_out o; memset(&o, 0, sizeof o);
    std::vector<uint32_t> pos(n), chr(n); std::vector<uint16_t> flag(n), slen(n);
    std::vector<uint64_t> so(n + 1), co(n + 1), mo(n + 1);
    std::vector<uint8_t> seq(n * sp.len_max + 64), cig(n * 64 + 4096), md(n * 96 + 4096);
    o.reads_cap = n; o.pos = pos.data(); o.flag = flag.data(); o.seq_len = slen.data(); o.chr = chr.data();
    o.seq_off = so.data(); o.seq = seq.data(); o.seq_cap = seq.size(); o.cigar_off = co.data(); o.cigar = cig.data(); o.cigar_cap = cig.size();
    o.md_off = mo.data(); o.md = md.data(); o.md_cap = md.size();
    if (cbcs_reads(&sp, cptr.data(), clen.data(), &o)) { fprintf(stderr, "generator failed\n"); return 1; }
    if (argc > 12) {                                           /* n_flags distinct FLAG values (strand bit included); argv[13] = 1: drawn at random from the first read on
                                                                  (otherwise the first n_flags reads carry one each, which fills the first snapshot at once) */
        const int random_head = argc > 13 && atoi(argv[13]);
        const uint32_t nf = (uint32_t)atoi(argv[12]);
        uint64_t x = sp.seed * 0x9e3779b97f4a7c15ull + 1u;
        for (uint64_t r = 0; r < n; r++) {
            x = x * 6364136223846793005ull + 1442695040888963407ull;
            const uint32_t k = (r < nf && !random_head) ? (uint32_t)r : (uint32_t)((x >> 33) % nf);
            flag[r] = (uint16_t)((k * 13u + 5u) % 4096u);        /* 13 is odd: distinct k < 4096 give distinct values */
        }
    }
    cbco_batch ob = { n, pos.data(), flag.data(), slen.data(), chr.data(), so.data(), seq.data(), co.data(), cig.data(), mo.data(), md.data() };
    cbco_genome og = { sp.n_chr, cptr.data(), clen.data(), nptr.data() };
    std::vector<cbcg_read_rec> recs(n + 1); std::vector<uint16_t> edits(3 * seq.size() + 64);
    const int64_t ne = cbco_extract(&ob, &og, recs.data(), edits.data(), edits.size());
    if (ne < 0) { fprintf(stderr, "oracle extract failed\n"); return 1; }

    /* ---- the cut (api.cu cut_blocks): generations on the schedule, never across a chromosome change */
    uint32_t sc[CBCG_GEN_MAX], sr[CBCG_GEN_MAX], last = 0, levels = 0;
    if (gen_mode) levels = cbcg_gen_schedule(n, CBCG_N_SUB, sc, sr, &last, nullptr);
    if (block_reads == 0xffffffffu) block_reads = gen_mode ? last : 1024u;
    std::vector<BlockDesc> hb;
    {
        uint32_t gen = 0, left = levels ? sc[0] : 0;
        uint64_t r = 0;
        while (r < n) {
            uint32_t want = gen < levels ? sr[gen] : block_reads;
            uint64_t e = r + 1;
            while (e < n && e - r < want && chr[e] == chr[r]) e++;
            BlockDesc d; memset(&d, 0, sizeof d);
            d.first_read = (uint32_t)r; d.n_reads = (uint32_t)(e - r); d.chr = chr[r]; d.gen = gen; d.base_pos = recs[r].pos;
            d.edit_base = recs[r].edit_off; d.n_edits = (uint32_t)((e < n ? recs[e].edit_off : (uint64_t)ne) - recs[r].edit_off);
            hb.push_back(d); r = e;
            if (gen < levels && --left == 0) { gen++; left = gen < levels ? sc[gen] : 0; }
        }
    }
    const uint32_t nb = (uint32_t)hb.size();
    bool fixed = true; uint32_t max_len = 0, max_block = 0;
    for (uint64_t r = 0; r < n; r++) { if (slen[r] != L) fixed = false; max_len = std::max<uint32_t>(max_len, slen[r]); }
    uint64_t ws_total = 0, pay_total = 0;
    for (auto &d : hb) { d.ws_off = ws_total; ws_total += ws_layout(L, d.n_reads, d.n_edits, 0, 1).total; d.payload_off = pay_total; pay_total += payload_cap_bytes(d.n_reads, d.n_edits, 0); max_block = std::max(max_block, d.n_reads); }
    std::vector<uint8_t> ws(ws_total + 256), scratch(pay_total + 256), fin((uint64_t)nb * fin_stride_dev() + 256);
    const SnapLayout sl = snap_layout(L);
    std::vector<uint8_t> snap_a(al(sl.total)), snap_b(al(sl.total));
    k2_snapshot_init(snap_a.data(), L);
    DevGenome dg; std::vector<uint64_t> goff(sp.n_chr), glen2(sp.n_chr); std::vector<uint8_t> gb;
    for (uint32_t c = 0; c < sp.n_chr; c++) { goff[c] = gb.size(); glen2[c] = clen[c]; gb.insert(gb.end(), chrs[c].begin(), chrs[c].end()); gb.resize(gb.size() + 1024, 'N'); }
    dg.n_chr = sp.n_chr; dg.bases = gb.data(); dg.chr_off = goff.data(); dg.chr_len = glen2.data();
    CoderParams P; memset(&P, 0, sizeof P);
    P.L = L; P.blocks = hb.data(); P.recs = recs.data(); P.edits = edits.data(); P.chr = chr.data(); P.genome = dg;
    P.ws = ws.data(); P.payload = scratch.data(); P.lean = 1; P.short_flush = 1; P.primed = 1; P.fixed_len = fixed ? 1 : 0; P.fin = fin.data();
    const uint32_t flag_target = cbcg_flag_target(max_block);

    auto run_generations = [&](int mode) -> int {
        uint8_t *cur = snap_a.data(), *other = snap_b.data();
        k2_snapshot_init(cur, L);
        uint32_t g0 = 0;
        while (g0 < nb) {
            uint32_t g1 = g0; while (g1 < nb && hb[g1].gen == hb[g0].gen) g1++;
            P.snap = cur;
            const uint32_t phases[3][2] = { { 0, 2 }, { 2, 3 }, { 3, 4 } };     /* decode: {A, B} -> C -> D; encode: any order */
            for (int ph = 0; ph < 3; ph++)
                for (uint32_t k = g0; k < g1; k++)
                    for (uint32_t q = phases[ph][0]; q < phases[ph][1]; q++) {
                        uint64_t item = 0;
                        const int rc = mode == MODE_ENC ? k2_run_role<MODE_ENC>(P, q, k, &item) : k2_run_role<MODE_DEC>(P, q, k, &item);
                        if (rc) { fprintf(stderr, "role %u of block %u failed: %d (item %llu)\n", q, k, rc, (unsigned long long)item); return rc; }
                    }
            if (g1 < nb) {
                const int rc = host_merge(hb.data(), g0, g1 - g0, L, cur, other, fin.data(), ws.data(), flag_target);
                if (rc) { fprintf(stderr, "merge failed: %d\n", rc); return rc; }
                std::swap(cur, other);
            }
            g0 = g1;
        }
        return 0;
    };

    /* ---- 1. encode, frame, compare with the oracle's container */
    if (run_generations(MODE_ENC)) return 1;
    std::vector<uint8_t> head, cont;
    container_head(head, max_len, L, n, nb, names, block_reads, gen_mode | CBCG_MODE_SPLIT4 | (fixed ? CBCG_MODE_FIXED_LEN : 0u), hb.data());
    cont = head;
    for (auto &d : hb) {
        uint64_t off = d.payload_off;
        d.payload_bytes = 0;
        for (uint32_t q = 0; q < CBCG_N_SUB; q++) { cont.insert(cont.end(), scratch.begin() + off, scratch.begin() + off + d.sub_bytes[q]); off += k2_sub_cap(q, d.n_reads, d.n_edits); d.payload_bytes += d.sub_bytes[q]; }
    }
    cbco_buf ref = { 0, 0, 0 };
    const int orc = gen_mode ? cbco_encode_blocked(&ob, &og, L, argc > 10 ? block_reads : 0xffffffffu, gen_mode | CBCG_MODE_SPLIT4, &ref)
                             : cbco_encode_blocked(&ob, &og, L, block_reads, CBCG_MODE_SPLIT4, &ref);
    if (orc) { fprintf(stderr, "oracle encode failed: %d\n", orc); return 1; }
    if (ref.size != cont.size() || memcmp(ref.data, cont.data(), cont.size())) {
        size_t d = 0; while (d < cont.size() && d < ref.size && cont[d] == ref.data[d]) d++;
        fprintf(stderr, "FAIL: container != oracle (sizes %zu vs %llu, first difference at byte %zu, head %zu bytes, %u blocks)\n", cont.size(), (unsigned long long)ref.size, d, head.size(), nb);
        return 1;
    }

    /* ---- 2. decode the container with the roles: records and edits of the input come back */
    std::vector<cbcg_read_rec> drecs(n + 1); std::vector<uint16_t> dedits(edits.size(), 0xffffu); std::vector<uint32_t> dchr(n, 0xffffffffu);
    memset(drecs.data(), 0xff, drecs.size() * sizeof(cbcg_read_rec));
    {
        uint64_t off = 0;
        for (auto &d : hb) { d.payload_off = off; off += d.payload_bytes; d.n_symbols = 0; d.pos_card = d.n_rows = d.pa_touched = 0; }
        std::vector<uint8_t> payload(cont.begin() + head.size(), cont.end()); payload.resize(payload.size() + 64);
        P.payload = payload.data(); P.recs = drecs.data(); P.edits = dedits.data(); P.chr = dchr.data();
        if (run_generations(MODE_DEC)) return 1;
    }
    for (uint64_t r = 0; r < n; r++) {
        if (memcmp(&drecs[r], &recs[r], sizeof(cbcg_read_rec)) || dchr[r] != chr[r]) { fprintf(stderr, "FAIL: decoded record %llu differs\n", (unsigned long long)r); return 1; }
    }
    if (memcmp(dedits.data(), edits.data(), (size_t)ne * 2)) { fprintf(stderr, "FAIL: decoded edits differ\n"); return 1; }
    printf("ok: %llu reads, %u blocks, %zu container bytes (gen_mode %u, block_reads %u, %s)\n", (unsigned long long)n, nb, cont.size(), gen_mode, block_reads, fixed ? "fixed length" : "variable length");
    cbco_buf_free(&ref);
    return 0;
}
