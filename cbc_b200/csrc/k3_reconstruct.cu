/*
 * k3_reconstruct.cu -- K3: read reconstruction (hot path part 3).
 *
 * Replaces reconstruct_read's base-building half (src/read_decompression.c:383-384 perfect match,
 * :418-437 deletions and fill, :442-458 SNP substitution, :464-515 insertions) followed by
 * print_line (src/compression.c:16-40): reverse-strand reads are reverse-complemented by the
 * first and again by the second, so the emitted line is the forward construction for both
 * strands. Output is SEQ + '\n' per read, reads in input order.
 *
 * B200 design. One CTA per tile of K3_TILE reads; the reference window under the tile is staged by a
 * 1-D TMA bulk copy; the tile's output bytes are assembled in shared memory (warp per read, one lane
 * per output byte) and leave with one TMA bulk store (16-byte aligned middle) plus byte stores for the
 * ragged ends. Output offsets come from a decoupled look-back over tile byte totals, so variable-length
 * reads need no separate scan pass.
 * Algorithmic HBM bytes per read: 16 (record) + 2*edits + 4 (chr) + L/coverage (window) + L + 1 (output).
 */
#include <algorithm>
#include "common.cuh"
#include "internal.h"

#define K3_TILE     128u
#define K3_WARPS    8u
#define K3_THREADS  (K3_WARPS * 32u)
#define K3_REF_CAP  8192u         /* most bytes of reference window staged per tile */
#define K3_CHUNKS   ((K3_TILE * (CBCG_MAX_READ_LEN + 1u) + 48u) / 16u)   /* 16-byte pieces of the largest tile image */

uint64_t reconstruct_num_tiles(uint64_t n_reads) { return (n_reads + K3_TILE - 1) / K3_TILE; }

/* Scratch of a read that is built base by base: edit positions. Every warp has room for K3_EDITS_SMALL edits of a kind
 * (nearly every read); a read with more of one kind waits for the CTA's one full-size scratch (255 of a kind, the
 * format's limit), which warp 0 works through after the others are done: per-warp full-size scratch was 16 KB of a
 * 49 KB CTA and cost two resident CTAs per SM. */
#define K3_EDITS_SMALL 64u
template <uint32_t CAP> struct K3Scratch {
    uint16_t cumdel[CAP];             /* cumulative deletion offsets */
    uint16_t snp_at[CAP];             /* aligned index of each SNP */
    uint16_t ins_at[CAP];             /* output index of each inserted base */
    uint8_t  snp_ch[CAP];
    uint8_t  ins_ch[CAP];
};
typedef K3Scratch<K3_EDITS_SMALL> K3Warp;
typedef K3Scratch<256u> K3Big;

struct K3Smem {
    uint64_t bar;
    uint64_t tile_base;
    uint32_t tile;
    uint32_t n_slow;
    uint32_t warp_tot[K3_WARPS];
    uint32_t out_off[K3_TILE + 1];    /* byte offset of each read's line inside the tile; [nr] = tile total */
    __align__(16) cbcg_read_rec rec[K3_TILE];
    uint32_t chr[K3_TILE];
    uint32_t so[K3_TILE + 1];         /* offset of the read's first base in ref[] (reads on the copy path), else 0 */
    uint8_t  slow[K3_TILE];           /* reads built base by base (indels, reads outside the window) */
    uint8_t  is_slow[K3_TILE];
    uint8_t  big[K3_TILE];            /* reads with more than K3_EDITS_SMALL edits of a kind */
    uint32_t n_big;
    K3Warp w[K3_WARPS];
    K3Big wbig;
    __align__(16) uint8_t dyn[16];    /* 32 bytes of front pad (pieces at a tile's edge read up to 16 bytes before a read's first base),
                                         ref_cap + 64 bytes of reference window, K3_TILE * (max_len + 1) + 48 bytes of tile image,
                                         one owner byte per 16-byte piece of the image (dynamic) */
};

/* 16 bytes of the staged window from byte offset `off` (any alignment, >= -32): two aligned 16-byte loads, the
 * word rotation as selects (registers cannot be indexed), the byte rotation as funnel shifts. */
__device__ __forceinline__ uint4 lds_16_unaligned(const uint8_t *base, int32_t off) {
    const uint4 q0 = *reinterpret_cast<const uint4 *>(base + (off & ~15));
    const uint4 q1 = *reinterpret_cast<const uint4 *>(base + (off & ~15) + 16);
    const bool r2 = (off & 8) != 0, r1 = (off & 4) != 0;
    const uint32_t x0 = r2 ? q0.z : q0.x, x1 = r2 ? q0.w : q0.y, x2 = r2 ? q1.x : q0.z, x3 = r2 ? q1.y : q0.w,
                   x4 = r2 ? q1.z : q1.x, x5 = r2 ? q1.w : q1.y;
    const uint32_t y0 = r1 ? x1 : x0, y1 = r1 ? x2 : x1, y2 = r1 ? x3 : x2, y3 = r1 ? x4 : x3, y4 = r1 ? x5 : x4;
    const uint32_t sh = ((uint32_t)off & 3u) * 8u;
    return make_uint4(__funnelshift_r(y0, y1, sh), __funnelshift_r(y1, y2, sh), __funnelshift_r(y2, y3, sh), __funnelshift_r(y3, y4, sh));
}
/* A piece that holds a line end: bytes below `rem` from a (the read that ends), byte rem = '\n', bytes above from b (the
 * next read). Only the word with the '\n' mixes the two sources; the words before it are a's, the ones after it b's. */
__device__ __forceinline__ uint4 k3_merge(const uint4 a, const uint4 b, uint32_t rem) {       /* rem in 0..15 */
    const uint32_t kk = rem >> 2, sh = (rem & 3u) * 8u;
    const uint32_t as = kk == 0u ? a.x : (kk == 1u ? a.y : (kk == 2u ? a.z : a.w));
    const uint32_t bs = kk == 0u ? b.x : (kk == 1u ? b.y : (kk == 2u ? b.z : b.w));
    const uint32_t lo = (1u << sh) - 1u;                                                    /* bytes below the '\n' */
    const uint32_t mix = (as & lo) | ((uint32_t)'\n' << sh) | (bs & ~(lo | (0xffu << sh)));
    uint4 v;
    v.x = kk == 0u ? mix : a.x;                                                             /* word 0 is never after the '\n' */
    v.y = kk == 1u ? mix : (kk > 1u ? a.y : b.y);
    v.z = kk == 2u ? mix : (kk > 2u ? a.z : b.z);
    v.w = kk == 3u ? mix : b.w;
    return v;
}

__global__ void __launch_bounds__(K3_THREADS, 6)
k3_reconstruct_kernel(uint64_t n_reads, const cbcg_read_rec *__restrict__ recs, const uint32_t *__restrict__ chr_of,
                      const uint16_t *__restrict__ edits, DevGenome g, uint8_t *__restrict__ out, uint64_t out_cap,
                      uint64_t *tile_desc, uint32_t *ticket, uint64_t *total_bytes, unsigned long long *err,
                      uint32_t max_len, uint32_t fixed_len, uint32_t ref_cap) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    K3Smem &S = *reinterpret_cast<K3Smem *>(smem_raw);
    uint8_t *const s_ref = S.dyn + 32u, *const s_out = s_ref + ref_cap + 64u;
    uint8_t *const s_own = s_out + ((K3_TILE * (max_len + 1u) + 48u + 15u) & ~15u);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    /* Tiles take their index from a ticket so that the look-back below cannot deadlock; with closed-form offsets
       (fixed_len) the CTA index will do, and the records are requested one barrier earlier. */
    uint32_t tile = blockIdx.x;
    if (!fixed_len) {
        if (tid == 0) S.tile = atomicAdd(ticket, 1u);
        __syncthreads();
        tile = S.tile;
    }
    if (tid == 0) {
        S.n_slow = 0; S.n_big = 0;
        mbar_init(&S.bar, 1);
        mbar_fence_init();
    }
    const uint64_t r0 = (uint64_t)tile * K3_TILE;
    const uint32_t nr = (uint32_t)min((uint64_t)K3_TILE, n_reads - r0);

    /* records (16 B each, coalesced) and chromosome ordinals */
    uint32_t my_bytes = 0;
    if (tid < nr) {
        uint4 v = reinterpret_cast<const uint4 *>(recs)[r0 + tid];
        *reinterpret_cast<uint4 *>(&S.rec[tid]) = v;
        S.chr[tid] = chr_of[r0 + tid];
        uint32_t len = S.rec[tid].len;
        if (len > max_len || len == 0 || (fixed_len && len != fixed_len)) {
            dev_set_error(err, CBCG_ERR_CORRUPT, r0 + tid);
            len = fixed_len; S.rec[tid].len = (uint16_t)len; S.rec[tid].pos = 0u;     /* pos 0: rebuilt as N's below */
        }
        my_bytes = len + 1u;
    }
    /* the copy path needs every line of the tile to span a whole 16-byte piece */
    const int tile_short = __syncthreads_or(tid < nr && my_bytes < 17u);

    /* reference window under the tile's first read */
    const uint32_t chr0 = S.chr[0], pos0 = S.rec[0].pos;
    uint64_t w0 = 0; uint32_t ref_bytes = 0;
    if (chr0 < g.n_chr) {
        const uint64_t clen0 = g.chr_len[chr0];
        w0 = pos0 ? ((uint64_t)(pos0 - 1u) & ~15ull) : 0ull;
        uint64_t avail = (clen0 + REF_PAD > w0) ? ((clen0 + REF_PAD - w0) & ~15ull) : 0ull;
        ref_bytes = (uint32_t)min((uint64_t)ref_cap, avail);
        /* position-sorted input: the tile's last read bounds the window (a read beyond it misses the window and is
           built from HBM) */
        const uint32_t pos_l = S.rec[nr - 1u].pos;
        if (S.chr[nr - 1u] == chr0 && pos_l >= pos0 && pos0) {
            const uint64_t need = ((uint64_t)(pos_l - 1u) - w0 + max_len + 8u + 15u) & ~15ull;
            if (need < ref_bytes) ref_bytes = (uint32_t)need;
        }
    }
    if (tid == 0) {
        mbar_expect_tx(&S.bar, ref_bytes);
        if (ref_bytes) tma_load_1d(s_ref, g.bases + g.chr_off[chr0] + w0, ref_bytes, &S.bar);
    }

    /* Which reads are plain copies of the window, possibly with a few substituted bases? (perfect matches :383-384
       and substitution-only reads :442-458: all of config 2.) They take the copy path; the others are listed. The
       first SNPs of a copy-path read are fetched now, so that their latency hides behind the window's. */
    bool fast = false; uint32_t ns = 0, e_first[4] = { 0u, 0u, 0u, 0u };
    const uint16_t *e_mine = edits;
    if (tid < nr) {
        const cbcg_read_rec &rec = S.rec[tid];
        const uint32_t len = rec.len, pos = rec.pos, chr = S.chr[tid];
        ns = rec.match ? 0u : rec.n_snps;
        bool ok = !tile_short && len > 0u && pos >= 1u && chr == chr0 && chr < g.n_chr && ref_bytes != 0u &&
                  (rec.match || (rec.n_dels == 0u && rec.n_ins == 0u));
        if (ok) ok = (uint64_t)(pos - 1u) + len <= g.chr_len[chr] && (uint64_t)(pos - 1u) >= w0 &&
                     ((uint64_t)(pos - 1u) - w0 + len + 8u <= ref_bytes);
        fast = ok;
        S.so[tid] = ok ? (uint32_t)((uint64_t)(pos - 1u) - w0) : 0u;
        S.is_slow[tid] = ok ? 0u : 1u;
        if (!ok) { S.slow[atomicAdd(&S.n_slow, 1u)] = (uint8_t)tid; ns = 0; }
        else if (ns) {
            e_mine = edits + rec.edit_off;
#pragma unroll
            for (uint32_t q = 0; q < 4u; q++) if (q < ns) e_first[q] = e_mine[q];
        }
    }
    if (tid == nr) S.so[nr] = 0u;

    /* CTA scan of line sizes (threads >= K3_TILE carry 0) */
    uint32_t incl = warp_incl_scan(my_bytes);
    if (lane == 31) S.warp_tot[warp] = incl;
    __syncthreads();
    uint32_t warp_base = 0, tile_total = 0;
#pragma unroll
    for (uint32_t k = 0; k < K3_WARPS; k++) { uint32_t t = S.warp_tot[k]; if (k < warp) warp_base += t; tile_total += t; }
    const uint32_t my_off = warp_base + incl - my_bytes;
    if (tid < nr) S.out_off[tid] = my_off;
    if (tid == nr) S.out_off[nr] = tile_total;
    /* Output offset of the tile. Equal-length reads (CBCG_MODE_FIXED_LEN containers): closed form. Otherwise a
       decoupled look-back over the tiles' byte totals, whose wait for the slowest of 32 neighbouring tiles is the
       longest stall of this kernel. */
    if (fixed_len) {
        if (tid == 0 && r0 + nr == n_reads) *total_bytes = n_reads * (uint64_t)(fixed_len + 1u);
    } else if (warp == 0) {
        uint64_t base = lookback_exclusive(tile_desc, tile, tile_total, err);
        if (lane == 0) {
            S.tile_base = base;
            if (r0 + nr == n_reads) *total_bytes = base + tile_total;
        }
    }
    __syncthreads();
    const uint64_t tile_base = fixed_len ? r0 * (uint64_t)(fixed_len + 1u) : S.tile_base;
    /* smem image is shifted so that smem offset == global offset (mod 16): the middle can leave by TMA */
    const uint32_t shift = (uint32_t)((uint64_t)(out + tile_base) & 15ull);
    uint8_t *img = s_out + shift;
    const uint32_t n_chunks = (shift + tile_total + 15u) >> 4;
    /* owner of the first byte of every 16-byte piece: read i owns the pieces that start inside its line */
    if (tid < nr) {
        const uint32_t c_lo = (tid == 0u) ? 0u : ((shift + my_off + 15u) >> 4);
        const uint32_t c_hi = (shift + my_off + my_bytes + 15u) >> 4;       /* first piece of the next read */
        for (uint32_t c = c_lo; c < c_hi && c < n_chunks; c++) s_own[c] = (uint8_t)tid;
    }
    __syncthreads();
    mbar_wait(&S.bar, 0);

    /* ---- copy path, one 16-byte piece of the tile image per lane and step. A piece lies inside one line, or holds the
       end of line i and the start of line i + 1 (lines are >= 17 bytes here). Reads off the copy path get filler
       bytes that their own builder overwrites below. */
    if (!tile_short && ref_bytes) {
        for (uint32_t c = tid; c < n_chunks; c += K3_THREADS) {
            const uint32_t i = s_own[c];
            const int32_t j0 = (int32_t)(16u * c) - (int32_t)(shift + S.out_off[i]);   /* read-relative index of the piece's first byte */
            const int32_t rem = (int32_t)S.rec[i].len - j0;                             /* bases of read i from there on */
            uint4 v = lds_16_unaligned(s_ref, (int32_t)S.so[i] + j0);
            if (rem < 16) v = k3_merge(v, lds_16_unaligned(s_ref, (int32_t)S.so[i + 1u] - rem - 1), (uint32_t)rem);   /* rem >= 0: the piece starts inside line i */
            *reinterpret_cast<uint4 *>(s_out + 16u * c) = v;
        }
    }
    __syncthreads();
    /* substituted bases of the copy-path reads: SNP k sits at sum_{i<k}(p_i + 1) + p_k (:442-458) */
    if (fast && ns) {
        const uint32_t len = S.rec[tid].len;
        uint8_t *dst = img + my_off;
        uint32_t at = 0; bool bad = false;
#pragma unroll
        for (uint32_t q = 0; q < 4u; q++) {
            if (q < ns && !bad) {
                at += CBCG_EDIT_DELTA(e_first[q]) + 1u;
                if (at - 1u >= len) bad = true; else dst[at - 1u] = (uint8_t)base_char(CBCG_EDIT_TARGET(e_first[q]));
            }
        }
        for (uint32_t q = 4u; q < ns && !bad; q++) {
            const uint32_t ed = e_mine[q];
            at += CBCG_EDIT_DELTA(ed) + 1u;
            if (at - 1u >= len) bad = true; else dst[at - 1u] = (uint8_t)base_char(CBCG_EDIT_TARGET(ed));
        }
        if (bad) { dev_set_error(err, CBCG_ERR_CORRUPT, r0 + tid); for (uint32_t j = 0; j < len; j++) dst[j] = 'N'; }
    }

    /* reads built base by base, a warp each: first pass with the warps' own scratch, second pass (warp 0) for the reads
       that need the full-size one. One copy of the builder: the scratch is reached through pointers. */
    struct { uint16_t *cumdel, *snp_at, *ins_at; uint8_t *snp_ch, *ins_ch; } W;
    for (uint32_t pass = 0; pass < 2u; pass++) {
    if (pass) { if (!__syncthreads_or(S.n_big != 0u)) break; }               /* CTA-uniform: n_big is final after the barrier */
    const uint32_t n_list = pass ? S.n_big : S.n_slow, cap = pass ? 256u : K3_EDITS_SMALL;
    if (pass) { W.cumdel = S.wbig.cumdel; W.snp_at = S.wbig.snp_at; W.ins_at = S.wbig.ins_at; W.snp_ch = S.wbig.snp_ch; W.ins_ch = S.wbig.ins_ch; }
    else { K3Warp &w = S.w[warp]; W.cumdel = w.cumdel; W.snp_at = w.snp_at; W.ins_at = w.ins_at; W.snp_ch = w.snp_ch; W.ins_ch = w.ins_ch; }
    for (uint32_t sidx = pass ? (warp ? n_list : 0u) : warp; sidx < n_list; sidx += pass ? 1u : K3_WARPS) {
        const uint32_t i = pass ? S.big[sidx] : S.slow[sidx];
        const cbcg_read_rec &rec = S.rec[i];
        const uint32_t len = rec.len, pos = rec.pos, chr = S.chr[i];
        uint8_t *dst = img + S.out_off[i];
        if (len == 0) { if (lane == 0) dst[0] = '\n'; continue; }
        bool bad = (pos == 0u) || chr >= g.n_chr;
        const uint32_t nd = rec.match ? 0u : rec.n_dels, ns = rec.match ? 0u : rec.n_snps, ni = rec.match ? 0u : rec.n_ins;
        if (ni > len) bad = true;
        const uint32_t aligned = len - (bad ? 0u : ni);
        uint64_t clen = 0; const uint8_t *gref = nullptr;
        if (!bad) {
            clen = g.chr_len[chr]; gref = g.bases + g.chr_off[chr];
            if ((uint64_t)(pos - 1u) + aligned + nd > clen) bad = true;
        }
        if (bad) {
            if (lane == 0) dev_set_error(err, CBCG_ERR_CORRUPT, r0 + i);
            for (uint32_t j = lane; j < len; j += 32u) dst[j] = 'N';
            if (lane == 0) dst[len] = '\n';
            continue;
        }
        const bool in_win = (chr == chr0) && ref_bytes && (uint64_t)(pos - 1u) >= w0 &&
                            ((uint64_t)(pos - 1u) - w0 + aligned + nd <= ref_bytes);
        const uint8_t *src = in_win ? (s_ref + (uint32_t)((uint64_t)(pos - 1u) - w0)) : (gref + (pos - 1u));

        if ((nd | ns | ni) == 0u) {                         /* perfect match or no edits: straight copy */
            for (uint32_t j = lane; j < len; j += 32u) dst[j] = src[j];
            if (lane == 0) dst[len] = '\n';
            continue;
        }
        const uint16_t *e = edits + rec.edit_off;
        bool malformed = false;
        if ((nd | ni) == 0u) {                              /* substitutions only: copy, then patch the SNP sites (:442-458) */
            for (uint32_t j = lane; j < len; j += 32u) dst[j] = src[j];
            if (lane == 0) dst[len] = '\n';
            __syncwarp();
            uint32_t carry = 0;                             /* SNP k sits at sum_{i<k}(p_i + 1) + p_k */
            for (uint32_t k0 = 0; k0 < ns; k0 += 32u) {
                const uint32_t k = k0 + lane; const uint32_t ed = (k < ns) ? e[k] : 0u;
                const uint32_t d = (k < ns) ? CBCG_EDIT_DELTA(ed) + 1u : 0u;
                const uint32_t s = warp_incl_scan(d) + carry;
                if (k < ns) { if (s - 1u >= len) malformed = true; else dst[s - 1u] = (uint8_t)base_char(CBCG_EDIT_TARGET(ed)); }
                carry = __shfl_sync(FULL_MASK, s, 31);
            }
            if (__any_sync(FULL_MASK, malformed)) {
                if (lane == 0) dev_set_error(err, CBCG_ERR_CORRUPT, r0 + i);
                for (uint32_t j = lane; j < len; j += 32u) dst[j] = 'N';
            }
            __syncwarp();
            continue;
        }
        if (max(nd, max(ns, ni)) > cap) {                   /* waits for the CTA's full-size scratch */
            if (lane == 0) S.big[atomicAdd(&S.n_big, 1u)] = (uint8_t)i;
            continue;
        }
        /* edit positions: prefix sums of the deltas (lists are short; 32 entries per step) */
        {
            uint32_t carry = 0;
            for (uint32_t k0 = 0; k0 < nd; k0 += 32u) {
                uint32_t k = k0 + lane, d = (k < nd) ? CBCG_EDIT_DELTA(e[k]) : 0u;
                uint32_t s = warp_incl_scan(d) + carry;
                if (k < nd) W.cumdel[k] = (uint16_t)s;
                carry = __shfl_sync(FULL_MASK, s, 31);
            }
            carry = 0;                                      /* SNP k sits at sum_{i<k}(p_i + 1) + p_k */
            for (uint32_t k0 = 0; k0 < ns; k0 += 32u) {
                uint32_t k = k0 + lane; uint32_t ed = (k < ns) ? e[nd + k] : 0u;
                uint32_t d = (k < ns) ? CBCG_EDIT_DELTA(ed) + 1u : 0u;
                uint32_t s = warp_incl_scan(d) + carry;
                if (k < ns) {
                    W.snp_at[k] = (uint16_t)(s - 1u); W.snp_ch[k] = (uint8_t)base_char(CBCG_EDIT_TARGET(ed));
                    if (s - 1u >= aligned) malformed = true;
                }
                carry = __shfl_sync(FULL_MASK, s, 31);
            }
            carry = 0;                                      /* insertion k lands at output index cum_k + k */
            for (uint32_t k0 = 0; k0 < ni; k0 += 32u) {
                uint32_t k = k0 + lane; uint32_t ed = (k < ni) ? e[nd + ns + k] : 0u;
                uint32_t d = (k < ni) ? CBCG_EDIT_DELTA(ed) : 0u;
                uint32_t s = warp_incl_scan(d) + carry;
                if (k < ni) {
                    W.ins_at[k] = (uint16_t)(s + k); W.ins_ch[k] = (uint8_t)base_char(CBCG_EDIT_TARGET(ed));
                    if (s > aligned) malformed = true;
                }
                carry = __shfl_sync(FULL_MASK, s, 31);
            }
        }
        malformed = __any_sync(FULL_MASK, malformed);
        __syncwarp();
        if (malformed) {
            if (lane == 0) dev_set_error(err, CBCG_ERR_CORRUPT, r0 + i);
            for (uint32_t j = lane; j < len; j += 32u) dst[j] = 'N';
            if (lane == 0) dst[len] = '\n';
            __syncwarp();
            continue;
        }
        for (uint32_t j = lane; j < len; j += 32u) {
            uint32_t q = 0; int hit = -1;
            for (uint32_t k = 0; k < ni; k++) { uint32_t at = W.ins_at[k]; q += (at < j); if (at == j) hit = (int)k; }
            uint32_t c;
            if (hit >= 0) c = W.ins_ch[hit];
            else {
                const uint32_t a = j - q;
                uint32_t sh = 0;
                for (uint32_t k = 0; k < nd; k++) sh += (W.cumdel[k] <= a);
                c = src[a + sh];
                for (uint32_t k = 0; k < ns; k++) if (W.snp_at[k] == a) c = W.snp_ch[k];
            }
            dst[j] = (uint8_t)c;
        }
        if (lane == 0) dst[len] = '\n';
        __syncwarp();
    }
    }
    fence_proxy_async_smem();          /* generic-proxy smem writes -> visible to the bulk store */
    __syncthreads();

    /* ---- tile image -> HBM */
    if (tile_base + tile_total > out_cap) { if (tid == 0) dev_set_error(err, CBCG_ERR_CAPACITY, r0); return; }
    uint8_t *gdst = out + tile_base;
    const uint32_t head = min(tile_total, (16u - shift) & 15u);
    const uint32_t mid = (tile_total - head) & ~15u;
    const uint32_t tail = tile_total - head - mid;
    if (tid == 0 && mid) {
        tma_store_1d(gdst + head, img + head, mid);
        tma_store_commit_wait();
    }
    if (tid >= 32 && tid < 32 + head) gdst[tid - 32] = img[tid - 32];
    if (tid >= 64 && tid < 64 + tail) gdst[head + mid + (tid - 64)] = img[head + mid + (tid - 64)];
}

void reconstruct_set_carveout(int pct) { cudaFuncSetAttribute(k3_reconstruct_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct); }   /* see k2_coder.cu */

int launch_reconstruct(uint64_t n_reads, const cbcg_read_rec *recs, const uint32_t *chr, const uint16_t *edits,
                       const DevGenome &g, uint8_t *out, uint64_t out_cap, uint32_t max_len, uint32_t fixed_len,
                       uint64_t *tile_desc, uint32_t *ticket, uint64_t *total_bytes,
                       unsigned long long *err, cudaStream_t st, cudaEvent_t ev_start, cudaEvent_t ev_stop, uint32_t ref_cap_hint) {
    if (n_reads == 0) return 0;
    const uint64_t tiles = reconstruct_num_tiles(n_reads);
    /* reference window: what a tile spans at the input's coverage (the caller's estimate), twice over; 0: the largest */
    const uint32_t ref_cap = ref_cap_hint ? std::min<uint32_t>(std::max<uint32_t>((ref_cap_hint + 511u) & ~511u, 1024u), K3_REF_CAP) : K3_REF_CAP;
    const auto smem_for = [](uint32_t rc, uint32_t ml) { const size_t img = ((size_t)K3_TILE * (ml + 1u) + 48u + 15u) & ~(size_t)15u; return sizeof(K3Smem) + 32 + rc + 64 + img + (img >> 4) + 32; };
    const size_t smem = smem_for(ref_cap, max_len);
    /* per device, so no process-wide cache (see launch_extract) */
    const size_t smem_max = smem_for(K3_REF_CAP, CBCG_MAX_READ_LEN);
    if (smem > smem_max) return -1;
    if (cudaFuncSetAttribute(k3_reconstruct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) != cudaSuccess) return -1;
    if (!fixed_len) {                                       /* closed-form offsets use neither the ticket nor the descriptors */
        if (cudaMemsetAsync(tile_desc, 0, tiles * sizeof(uint64_t), st) != cudaSuccess) return -1;
        if (cudaMemsetAsync(ticket, 0, sizeof(uint32_t), st) != cudaSuccess) return -1;
    }
    if (ev_start) cudaEventRecord(ev_start, st);
    k3_reconstruct_kernel<<<(unsigned)tiles, K3_THREADS, smem, st>>>(n_reads, recs, chr, edits, g, out, out_cap,
                                                                    tile_desc, ticket, total_bytes, err, max_len, fixed_len, ref_cap);
    if (ev_stop) cudaEventRecord(ev_stop, st);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
