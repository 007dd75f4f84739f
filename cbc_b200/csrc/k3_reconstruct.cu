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
#include "common.cuh"
#include "internal.h"

#define K3_TILE     128u
#define K3_WARPS    8u
#define K3_THREADS  (K3_WARPS * 32u)
#define K3_REF_CAP  8192u

uint64_t reconstruct_num_tiles(uint64_t n_reads) { return (n_reads + K3_TILE - 1) / K3_TILE; }

struct K3Warp {                       /* per-warp scratch: edit positions of the read being built */
    uint16_t cumdel[256];             /* cumulative deletion offsets */
    uint16_t snp_at[256];             /* aligned index of each SNP */
    uint16_t ins_at[256];             /* output index of each inserted base */
    uint8_t  snp_ch[256];
    uint8_t  ins_ch[256];
};

struct K3Smem {
    uint64_t bar;
    uint64_t tile_base;
    uint32_t tile;
    uint32_t warp_tot[K3_WARPS];
    uint32_t out_off[K3_TILE];        /* byte offset of each read's line inside the tile */
    __align__(16) cbcg_read_rec rec[K3_TILE];
    uint32_t chr[K3_TILE];
    uint32_t fast[K3_TILE];           /* offset of the read's bases in ref[] when it takes the word-copy path, else ~0 */
    uint8_t  nsnp[K3_TILE];           /* SNPs to patch on that path */
    K3Warp w[K3_WARPS];
    __align__(16) uint8_t ref[K3_REF_CAP + 16];
    __align__(16) uint8_t out[16];    /* really K3_TILE * (max_len + 1) + 32 (dynamic) */
};

__global__ void __launch_bounds__(K3_THREADS)
k3_reconstruct_kernel(uint64_t n_reads, const cbcg_read_rec *__restrict__ recs, const uint32_t *__restrict__ chr_of,
                      const uint16_t *__restrict__ edits, DevGenome g, uint8_t *__restrict__ out, uint64_t out_cap,
                      uint64_t *tile_desc, uint32_t *ticket, uint64_t *total_bytes, unsigned long long *err,
                      uint32_t max_len) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    K3Smem &S = *reinterpret_cast<K3Smem *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    if (tid == 0) {
        S.tile = atomicAdd(ticket, 1u);
        mbar_init(&S.bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const uint32_t tile = S.tile;
    const uint64_t r0 = (uint64_t)tile * K3_TILE;
    const uint32_t nr = (uint32_t)min((uint64_t)K3_TILE, n_reads - r0);

    /* records (16 B each, coalesced) and chromosome ordinals */
    uint32_t my_bytes = 0;
    if (tid < nr) {
        uint4 v = reinterpret_cast<const uint4 *>(recs)[r0 + tid];
        *reinterpret_cast<uint4 *>(&S.rec[tid]) = v;
        S.chr[tid] = chr_of[r0 + tid];
        uint32_t len = S.rec[tid].len;
        if (len > max_len || len == 0) { dev_set_error(err, CBCG_ERR_CORRUPT, r0 + tid); len = 0; S.rec[tid].len = 0; }
        my_bytes = len + 1u;
    }
    __syncthreads();

    /* reference window under the tile's first read */
    const uint32_t chr0 = S.chr[0], pos0 = S.rec[0].pos;
    uint64_t w0 = 0; uint32_t ref_bytes = 0;
    if (chr0 < g.n_chr) {
        const uint64_t clen0 = g.chr_len[chr0];
        w0 = pos0 ? ((uint64_t)(pos0 - 1u) & ~15ull) : 0ull;
        uint64_t avail = (clen0 + REF_PAD > w0) ? ((clen0 + REF_PAD - w0) & ~15ull) : 0ull;
        ref_bytes = (uint32_t)min((uint64_t)K3_REF_CAP, avail);
    }
    if (tid == 0) {
        mbar_expect_tx(&S.bar, ref_bytes);
        if (ref_bytes) tma_load_1d(S.ref, g.bases + g.chr_off[chr0] + w0, ref_bytes, &S.bar);
    }

    /* CTA scan of line sizes (threads >= K3_TILE carry 0) */
    uint32_t incl = warp_incl_scan(my_bytes);
    if (lane == 31) S.warp_tot[warp] = incl;
    __syncthreads();
    uint32_t warp_base = 0, tile_total = 0;
#pragma unroll
    for (uint32_t k = 0; k < K3_WARPS; k++) { uint32_t t = S.warp_tot[k]; if (k < warp) warp_base += t; tile_total += t; }
    if (tid < nr) S.out_off[tid] = warp_base + incl - my_bytes;
    if (warp == 0) {
        uint64_t base = lookback_exclusive(tile_desc, tile, tile_total, err);
        if (lane == 0) {
            S.tile_base = base;
            if (r0 + nr == n_reads) *total_bytes = base + tile_total;
        }
    }
    __syncthreads();
    const uint64_t tile_base = S.tile_base;
    /* smem image is shifted so that smem offset == global offset (mod 16): the middle can leave by TMA */
    const uint32_t shift = (uint32_t)((uint64_t)(out + tile_base) & 15ull);
    uint8_t *img = S.out + shift;
    /* Which reads are plain copies of the window, possibly with a few substituted bases? (perfect matches :383-384
       and substitution-only reads :442-458: all of config 2.) They take the word-copy path below. */
    if (tid < nr) {
        const cbcg_read_rec &rec = S.rec[tid];
        const uint32_t len = rec.len, pos = rec.pos, chr = S.chr[tid];
        const uint32_t ns = rec.match ? 0u : rec.n_snps;
        bool ok = len > 0u && pos >= 1u && chr == chr0 && chr < g.n_chr && ref_bytes != 0u &&
                  (rec.match || (rec.n_dels == 0u && rec.n_ins == 0u && ns <= 16u));
        if (ok) ok = (uint64_t)(pos - 1u) + len <= g.chr_len[chr] && (uint64_t)(pos - 1u) >= w0 &&
                     ((uint64_t)(pos - 1u) - w0 + len + 8u <= ref_bytes);
        S.fast[tid] = ok ? (uint32_t)((uint64_t)(pos - 1u) - w0) : 0xffffffffu;
        S.nsnp[tid] = (uint8_t)ns;
    }
    __syncthreads();
    mbar_wait(&S.bar, 0);

    /* ---- word-copy path: half a warp per read, one aligned 4-byte word of the tile image per lane and step
       (unaligned source word = two shared loads and a funnel shift); the line's first and last words, shared with
       the neighbouring reads, go out byte by byte. */
    {
        const uint32_t half = lane >> 4, sub = lane & 15u;
        for (uint32_t i0 = warp * 2u; i0 < nr; i0 += K3_WARPS * 2u) {
            const uint32_t i = i0 + half;
            uint32_t so = 0xffffffffu, len = 0, doff = 0, ns = 0;
            if (i < nr) { so = S.fast[i]; len = S.rec[i].len; doff = shift + S.out_off[i]; ns = S.nsnp[i]; }
            if (so != 0xffffffffu) {
                const uint32_t w_last = (doff + len) >> 2;
                for (uint32_t w = (doff >> 2) + sub; w <= w_last; w += 16u) {
                    const int32_t j0 = (int32_t)(4u * w) - (int32_t)doff;          /* read-relative index of the word's first byte */
                    if (j0 >= 0 && (uint32_t)j0 + 3u <= len) {
                        uint32_t v = lds_u32_unaligned(S.ref, so + (uint32_t)j0);
                        if ((uint32_t)j0 + 3u == len) v = (v & 0x00ffffffu) | ((uint32_t)'\n' << 24);
                        reinterpret_cast<uint32_t *>(S.out)[w] = v;
                    } else {
#pragma unroll
                        for (uint32_t q = 0; q < 4u; q++) {
                            const int32_t j = j0 + (int32_t)q;
                            if (j >= 0 && (uint32_t)j <= len) S.out[4u * w + q] = ((uint32_t)j == len) ? (uint8_t)'\n' : S.ref[so + (uint32_t)j];
                        }
                    }
                }
            } else ns = 0;
            __syncwarp();
            if (__any_sync(FULL_MASK, ns != 0u)) {           /* SNP k sits at sum_{i<k}(p_i + 1) + p_k: 16-lane scan per read */
                uint32_t ed = 0, d = 0;
                if (sub < ns) { ed = edits[S.rec[i].edit_off + sub]; d = CBCG_EDIT_DELTA(ed) + 1u; }
                uint32_t sc = d;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL_MASK, sc, o, 16); if (sub >= (uint32_t)o) sc += t; }
                if (sub < ns) {
                    if (sc - 1u >= len) dev_set_error(err, CBCG_ERR_CORRUPT, r0 + i);
                    else S.out[doff + sc - 1u] = (uint8_t)base_char(CBCG_EDIT_TARGET(ed));
                }
            }
        }
    }

    K3Warp &W = S.w[warp];
    for (uint32_t i = warp; i < nr; i += K3_WARPS) {
        if (S.fast[i] != 0xffffffffu) continue;             /* done above */
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
        const uint8_t *src = in_win ? (S.ref + (uint32_t)((uint64_t)(pos - 1u) - w0)) : (gref + (pos - 1u));

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

int launch_reconstruct(uint64_t n_reads, const cbcg_read_rec *recs, const uint32_t *chr, const uint16_t *edits,
                       const DevGenome &g, uint8_t *out, uint64_t out_cap, uint32_t max_len,
                       uint64_t *tile_desc, uint32_t *ticket, uint64_t *total_bytes,
                       unsigned long long *err, cudaStream_t st, cudaEvent_t ev_start, cudaEvent_t ev_stop) {
    if (n_reads == 0) return 0;
    const uint64_t tiles = reconstruct_num_tiles(n_reads);
    const size_t smem = sizeof(K3Smem) + (size_t)K3_TILE * (max_len + 1u) + 64;
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(k3_reconstruct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        configured = smem;
    }
    if (cudaMemsetAsync(tile_desc, 0, tiles * sizeof(uint64_t), st) != cudaSuccess) return -1;
    if (cudaMemsetAsync(ticket, 0, sizeof(uint32_t), st) != cudaSuccess) return -1;
    if (ev_start) cudaEventRecord(ev_start, st);
    k3_reconstruct_kernel<<<(unsigned)tiles, K3_THREADS, smem, st>>>(n_reads, recs, chr, edits, g, out, out_cap,
                                                                    tile_desc, ticket, total_bytes, err, max_len);
    if (ev_stop) cudaEventRecord(ev_stop, st);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
