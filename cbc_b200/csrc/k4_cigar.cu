/*
 * k4_cigar.cu -- CIGAR recovery.
 *
 * Upstream declares `reconstructCigar(Dels, Insers, numDels, numIns, totalReadLength, recCigar)` and a per-read
 * `cigarFlags` ("to check if the cigar can be recovered from indels", include/sam_block.h:179,443) and sketches the
 * decoder (decompress_cigar, src/read_decompression.c:91-113: a flag per read, the CIGAR text itself when the flag is
 * 0), but implements neither. This file builds it over the edit records the path already has on the device:
 *
 *   implied CIGAR: the deletions and insertions of a read are stored in M coordinates (bases of M operations
 *     consumed before the event: Dels[k] / Insers[k].pos are deltas against the previous event of the same kind,
 *     src/read_compression.c:321-352), so the operations follow from merging the two lists: (gap)M, then the
 *     insertions at that coordinate as one I, then the deletions as one D; whatever is left of len - n_ins is the
 *     last M. The reference codes soft clips as insertions (:358-479), so a clipped read implies I where it had S.
 *   class of a read (encoder): 0 the implied text is the CIGAR; 1 / 2 / 3 it is after turning the first / the last /
 *     both end operations from I into S; 4 anything else (=, X, H, P, I next to D the other way round, non-canonical
 *     numbers, more than CIG_MAX_OPS operations): the text is kept verbatim.
 *
 * k4_cigar_class_kernel: a thread per read, the class from the read's record, edits and CIGAR text.
 * k4_cigar_emit_kernel: a thread per read, 128 reads per CTA: the text of every read ('\n'-terminated) from record,
 *   edits, class and the verbatim texts, written at offsets from a CTA scan and the same decoupled look-back across
 *   tiles as K1 / K3 (tile order from an atomic ticket).
 * The host side (section layout, api.cu: cbcg_cigar_pack / cbcg_cigar_unpack) stores only the reads whose class is not 0.
 */
#include "common.cuh"
#include "internal.h"

#define CIG_MAX_OPS 48u
#define CIG_TILE    128u

/* operations of the implied CIGAR as (count << 8 | op); returns their number, or 0xffffffff when they do not fit */
__device__ static uint32_t cig_implied(const cbcg_read_rec &r, const uint16_t *__restrict__ e, uint32_t *ops) {
    const uint32_t nd = r.n_dels, ni = r.n_ins;
    const uint16_t *dels = e, *ins = e + nd + r.n_snps;
    if (ni > r.len) return 0xffffffffu;
    const uint32_t total_m = (uint32_t)r.len - ni;
    uint32_t cur = 0, kd = 0, ki = 0, dc = 0, ic = 0, n = 0;
    while (kd < nd || ki < ni) {
        const uint32_t dn = kd < nd ? dc + CBCG_EDIT_DELTA(dels[kd]) : 0xffffffffu;
        const uint32_t in = ki < ni ? ic + CBCG_EDIT_DELTA(ins[ki]) : 0xffffffffu;
        const uint32_t at = min(dn, in);
        if (at > cur) { if (n >= CIG_MAX_OPS) return 0xffffffffu; ops[n++] = ((at - cur) << 8) | 'M'; cur = at; }
        if (in == at) {
            uint32_t k = 0;
            while (ki < ni && ic + CBCG_EDIT_DELTA(ins[ki]) == at) { ic = at; ki++; k++; }
            if (n >= CIG_MAX_OPS) return 0xffffffffu;
            ops[n++] = (k << 8) | 'I';
        }
        if (dn == at) {
            uint32_t k = 0;
            while (kd < nd && dc + CBCG_EDIT_DELTA(dels[kd]) == at) { dc = at; kd++; k++; }
            if (n >= CIG_MAX_OPS) return 0xffffffffu;
            ops[n++] = (k << 8) | 'D';
        }
    }
    if (total_m > cur) { if (n >= CIG_MAX_OPS) return 0xffffffffu; ops[n++] = ((total_m - cur) << 8) | 'M'; }
    else if (total_m < cur) return 0xffffffffu;              /* events beyond the read: not a record K1 writes */
    return n;
}

__global__ void __launch_bounds__(128)
k4_cigar_class_kernel(uint64_t n_reads, const cbcg_read_rec *__restrict__ recs, const uint16_t *__restrict__ edits,
                      const uint64_t *__restrict__ cigar_off, const uint8_t *__restrict__ cigar, uint8_t *__restrict__ cls) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const cbcg_read_rec rec = recs[r];
    uint32_t ops[CIG_MAX_OPS];
    const uint32_t n = cig_implied(rec, edits + rec.edit_off, ops);
    const uint8_t *t = cigar + cigar_off[r];
    const uint32_t tl = (uint32_t)(cigar_off[r + 1] - cigar_off[r]);
    uint32_t c = 0, i = 0, k = 0;
    bool same = n != 0xffffffffu;
    while (same && i < tl) {
        /* canonical numbers only (no leading zero, 1 .. 65535): then equal operations mean equal text */
        uint32_t num = 0, digits = 0;
        const uint32_t first = t[i];
        while (i < tl && t[i] >= '0' && t[i] <= '9' && digits < 6u) { num = num * 10u + (uint32_t)(t[i] - '0'); i++; digits++; }
        if (digits == 0u || digits > 5u || first == '0' || num > 65535u || i >= tl || k >= n) { same = false; break; }
        const uint32_t op = t[i++], want = ops[k];
        if ((want >> 8) != num) { same = false; break; }
        if ((want & 0xffu) != op) {
            if (op == 'S' && (want & 0xffu) == 'I' && (k == 0u || k + 1u == n)) c |= (k == 0u) ? 1u : 2u;   /* a one-operation CIGAR cannot be all clip: k == 0 wins */
            else { same = false; break; }
        }
        k++;
    }
    if (same && k != n) same = false;
    cls[r] = same ? (uint8_t)c : (uint8_t)4u;
}

__device__ __forceinline__ uint32_t cig_put_num(uint8_t *o, uint32_t v) {
    uint32_t d = v >= 10000u ? 5u : v >= 1000u ? 4u : v >= 100u ? 3u : v >= 10u ? 2u : 1u;
    for (uint32_t j = d; j-- > 0u;) { o[j] = (uint8_t)('0' + v % 10u); v /= 10u; }
    return d;
}

__global__ void __launch_bounds__(CIG_TILE)
k4_cigar_emit_kernel(uint64_t n_reads, const cbcg_read_rec *__restrict__ recs, const uint16_t *__restrict__ edits,
                     const uint8_t *__restrict__ cls, const uint64_t *__restrict__ exc_read, const uint64_t *__restrict__ exc_off,
                     const uint8_t *__restrict__ exc_text, uint64_t n_exc, uint8_t *__restrict__ out, uint64_t out_cap,
                     uint64_t *tile_desc, uint32_t *ticket, uint64_t *total_bytes, unsigned long long *err) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[CIG_TILE / 32u];
    __shared__ uint64_t s_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t r = (uint64_t)tile * CIG_TILE + tid;
    uint8_t text[CIG_MAX_OPS * 6u + 2u];
    uint32_t len = 0; const uint8_t *verb = nullptr;
    if (r < n_reads) {
        const uint32_t c = cls[r];
        if (c >= 4u) {                                        /* verbatim: its entry by binary search over the listed reads */
            uint64_t lo = 0, hi = n_exc;
            while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (exc_read[mid] < r) lo = mid + 1; else hi = mid; }
            if (lo < n_exc && exc_read[lo] == r) { verb = exc_text + exc_off[lo]; len = (uint32_t)(exc_off[lo + 1] - exc_off[lo]); }
            else dev_set_error(err, CBCG_ERR_FORMAT, r);
        } else {
            const cbcg_read_rec rec = recs[r];
            uint32_t ops[CIG_MAX_OPS];
            const uint32_t n = cig_implied(rec, edits + rec.edit_off, ops);
            if (n == 0xffffffffu) dev_set_error(err, CBCG_ERR_FORMAT, r);
            else for (uint32_t k = 0; k < n; k++) {
                uint32_t op = ops[k] & 0xffu;
                if (op == 'I' && ((k == 0u && (c & 1u)) || (k + 1u == n && k != 0u && (c & 2u)))) op = 'S';
                len += cig_put_num(text + len, ops[k] >> 8);
                text[len++] = (uint8_t)op;
            }
        }
        len += 1u;                                            /* '\n' */
    }
    const uint32_t incl = warp_incl_scan(len);
    if (lane == 31u) s_warp[warp] = incl;
    __syncthreads();
    uint32_t warp_base = 0, tile_total = 0;
#pragma unroll
    for (uint32_t k = 0; k < CIG_TILE / 32u; k++) { const uint32_t t = s_warp[k]; if (k < warp) warp_base += t; tile_total += t; }
    if (warp == 0) {
        const uint64_t base = lookback_exclusive(tile_desc, tile, tile_total, err);
        if (lane == 0) { s_base = base; if ((uint64_t)(tile + 1u) * CIG_TILE >= n_reads) *total_bytes = base + tile_total; }
    }
    __syncthreads();
    if (r >= n_reads) return;
    const uint64_t at = s_base + warp_base + incl - len;
    if (at + len > out_cap) { dev_set_error(err, CBCG_ERR_CAPACITY, r); return; }
    const uint8_t *src = verb ? verb : text;
    for (uint32_t j = 0; j + 1u < len; j++) out[at + j] = src[j];
    out[at + len - 1u] = '\n';
}

int launch_cigar_class(uint64_t n_reads, const cbcg_read_rec *recs, const uint16_t *edits, const uint64_t *cigar_off,
                       const uint8_t *cigar, uint8_t *cls, cudaStream_t st) {
    if (!n_reads) return 0;
    k4_cigar_class_kernel<<<(unsigned)((n_reads + 127u) / 128u), 128, 0, st>>>(n_reads, recs, edits, cigar_off, cigar, cls);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

uint64_t cigar_num_tiles(uint64_t n_reads) { return (n_reads + CIG_TILE - 1u) / CIG_TILE; }

int launch_cigar_emit(uint64_t n_reads, const cbcg_read_rec *recs, const uint16_t *edits, const uint8_t *cls,
                      const uint64_t *exc_read, const uint64_t *exc_off, const uint8_t *exc_text, uint64_t n_exc,
                      uint8_t *out, uint64_t out_cap, uint64_t *tile_desc, uint32_t *ticket, uint64_t *total_bytes,
                      unsigned long long *err, cudaStream_t st) {
    if (!n_reads) return 0;
    const uint64_t tiles = cigar_num_tiles(n_reads);
    if (cudaMemsetAsync(tile_desc, 0, tiles * sizeof(uint64_t), st) != cudaSuccess) return -1;
    if (cudaMemsetAsync(ticket, 0, sizeof(uint32_t), st) != cudaSuccess) return -1;
    k4_cigar_emit_kernel<<<(unsigned)tiles, CIG_TILE, 0, st>>>(n_reads, recs, edits, cls, exc_read, exc_off, exc_text, n_exc,
                                                               out, out_cap, tile_desc, ticket, total_bytes, err);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
