/*
 * k0_unpack.cu -- compact batch (cbcg_batch_compact) -> the SoA batch K1 reads (DevBatch).
 *
 * The link between host and device is the narrowest pipe on the path (about 55 GB/s against 6.5 TB/s of HBM), so a
 * batch crosses it packed: 2 bits per base, text lengths instead of offsets, chromosome runs instead of a word per read.
 * One CTA per tile of 128 reads rebuilds, in HBM, what load_sam_line (src/sam_file_allocation.c:437-529) would have
 * produced: the byte-per-base SEQ pool and the three offset arrays (a CTA scan over the tile's lengths on top of one
 * host-computed offset per tile), and the chromosome ordinal of every read. Bases other than A, C, G, T are patched in
 * from the exception list by a second, tiny kernel.
 */
#include "common.cuh"
#include "internal.h"

#define K0_TILE 128u

struct UnpackParams {
    uint64_t r_begin, r_end, n_reads;
    const uint16_t *seq_len, *cigar_len, *md_len;
    const uint8_t *seq2;
    const uint64_t *tile_base;          /* per tile of the whole batch: seq bytes, seq2 bytes, cigar bytes, md bytes before it */
    const uint64_t *run_first; const uint32_t *run_chr; uint32_t n_runs;
    uint64_t *seq_off, *cigar_off, *md_off; uint8_t *seq; uint32_t *chr;
    unsigned long long *err;
};

__global__ void __launch_bounds__(K0_TILE) k0_unpack_kernel(UnpackParams P) {
    __shared__ uint32_t wsum[4][4];
    __shared__ uint64_t s_seq[K0_TILE], s_seq2[K0_TILE];
    __shared__ uint16_t s_len[K0_TILE];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t tile = P.r_begin / K0_TILE + blockIdx.x;
    const uint64_t r = tile * K0_TILE + tid;
    const bool live = r < P.r_end;
    uint32_t v[4] = { 0, 0, 0, 0 };
    if (live) {
        v[0] = P.seq_len[r]; v[1] = (v[0] + 3u) >> 2; v[2] = P.cigar_len[r]; v[3] = P.md_len[r];
        if (v[0] == 0u || v[0] > CBCG_MAX_READ_LEN) { dev_set_error(P.err, CBCG_ERR_INPUT, r); v[0] = v[1] = 0; }
    }
    uint32_t incl[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t x = v[q];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(FULL_MASK, x, o); if (lane >= (uint32_t)o) x += y; }
        incl[q] = x;
        if (lane == 31) wsum[q][warp] = x;
    }
    __syncthreads();
    if (tid < 4u) {                                          /* the lengths of the tile against the host's offset of the next one */
        uint32_t tot = 0;
        for (uint32_t k = 0; k < K0_TILE / 32u; k++) tot += wsum[tid][k];
        if (P.tile_base[(tile + 1u) * 4u + tid] - P.tile_base[tile * 4u + tid] != (uint64_t)tot) dev_set_error(P.err, CBCG_ERR_INPUT, tile * K0_TILE);
    }
    uint64_t off[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t wb = 0;
        for (uint32_t k = 0; k < warp; k++) wb += wsum[q][k];
        off[q] = P.tile_base[tile * 4u + q] + wb + incl[q] - v[q];
    }
    if (live) {
        P.seq_off[r] = off[0]; P.cigar_off[r] = off[2]; P.md_off[r] = off[3];
        /* the end offset of the range's last read: K1 runs on this range before the next one is unpacked */
        if (r + 1 == P.r_end) { P.seq_off[r + 1] = off[0] + v[0]; P.cigar_off[r + 1] = off[2] + v[2]; P.md_off[r + 1] = off[3] + v[3]; }
        uint32_t lo = 0, hi = P.n_runs;                       /* the chromosome run this read lies in */
        while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (P.run_first[mid] <= r) lo = mid; else hi = mid; }
        P.chr[r] = P.run_chr[lo];
    }
    s_seq[tid] = off[0]; s_seq2[tid] = off[1]; s_len[tid] = (uint16_t)v[0];
    __syncthreads();
    /* SEQ: a warp per read, a lane per packed byte (four bases): consecutive lanes write consecutive bytes */
    for (uint32_t k = warp; k < K0_TILE; k += K0_TILE / 32u) {
        const uint32_t len = s_len[k];
        if (!len) continue;
        const uint8_t *src = P.seq2 + s_seq2[k];
        uint8_t *dst = P.seq + s_seq[k];
        for (uint32_t j = lane; 4u * j < len; j += 32u) {
            const uint32_t byte = src[j];
            const uint32_t c = 0x54474341u;                   /* "ACGT", little endian */
#pragma unroll
            for (uint32_t q = 0; q < 4u; q++) if (4u * j + q < len) dst[4u * j + q] = (uint8_t)(c >> (8u * ((byte >> (2u * q)) & 3u)));
        }
    }
}

__global__ void k0_patch_kernel(const uint32_t *exc_read, const uint16_t *exc_base, const uint8_t *exc_char, uint64_t n_exc,
                                const uint64_t *seq_off, uint8_t *seq) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_exc) seq[seq_off[exc_read[i]] + exc_base[i]] = exc_char[i];
}

int launch_unpack(uint64_t r_begin, uint64_t r_end, uint64_t n_reads, const uint16_t *seq_len, const uint16_t *cigar_len, const uint16_t *md_len,
                  const uint8_t *seq2, const uint64_t *tile_base, const uint64_t *run_first, const uint32_t *run_chr, uint32_t n_runs,
                  uint64_t *seq_off, uint64_t *cigar_off, uint64_t *md_off, uint8_t *seq, uint32_t *chr, unsigned long long *err, cudaStream_t st) {
    if (r_end <= r_begin) return 0;
    UnpackParams P = { r_begin, r_end, n_reads, seq_len, cigar_len, md_len, seq2, tile_base, run_first, run_chr, n_runs, seq_off, cigar_off, md_off, seq, chr, err };
    const uint64_t tiles = (r_end - r_begin + K0_TILE - 1) / K0_TILE;      /* r_begin is a multiple of the tile */
    k0_unpack_kernel<<<(unsigned)tiles, K0_TILE, 0, st>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
int launch_patch(const uint32_t *exc_read, const uint16_t *exc_base, const uint8_t *exc_char, uint64_t n_exc, const uint64_t *seq_off, uint8_t *seq, cudaStream_t st) {
    if (!n_exc) return 0;
    k0_patch_kernel<<<(unsigned)((n_exc + 255u) / 256u), 256, 0, st>>>(exc_read, exc_base, exc_char, n_exc, seq_off, seq);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
void unpack_set_carveout(int pct) {
    cudaFuncSetAttribute(k0_unpack_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k0_patch_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}
