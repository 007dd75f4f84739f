/* common.cuh -- warp primitives, TMA bulk-copy + mbarrier helpers, single-pass look-back scan. sm_100a only. */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "cbcg.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "cbc_b200 kernels are written for sm_100a (B200) only"
#endif

#define FULL_MASK 0xffffffffu

/* Device-side error word: first failure wins. (code << 40 | item index). */
__device__ __forceinline__ void dev_set_error(unsigned long long *err, int code, uint64_t item) {
    unsigned long long v = ((unsigned long long)(uint32_t)(-code) << 40) | (item & 0xffffffffffull);
    atomicCAS(err, 0ull, v);
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) { return __reduce_add_sync(FULL_MASK, v); }
__device__ __forceinline__ uint32_t warp_min(uint32_t v) { return __reduce_min_sync(FULL_MASK, v); }

__device__ __forceinline__ uint64_t warp_sum64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

/* Inclusive warp scan. */
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    const uint32_t lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FULL_MASK, v, o);
        if (lane >= (uint32_t)o) v += t;
    }
    return v;
}

/* ---------------------------------------------------------------- base codes
 * char2basepair / basepair2char: reference src/sam_models.c:11-45. */
__device__ __forceinline__ uint32_t base_code(uint32_t c) {
    return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
}
__device__ __forceinline__ uint32_t base_char(uint32_t b) {
    return b == 0 ? 'A' : b == 1 ? 'C' : b == 2 ? 'G' : b == 3 ? 'T' : 'N';
}

/* ---------------------------------------------------------------- mbarrier + TMA 1-D bulk copy */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
/* global -> shared, 16-byte aligned addresses and size; completes on the mbarrier (UBLKCP in SASS). */
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
/* shared -> global, 16-byte aligned addresses and size. */
__device__ __forceinline__ void tma_store_1d(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

/* ---------------------------------------------------------------- unaligned 32-bit reads from shared */
__device__ __forceinline__ uint32_t lds_u32_unaligned(const uint8_t *base, uint32_t off) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(base + (off & ~3u));
    return __funnelshift_r(w[0], w[1], (off & 3u) * 8u);
}

/* ---------------------------------------------------------------- decoupled look-back (single-pass scan)
 * One 64-bit descriptor per tile: status in bits 62..63 (0 invalid, 1 aggregate, 2 inclusive prefix),
 * value in the low 62 bits. Tiles take their index from an atomic ticket, so every lower-numbered tile
 * is already resident: the wait cannot deadlock. Called by warp 0 of the CTA; returns the exclusive
 * prefix of this tile (sum of all lower tiles' totals). */
#define LB_AGG   (1ull << 62)
#define LB_PFX   (2ull << 62)
#define LB_VAL   ((1ull << 62) - 1)

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint64_t lookback_exclusive(uint64_t *desc, uint32_t tile, uint64_t total,
                                                       unsigned long long *err) {
    const uint32_t lane = lane_id();
    if (tile == 0) {
        if (lane == 0) st_volatile_u64(&desc[0], LB_PFX | total);
        return 0;
    }
    if (lane == 0) st_volatile_u64(&desc[tile], LB_AGG | total);
    uint64_t excl = 0;
    int64_t j = (int64_t)tile - 1;
    for (;;) {
        int64_t idx = j - (int64_t)lane;
        uint64_t d = (idx >= 0) ? ld_volatile_u64(&desc[idx]) : LB_PFX;
        uint32_t spins = 0;
        /* only the tiles between this one and the nearest published prefix matter: wait for those, not for all 32 */
        for (;;) {
            const uint32_t inv = __ballot_sync(FULL_MASK, (d >> 62) == 0);
            const uint32_t pfx = __ballot_sync(FULL_MASK, (d >> 62) == 2);
            const uint32_t upto = pfx ? ((2u << ((uint32_t)__ffs(pfx) - 1u)) - 1u) : 0xffffffffu;   /* lanes 0..first prefix */
            if (!(inv & upto)) break;
            if ((d >> 62) == 0) d = ld_volatile_u64(&desc[idx]);
            if (++spins > (1u << 24)) {            /* never expected: bail out loudly instead of hanging */
                if (lane == 0) dev_set_error(err, CBCG_ERR_INTERNAL, tile);
                d = LB_PFX;
            }
        }
        uint32_t pmask = __ballot_sync(FULL_MASK, (d >> 62) == 2);
        uint32_t first = pmask ? (uint32_t)(__ffs(pmask) - 1) : 32u;
        excl += warp_sum64(lane <= first ? (d & LB_VAL) : 0ull);
        if (pmask) break;
        j -= 32;
    }
    if (lane == 0) st_volatile_u64(&desc[tile], LB_PFX | (excl + total));
    return excl;
}
