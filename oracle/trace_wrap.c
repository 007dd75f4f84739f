/*
 * oracle/trace_wrap.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Link-time symbol tracer for the unmodified reference encoder. Linked with
 *   -Wl,--wrap=send_value_to_as -Wl,--wrap=alloc_read_models_t
 *   -Wl,--wrap=alloc_rname_models_t -Wl,--wrap=initialize_stream_model_codebook
 * it records, for every symbol the reference hands to its arithmetic coder
 * (send_value_to_as, src/stream_model.c:53), which model it went through and
 * the symbol value, without changing the bytes the reference writes.
 *
 * Record format (little endian, 8 bytes): u32 (stream << 24 | ctx), u32 symbol.
 * Stream ids are the ones in include/cbcg.h (CBCG_STREAM_*).
 * Output file: $CBC_TRACE_OUT (default: cbc_trace.bin).
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "sam_block.h"

enum { S_CODEBOOK = 0, S_SAME_REF = 1, S_RNAME = 2, S_RLENGTH = 3, S_POS = 4, S_POS_ALPHA = 5,
       S_FLAG = 6, S_MATCH = 7, S_SNPS = 8, S_INDELS = 9, S_VAR = 10, S_CHARS = 11, S_OTHER = 15 };

read_models __real_alloc_read_models_t(uint32_t read_length);
rname_models __real_alloc_rname_models_t(void);
stream_model *__real_initialize_stream_model_codebook(uint32_t rescale);
void __real_send_value_to_as(Arithmetic_stream as, stream_model model, int32_t x);

#define MAP_BITS 19
#define MAP_SIZE (1u << MAP_BITS)
static const void *map_key[MAP_SIZE];
static uint32_t map_val[MAP_SIZE];
static FILE *trace_fp;

static uint32_t slot_of(const void *p) {
    uint64_t h = (uint64_t)(uintptr_t)p;
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 29;
    return (uint32_t)h & (MAP_SIZE - 1);
}

static void map_put(const void *p, uint32_t stream, uint32_t ctx) {
    uint32_t s = slot_of(p);
    while (map_key[s] && map_key[s] != p) s = (s + 1) & (MAP_SIZE - 1);
    map_key[s] = p;
    map_val[s] = (stream << 24) | ctx;
}

static uint32_t map_get(const void *p) {
    uint32_t s = slot_of(p);
    while (map_key[s]) {
        if (map_key[s] == p) return map_val[s];
        s = (s + 1) & (MAP_SIZE - 1);
    }
    return (uint32_t)S_OTHER << 24;
}

static void put_array(stream_model *arr, uint32_t n, uint32_t stream) {
    for (uint32_t i = 0; i < n; i++) map_put(arr[i], stream, i);
}

read_models __wrap_alloc_read_models_t(uint32_t read_length) {
    read_models m = __real_alloc_read_models_t(read_length);
    put_array(m->flag, 1, S_FLAG);
    put_array(m->pos, 1, S_POS);
    put_array(m->pos_alpha, 4, S_POS_ALPHA);
    put_array(m->match, 256, S_MATCH);
    put_array(m->snps, 1, S_SNPS);
    put_array(m->indels, 1, S_INDELS);
    put_array(m->var, 0xffff, S_VAR);
    put_array(m->chars, 6, S_CHARS);
    put_array(m->rlength, 4, S_RLENGTH);
    return m;
}

rname_models __wrap_alloc_rname_models_t(void) {
    rname_models m = __real_alloc_rname_models_t();
    put_array(m->same_ref, 1, S_SAME_REF);
    put_array(m->rname, 256, S_RNAME);
    return m;
}

stream_model *__wrap_initialize_stream_model_codebook(uint32_t rescale) {
    stream_model *m = __real_initialize_stream_model_codebook(rescale);
    put_array(m, 4, S_CODEBOOK);
    return m;
}

void __wrap_send_value_to_as(Arithmetic_stream as, stream_model model, int32_t x) {
    if (!trace_fp) {
        const char *path = getenv("CBC_TRACE_OUT");
        trace_fp = fopen(path ? path : "cbc_trace.bin", "wb");
        if (!trace_fp) { perror("cbc_trace: open"); exit(2); }
        setvbuf(trace_fp, NULL, _IOFBF, 1 << 20);
    }
    uint32_t rec[2] = { map_get(model), (uint32_t)x };
    fwrite(rec, sizeof rec, 1, trace_fp);
    __real_send_value_to_as(as, model, x);
}

__attribute__((destructor)) static void trace_close(void) {
    if (trace_fp) fclose(trace_fp);
}
