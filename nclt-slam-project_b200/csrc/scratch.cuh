// Per-context device scratch: a stack carved inside one API call, backed by a list of chunks so
// that a nested call can always get more memory without moving what is already carved.
#pragma once
#include "common.cuh"

struct Carver {
    nclt_ctx* c;
    explicit Carver(nclt_ctx* ctx) : c(ctx) {}
    template <typename T>
    T* take(size_t n) {
        size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
        // nclt_scratch_reserve() guaranteed room in the current chunk for everything carved after it
        T* p = reinterpret_cast<T*>(static_cast<char*>(c->scratch) + c->scratch_off);
        c->scratch_off += bytes;
        return p;
    }
};
static inline size_t pad256(size_t b) { return (b + 255) & ~size_t(255); }

struct ScratchScope {   // restores the stack when the API call returns
    nclt_ctx* c;
    size_t saved_off;
    int saved_chunk;
    explicit ScratchScope(nclt_ctx* ctx) : c(ctx), saved_off(ctx->scratch_off), saved_chunk(ctx->scratch_chunk) {}
    ~ScratchScope() {
        // a scope opened before the first chunk existed (saved_chunk == -1) returns to chunk 0, offset 0:
        // restoring -1 would make every later top-level call look "idle and too small" and re-allocate
        int chunk = saved_chunk >= 0 ? saved_chunk : (c->scratch_chunks.empty() ? -1 : 0);
        c->scratch_off = saved_off;
        c->scratch_chunk = chunk;
        if (chunk >= 0 && chunk < (int)c->scratch_chunks.size()) {
            c->scratch = c->scratch_chunks[chunk].first;
            c->scratch_bytes = c->scratch_chunks[chunk].second;
        }
    }
};
