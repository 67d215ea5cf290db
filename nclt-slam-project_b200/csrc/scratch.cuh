// Per-context device scratch: one grow-only buffer carved as a stack inside one API call.
#pragma once
#include "common.cuh"

struct Carver {
    nclt_ctx* c;
    size_t off;
    explicit Carver(nclt_ctx* ctx) : c(ctx), off(ctx->scratch_off) {}
    template <typename T>
    T* take(size_t n) {
        size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
        T* p = reinterpret_cast<T*>(static_cast<char*>(c->scratch) + off);
        off += bytes;
        c->scratch_off = off;
        return p;
    }
};
static inline size_t pad256(size_t b) { return (b + 255) & ~size_t(255); }

struct ScratchScope {   // resets the stack when the outermost API call returns
    nclt_ctx* c;
    size_t saved;
    explicit ScratchScope(nclt_ctx* ctx) : c(ctx), saved(ctx->scratch_off) {}
    ~ScratchScope() { c->scratch_off = saved; }
};
