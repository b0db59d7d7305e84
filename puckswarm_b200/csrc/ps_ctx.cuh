// puckswarm_b200 -- execution context of the per-arena programs.
//
// Every per-arena phase (ps_physics.cuh, ps_sense.cuh, ...) is written once against a small "CTA context":
// thread id / count, barrier, block-wide exclusive scan and shared-memory atomics. On the GPU the context
// is one thread block (CtaCtx: __syncthreads, warp ballot + shuffle scans). The CPU-only unit tests
// instantiate the same templates with HostCtx (one virtual thread, no barriers) to check the device
// logic against the oracle without a GPU; HostCtx is never compiled into the product library.
#pragma once
#include "ps_math.cuh"

namespace ps {

#define PS_FOR(i, n) for (int i = ctx.tid(); i < (n); i += ctx.nthreads())

#if defined(__CUDACC__)
struct CtaCtx {
    int* scanScratch;   // shared: [33] ints
    __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
    __device__ __forceinline__ int nthreads() const { return (int)blockDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    // Block-wide exclusive scan of a 0/1 flag; every thread of the block must call it. Returns this thread's
    // offset and the block total. Warp ballots give the intra-warp part, one warp scans the warp totals.
    __device__ __forceinline__ int exscanFlag(bool flag, int& total) const {
        unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
        unsigned ballot = __ballot_sync(0xFFFFFFFFu, flag);
        int inWarp = __popc(ballot & ((1u << lane) - 1u));
        int nWarps = (blockDim.x + 31) >> 5;
        __syncthreads();   // protects scanScratch against the previous call
        if (lane == 0) scanScratch[warp] = __popc(ballot);
        __syncthreads();
        if (warp == 0) {
            int v = (int)lane < nWarps ? scanScratch[lane] : 0;
            int incl = v;
            for (int o = 1; o < 32; o <<= 1) {
                int n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if ((int)lane >= o) incl += n;
            }
            if ((int)lane < nWarps) scanScratch[lane] = incl - v;
            if (lane == 31) scanScratch[32] = incl;
        }
        __syncthreads();
        total = scanScratch[32];
        return scanScratch[warp] + inWarp;
    }
    // Same for arbitrary non-negative ints.
    __device__ __forceinline__ int exscanInt(int value, int& total) const {
        unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
        int incl = value;
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if ((int)lane >= o) incl += n;
        }
        int nWarps = (blockDim.x + 31) >> 5;
        __syncthreads();
        if (lane == 31) scanScratch[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int v = (int)lane < nWarps ? scanScratch[lane] : 0;
            int wi = v;
            for (int o = 1; o < 32; o <<= 1) {
                int n = __shfl_up_sync(0xFFFFFFFFu, wi, o);
                if ((int)lane >= o) wi += n;
            }
            if ((int)lane < nWarps) scanScratch[lane] = wi - v;
            if (lane == 31) scanScratch[32] = wi;
        }
        __syncthreads();
        total = scanScratch[32];
        return scanScratch[warp] + incl - value;
    }
    __device__ __forceinline__ int atomicAddI(int* p, int v) const { return atomicAdd(p, v); }
    __device__ __forceinline__ void atomicOrU(unsigned* p, unsigned v) const { atomicOr(p, v); }
    __device__ __forceinline__ void atomicMaxU(unsigned* p, unsigned v) const { atomicMax(p, v); }
    __device__ __forceinline__ void atomicMinI(int* p, int v) const { atomicMin(p, v); }
    __device__ __forceinline__ int atomicMinRet(int* p, int v) const { return atomicMin(p, v); }
    __device__ __forceinline__ void atomicMinU(unsigned* p, unsigned v) const { atomicMin(p, v); }
    __device__ __forceinline__ void atomicMaxI(int* p, int v) const { atomicMax(p, v); }
};
#endif

#if defined(PS_HOST_EMU)
struct HostCtx {
    int tid() const { return 0; }
    int nthreads() const { return 1; }
    void sync() const {}
    int exscanFlag(bool flag, int& total) const {
        total = flag ? 1 : 0;
        return 0;
    }
    int exscanInt(int value, int& total) const {
        total = value;
        return 0;
    }
    int atomicAddI(int* p, int v) const {
        int o = *p;
        *p += v;
        return o;
    }
    void atomicOrU(unsigned* p, unsigned v) const { *p |= v; }
    void atomicMaxU(unsigned* p, unsigned v) const {
        if (v > *p) *p = v;
    }
    void atomicMinI(int* p, int v) const {
        if (v < *p) *p = v;
    }
    int atomicMinRet(int* p, int v) const {
        int o = *p;
        if (v < o) *p = v;
        return o;
    }
    void atomicMinU(unsigned* p, unsigned v) const {
        if (v < *p) *p = v;
    }
    void atomicMaxI(int* p, int v) const {
        if (v > *p) *p = v;
    }
};
#endif

// Round n up to the next multiple of the thread count: loops that contain a block-wide scan must have
// the same trip count in every thread.
template <class Ctx>
PS_HD int roundUpThreads(const Ctx& ctx, int n) {
    int t = ctx.nthreads();
    return (n + t - 1) / t * t;
}

}  // namespace ps
