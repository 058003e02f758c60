// common.cuh -- constants, workspace carving, BatchNorm bookkeeping and small reduction kernels
// shared by the tower and head code.
#pragma once
#include "platform.h"
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include "../../include/pgpd.h"

namespace pgpd {

constexpr int C1 = 64, C2 = 128, C3 = 1024;   // tower widths (pointnet.py:12-14,127-129)
constexpr int H1 = 512, H2 = 256;             // head widths  (pointnet.py:16-18,182-184)
constexpr float BN_EPS = 1e-5f;               // nn.BatchNorm1d defaults (pointnet.py:21-25)
constexpr float BN_MOM = 0.1f;

// ---- workspace carving: the same code measures (base == nullptr) and assigns ------------------
struct Carver {
    char* base;
    size_t off;
    explicit Carver(void* b) : base(reinterpret_cast<char*>(b)), off(0) {}
    template <class T> T* take(size_t n) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

// per-layer BatchNorm state kept between forward and backward
struct BnState {
    float* mean;    // mean of u = W a (bias-free pre-activation); the BN mean is mean + bias
    float* rstd;    // 1/sqrt(var + eps)
    float* scale;   // gamma * rstd
    float* shift;   // train: beta - scale*mean ; eval: beta + scale*(bias - running_mean)
    void carve(Carver& c, int C) {
        mean = c.take<float>(C); rstd = c.take<float>(C); scale = c.take<float>(C); shift = c.take<float>(C);
    }
};

// ---- ordered 32-bit encoding of floats (unsigned compare == float compare) ---------------------
// NaN (either sign) maps to the largest key, so a NaN activation wins the max-pool like it does in torch's MaxPool1d.
__device__ __forceinline__ unsigned ord_encode(float f) {
    if (f != f) return 0xFFFFFFFFu;
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_decode(unsigned k) {
    unsigned u = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
    return __uint_as_float(u);
}

// relu that keeps NaN (fmaxf(NaN, 0) is 0; torch.relu(NaN) is NaN)
__device__ __forceinline__ float relu_nan(float v) {
#if defined(PGPD_EMU)
    return v < 0.f ? 0.f : v;
#else
    float r;
    asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(v));      // one instruction; NaN if v is NaN
    return r;
#endif
}

// fp16 operand range of the tensor-core path: activations are pre-scaled by 2^4 before the hi/lo split, so anything
// above 60000/16 would saturate.  Producers of activations flag such values (and NaN) per cloud instead of clamping
// silently; the pooled feature of a flagged cloud is written as NaN (see k_tail_l3).
constexpr float TC_ACT_LIMIT = 60000.0f / 16.0f;

// ---- last-block-done: the block that arrives last at `counter` (zero before the launch) returns true and resets it ----
// Used by the fused tail kernels: per-block partial results are written first, the last block reduces them in a fixed
// order (deterministic) and finalises.  All threads of the block must call it.
__device__ __forceinline__ bool last_block_done(unsigned* counter, unsigned nblocks) {
    __shared__ int s_is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(counter, 1u);
        s_is_last = (t == nblocks - 1u) ? 1 : 0;
        if (s_is_last) *counter = 0u;
    }
    __syncthreads();
    const bool last = s_is_last != 0;
    if (last) __threadfence();
    return last;
}

// ---- BatchNorm finalisation ---------------------------------------------------------------------
// Given the batch mean of the bias-free pre-activation and its biased variance, write the folded
// affine and update the running statistics exactly as nn.BatchNorm1d does in training mode
// (momentum 0.1, unbiased running variance).
__device__ __forceinline__ void bn_finalize_train(int c, double mean_u, double var, double count,
                                                  const float* bias, pgpd_bn bn, BnState st) {
    if (var < 0.0) var = 0.0;
    float rstd = (float)(1.0 / sqrt(var + (double)BN_EPS));
    float sc = bn.gamma[c] * rstd;
    st.mean[c] = (float)mean_u;
    st.rstd[c] = rstd;
    st.scale[c] = sc;
    st.shift[c] = bn.beta[c] - sc * (float)mean_u;
    float mu = (float)mean_u + (bias ? bias[c] : 0.f);
    float unbiased = (float)(var * (count / (count - 1.0)));
    bn.running_mean[c] = (1.f - BN_MOM) * bn.running_mean[c] + BN_MOM * mu;
    bn.running_var[c] = (1.f - BN_MOM) * bn.running_var[c] + BN_MOM * unbiased;
    if (c == 0 && bn.num_batches_tracked) *bn.num_batches_tracked += 1;
}

// ---- optional CUDA-event timing of the dominant kernel ---------------------------------------------
struct Profiler {
    bool on = false;
    bool sticky = false;     // keep the recorded pairs across reads (pairs captured into a CUDA graph are re-recorded by every replay)
#ifndef PGPD_EMU
    static constexpr int MAXP = 8192;
    cudaEvent_t* ev = nullptr;
    int n = 0;
    static void record(cudaEvent_t e, cudaStream_t s) {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(s, &st);
        if (st == cudaStreamCaptureStatusActive) cudaEventRecordWithFlags(e, s, cudaEventRecordExternal);   // becomes a graph node
        else cudaEventRecord(e, s);
    }
    void begin(cudaStream_t s) {
        if (!on || n >= MAXP) return;
        if (!ev) { ev = new cudaEvent_t[2 * MAXP]; for (int i = 0; i < 2 * MAXP; ++i) cudaEventCreate(&ev[i]); }
        record(ev[2 * n], s);
    }
    void end(cudaStream_t s) {
        if (!on || n >= MAXP || !ev) return;
        record(ev[2 * n + 1], s);
        ++n;
    }
    int read(int* launches, float* total_ms) {
        float tot = 0.f;
        for (int i = 0; i < n; ++i) {
            if (cudaEventSynchronize(ev[2 * i + 1]) != cudaSuccess) return -1;
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]) != cudaSuccess) return -1;
            tot += ms;
        }
        *launches = n; *total_ms = tot;
        if (!sticky) n = 0;
        return 0;
    }
    void reset() { n = 0; }
#else
    void begin(cudaStream_t) {}
    void end(cudaStream_t) {}
    int read(int* launches, float* total_ms) { *launches = 0; *total_ms = 0.f; return 0; }
    void reset() {}
#endif
};
inline Profiler& profiler() { static thread_local Profiler p; return p; }

inline dim3 grid1d(size_t n, int block) { return dim3((unsigned)((n + block - 1) / block)); }

}  // namespace pgpd
