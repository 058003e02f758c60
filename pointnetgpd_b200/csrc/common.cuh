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
__device__ __forceinline__ float relu_nan(float v) { return v < 0.f ? 0.f : v; }

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

// eval mode: y = gamma*(u + b - rm)/sqrt(rv+eps) + beta
__global__ void k_bn_eval_affine(int C, const float* bias, pgpd_bn bn, BnState st) {
    int c = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (c >= C) return;
    float rstd = 1.0f / sqrtf(bn.running_var[c] + BN_EPS);
    float sc = bn.gamma[c] * rstd;
    float b = bias ? bias[c] : 0.f;
    st.mean[c] = bn.running_mean[c] - b;
    st.rstd[c] = rstd;
    st.scale[c] = sc;
    st.shift[c] = bn.beta[c] + sc * (b - bn.running_mean[c]);
}

// ---- deterministic two-stage column reduction ------------------------------------------------------
// stage 1: tmp[s][c] = sum over the rows of slice s of part[row][c]   (double accumulation, fixed order)
// grid (ceil(C/32), S), block 256 = 32 columns x 8 row lanes
constexpr int REDUCE_MAX_SLICES = 32;
template <class T>
__global__ void k_colreduce_stage1(const T* __restrict__ part, int nblk, int C, double* __restrict__ tmp) {
    __shared__ double sh[8][33];
    const int tid = (int)threadIdx.x, cx = tid & 31, ry = tid >> 5;
    const int c = (int)blockIdx.x * 32 + cx;
    const int S = (int)gridDim.y, sl = (int)blockIdx.y;
    const int per = (nblk + S - 1) / S;
    const int r0 = sl * per, r1 = (r0 + per < nblk) ? r0 + per : nblk;
    double acc = 0.0;
    if (c < C) {
#pragma unroll 4
        for (int i = r0 + ry; i < r1; i += 8) acc += (double)part[(size_t)i * C + c];
    }
    sh[ry][cx] = acc;
    __syncthreads();
    if (ry == 0 && c < C) {
        double t = 0.0;
        for (int q = 0; q < 8; ++q) t += sh[q][cx];
        tmp[(size_t)sl * C + c] = t;
    }
}

inline int reduce_slices(int nblk) {
    int s = (nblk + 127) / 128;
    return s < 1 ? 1 : (s > REDUCE_MAX_SLICES ? REDUCE_MAX_SLICES : s);
}

// runs stage 1 and returns the number of slices S; afterwards tmp holds [S][C] doubles
template <class T>
inline int colreduce(const T* part, int nblk, int C, double* tmp, cudaStream_t s) {
    const int S = reduce_slices(nblk);
    launch(k_colreduce_stage1<T>, dim3((unsigned)((C + 31) / 32), (unsigned)S), dim3(256), 0, s, part, nblk, C, tmp);
    return S;
}

// one-kernel variant for at most 1024 partial rows: out[c] = (float) sum_rows part[row][c]; block = 32 columns x 32 row lanes
// (each lane sums its rows in order, the 32 lane sums are added in lane order: deterministic)
template <class T>
__global__ void k_colreduce_direct(const T* __restrict__ part, int nblk, int C, float* __restrict__ out) {
    __shared__ double sh[32][33];
    const int tid = (int)threadIdx.x, cx = tid & 31, ry = tid >> 5;
    const int c = (int)blockIdx.x * 32 + cx;
    double acc = 0.0;
    if (c < C) {
#pragma unroll 4
        for (int i = ry; i < nblk; i += 32) acc += (double)part[(size_t)i * C + c];
    }
    sh[ry][cx] = acc;
    __syncthreads();
    if (ry == 0 && c < C) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < 32; ++q) t += sh[q][cx];
        out[c] = (float)t;
    }
}

// slices of centred-square sums -> variance -> BatchNorm finalisation (train)
// The squares were centred on `centre` (nullptr: on mean_u itself): sum (u-c)^2 = sum (u-mu)^2 + count (mu-c)^2.
__global__ void k_bn_finalize_from_css(const double* tmp, int S, int C, const float* mean_u, double count,
                                       const float* bias, pgpd_bn bn, BnState st, const float* centre) {
    int c = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (c >= C) return;
    double s = 0.0;
    for (int i = 0; i < S; ++i) s += tmp[(size_t)i * C + c];
    const double mu = (double)mean_u[c];
    double var = s / count;
    if (centre) { const double d = mu - (double)centre[c]; var -= d * d; }
    bn_finalize_train(c, mu, var, count, bias, bn, st);
}

// out[c] = (float) sum_i tmp[i][c]
__global__ void k_reduce_f(const double* tmp, int S, int C, float* out) {
    int c = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (c >= C) return;
    double s = 0.0;
    for (int i = 0; i < S; ++i) s += tmp[(size_t)i * C + c];
    out[c] = (float)s;
}

// vsum[k] = scale * sum_s tmp[s][k]  (the slices of a two-stage column reduction), then
// mean_out[r] = (sum_k W[r][k] * vsum[k]) * inv   in double.  One warp per row (coalesced row reads, fixed shuffle tree);
// block = 256 threads = 8 rows, grid = ceil(rows / 8); cols <= 128.  Block 0 also stores vsum (kept for the backward).
__global__ void k_matvec_mean(const float* __restrict__ W, int rows, int cols, const double* __restrict__ tmp, int S, double scale,
                              double inv, float* __restrict__ mean_out, double* __restrict__ vsum_out) {
    __shared__ double vs[128];
    const int tid = (int)threadIdx.x;
    if (tid < cols) {
        double t = 0.0;
        for (int i = 0; i < S; ++i) t += tmp[(size_t)i * cols + tid];
        t *= scale;
        vs[tid] = t;
        if (blockIdx.x == 0 && vsum_out) vsum_out[tid] = t;
    }
    __syncthreads();
    const int r = (int)blockIdx.x * 8 + (tid >> 5), lane = tid & 31;
    double s = 0.0;
    if (r < rows)
        for (int k = lane; k < cols; k += 32) s += (double)W[(size_t)r * cols + k] * vs[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (r < rows && lane == 0) mean_out[r] = (float)(s * inv);
}

// out[c] = (float) sum over the nblk partial rows; one kernel for up to 1024 rows, the two-stage reduction otherwise
template <class T>
inline void colreduce_to_float(const T* part, int nblk, int C, float* out, double* tmp, cudaStream_t s) {
    if (nblk <= 1024) {
        launch(k_colreduce_direct<T>, dim3((unsigned)((C + 31) / 32)), dim3(1024), 0, s, part, nblk, C, out);
    } else {
        const int S = colreduce<T>(part, nblk, C, tmp, s);
        launch(k_reduce_f, dim3((unsigned)((C + 255) / 256)), dim3(256), 0, s, (const double*)tmp, S, C, out);
    }
}

__global__ void k_fill(float* p, size_t n, float v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// BatchNorm backward bookkeeping from slices [S][2][C] of (sum dz, sum dz*yhat):
//   dgamma = sum dz*yhat ; dbeta = sum dz ; m1 = dbeta/count ; m2 = dgamma/count
__global__ void k_bn_bwd_finalize(const double* tmp, int S, int C, double count,
                                  float* dgamma, float* dbeta, float* m1, float* m2) {
    int c = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (c >= C) return;
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < S; ++i) {
        s1 += tmp[((size_t)i * 2 + 0) * C + c];
        s2 += tmp[((size_t)i * 2 + 1) * C + c];
    }
    dgamma[c] = (float)s2;
    dbeta[c] = (float)s1;
    m1[c] = (float)(s1 / count);
    m2[c] = (float)(s2 / count);
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
