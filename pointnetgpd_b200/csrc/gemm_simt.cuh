// gemm_simt.cuh -- generic fp32 CUDA-core tile GEMM with functor loaders and epilogues.
//
// C[m][n] = sum_k A(m,k) * B(k,n) over one BM x BN tile per block; A and B elements come from
// problem functors (so BatchNorm/ReLU/transform prologues are applied while loading) and the
// accumulator tile goes to a problem epilogue (so max-pool / statistics / masks are fused).
// This is the always-available fp32 path: every tensor-core kernel in this library is checked
// against it, and it is what PGPD_F_SIMT selects.
#pragma once
#include "platform.h"

namespace pgpd {

template <int BM_, int BN_, int BK_, int TM_, int TN_>
struct TileCfg {
    static constexpr int BM = BM_, BN = BN_, BK = BK_, TM = TM_, TN = TN_;
    static constexpr int TX = BN / TN, TY = BM / TM, NT = TX * TY;
    static constexpr int RED_ELEMS = (TX * BM > TY * BN) ? TX * BM : TY * BN;
    static constexpr int RED_BYTES = 8 * RED_ELEMS;
    static_assert(TM % 4 == 0 && TN % 4 == 0, "thread tile must be a multiple of 4");
    // rows/cols owned by a thread are split in chunks of 4 that are BM/(TM/4) apart, so that
    // shared-memory reads are conflict-free 128-bit loads
    __host__ __device__ static constexpr int row_of(int ty, int i) { return (i >> 2) * (BM / (TM / 4)) + ty * 4 + (i & 3); }
    __host__ __device__ static constexpr int col_of(int tx, int j) { return (j >> 2) * (BN / (TN / 4)) + tx * 4 + (j & 3); }
};

using CfgBig = TileCfg<128, 128, 16, 8, 8>;    // 256 threads
using CfgTall = TileCfg<128, 64, 16, 8, 4>;    // 256 threads
using CfgSmall = TileCfg<64, 64, 16, 4, 4>;    // 256 threads
using CfgHead = TileCfg<32, 32, 32, 4, 4>;     // 64 threads: many small blocks for the skinny head GEMMs

// Deterministic block reductions through shared memory ---------------------------------------
// sum over the rows of the tile (per column): v[j] holds this thread's partial for col_of(tx,j)
template <class Cfg>
__device__ __forceinline__ float block_col_sum(const float (&v)[Cfg::TN], int ty, int tx, float* red) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < Cfg::TN; ++j) red[ty * Cfg::BN + Cfg::col_of(tx, j)] = v[j];
    __syncthreads();
    float s = 0.f;
    int t = (int)threadIdx.x;
    if (t < Cfg::BN) {
        for (int r = 0; r < Cfg::TY; ++r) s += red[r * Cfg::BN + t];
    }
    return s;  // valid for threads t < BN: the sum of column t
}

// sum over the columns of the tile (per row)
template <class Cfg>
__device__ __forceinline__ float block_row_sum(const float (&v)[Cfg::TM], int ty, int tx, float* red) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < Cfg::TM; ++i) red[tx * Cfg::BM + Cfg::row_of(ty, i)] = v[i];
    __syncthreads();
    float s = 0.f;
    int t = (int)threadIdx.x;
    if (t < Cfg::BM) {
        for (int r = 0; r < Cfg::TX; ++r) s += red[r * Cfg::BM + t];
    }
    return s;  // valid for threads t < BM: the sum of row t
}

// max over the columns of the tile (per row) of 64-bit keys
template <class Cfg>
__device__ __forceinline__ unsigned long long block_row_max_u64(const unsigned long long (&v)[Cfg::TM], int ty, int tx,
                                                                unsigned long long* red) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < Cfg::TM; ++i) red[tx * Cfg::BM + Cfg::row_of(ty, i)] = v[i];
    __syncthreads();
    unsigned long long s = 0ull;
    int t = (int)threadIdx.x;
    if (t < Cfg::BM) {
        for (int r = 0; r < Cfg::TX; ++r) {
            unsigned long long o = red[r * Cfg::BM + t];
            s = o > s ? o : s;
        }
    }
    return s;
}

// The kernel -------------------------------------------------------------------------------------
// Problem P provides:
//   static constexpr bool A_KFAST;   // A(m,k) contiguous along k in memory (else along m)
//   static constexpr bool B_NFAST;   // B(k,n) contiguous along n in memory (else along k)
//   static constexpr int  SCRATCH;   // floats of per-block shared scratch
//   struct Blk { int m0, n0, k0, k1; ... };
//   __device__ void  setup(Blk&) const;                       // from blockIdx
//   __device__ void  prologue(const Blk&, float* scratch) const;   // all threads; must end synchronised
//   __device__ float loadA(const Blk&, const float* scratch, int m, int k) const;  // bounds -> 0
//   __device__ float loadB(const Blk&, const float* scratch, int k, int n) const;
//   __device__ void  epilogue(const Blk&, const float* scratch, float (&acc)[TM][TN], int ty, int tx, void* red) const;
template <class Cfg, class P>
__global__ void __launch_bounds__(Cfg::NT) gemm_kernel(P p) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, TM = Cfg::TM, TN = Cfg::TN, NT = Cfg::NT;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    __shared__ __align__(16) unsigned char red[Cfg::RED_BYTES];
    __shared__ __align__(16) float scratch[P::SCRATCH > 0 ? P::SCRATCH : 4];

    typename P::Blk blk;
    p.setup(blk);
    p.prologue(blk, scratch);

    const int tid = (int)threadIdx.x;
    const int tx = tid % Cfg::TX, ty = tid / Cfg::TX;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // software pipeline: the global loads of tile k+1 are issued (into registers) before tile k is multiplied
    constexpr int EA = (BM * BK) / NT, EB = (BN * BK) / NT;
    static_assert((BM * BK) % NT == 0 && (BN * BK) % NT == 0, "tile elements must divide evenly over the threads");
    float ra[EA], rb[EB];
    auto load_regs = [&](int kk) {
#pragma unroll
        for (int q = 0; q < EA; ++q) {
            const int e = tid + q * NT;
            int m, k;
            if (P::A_KFAST) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
            ra[q] = (kk + k < blk.k1) ? p.loadA(blk, scratch, blk.m0 + m, kk + k) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < EB; ++q) {
            const int e = tid + q * NT;
            int n, k;
            if (P::B_NFAST) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
            rb[q] = (kk + k < blk.k1) ? p.loadB(blk, scratch, kk + k, blk.n0 + n) : 0.f;
        }
    };
    auto store_regs = [&]() {
#pragma unroll
        for (int q = 0; q < EA; ++q) {
            const int e = tid + q * NT;
            int m, k;
            if (P::A_KFAST) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
            As[k][m] = ra[q];
        }
#pragma unroll
        for (int q = 0; q < EB; ++q) {
            const int e = tid + q * NT;
            int n, k;
            if (P::B_NFAST) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
            Bs[k][n] = rb[q];
        }
    };
    if (blk.k0 < blk.k1) load_regs(blk.k0);
    for (int kk = blk.k0; kk < blk.k1; kk += BK) {
        store_regs();
        __syncthreads();
        if (kk + BK < blk.k1) load_regs(kk + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int c = 0; c < TM / 4; ++c) {
                float4 v = *reinterpret_cast<const float4*>(&As[k][Cfg::row_of(ty, c * 4)]);
                a[c * 4 + 0] = v.x; a[c * 4 + 1] = v.y; a[c * 4 + 2] = v.z; a[c * 4 + 3] = v.w;
            }
#pragma unroll
            for (int c = 0; c < TN / 4; ++c) {
                float4 v = *reinterpret_cast<const float4*>(&Bs[k][Cfg::col_of(tx, c * 4)]);
                b[c * 4 + 0] = v.x; b[c * 4 + 1] = v.y; b[c * 4 + 2] = v.z; b[c * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    p.epilogue(blk, scratch, acc, ty, tx, (void*)red);
}

template <class Cfg, class P>
inline void launch_gemm(const P& p, dim3 grid, cudaStream_t stream) {
    launch(gemm_kernel<Cfg, P>, grid, dim3(Cfg::NT), 0, stream, p);
}

}  // namespace pgpd
