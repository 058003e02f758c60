// tc_gemm.cuh -- tcgen05 GEMM for the dense FC heads (pointnet.py:35-37, 191-193 and their backward):
//
//     C[m][n] = sum_k A(m,k) B(n,k)            fp32 in / fp32 out, fp32-grade 3-pass hi/lo fp16 products, TMEM accumulator
//
// Each operand is a row-major fp32 matrix in global memory, used either along its rows (the contraction index is the
// contiguous one: "K-major", e.g. X[b][i] and W[j][i] in U = X W^T) or across them (the contraction index is the ROW
// index: "MN-major", e.g. dU[b][j] and X[b][i] in dW = dU^T X).  In both cases a shared-memory operand tile is a stack
// of 128-byte rows (64 fp16) with the 128-byte swizzle -- a row is a 64-element piece of a global row -- and only the
// UMMA descriptor differs (validated by tests/tc_probe and by the tower kernels that use both forms).
// One CTA = one 128 x 128 output tile over one K slice (split-K over blockIdx.z; the partials are summed in a fixed
// order by k_splitk_finish); 8 loader warps convert the operands on the way in (double-buffered 64-wide K chunks),
// one thread issues the MMAs, 4 warps drain the accumulator.
// Operand scaling (powers of two, exact): activations x 2^4; weights and gradients by a per-tensor factor derived
// from a device-side max |x| (k_absmax2 / the BatchNorm-backward kernel), so that hi/lo stay in fp16's normal range.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace pgpd { namespace tc {

constexpr int GM_T = 128;                  // output tile edge
constexpr int GM_KC = 64;                  // K chunk
constexpr int GM_THREADS = 416;            // warps 0-3 epilogue, 4-7 A loaders, 8-11 B loaders, 12 MMA issuer
constexpr int GM_OP_BYTES = 32768;         // one operand chunk, hi + lo
constexpr int GM_OFF_MISC = 4 * GM_OP_BYTES;
constexpr int GM_SMEM_BYTES = GM_OFF_MISC + 256 + 1024;

struct GemmOp {
    const float* p;        // row-major matrix
    int ld;                // its row length (floats)
    int mode;              // 0: rows indexed by m (or n), contraction along the row; 1: rows indexed by k
    const unsigned* absmax_bits;   // device array of n_absmax bit patterns of partial max |x| (non-negative floats), or nullptr
    int n_absmax;
    float fixed_scale;     // used when absmax_bits == nullptr
};

struct GemmParams {
    GemmOp A, B;
    int M, N, K, kslice;   // kslice: K range per blockIdx.z (multiple of GM_KC), or K when there is no split
    float* C;              // [M][N] (no split) or partials [gridDim.z][M][N]
    const float* mask;     // optional [M][N]: output zeroed where mask <= 0 (only without split-K)
};

// partial max |x| over up to two arrays: out[which * gridDim.x + block] (bit patterns; no atomics, nothing to zero first: the
// consumer takes the max over the gridDim.x values, see gemm_op_scale).
// grid = (blocks, 2), block = 256; array lengths must be multiples of 4 and the pointers 16-byte aligned.
__global__ void k_absmax2(const float* __restrict__ a, size_t na, const float* __restrict__ b, size_t nb, unsigned* __restrict__ out) {
    __shared__ float sh[8];
    const int tid = (int)threadIdx.x, which = (int)blockIdx.y;
    const float4* p = reinterpret_cast<const float4*>(which ? b : a);
    const size_t n4 = (which ? nb : na) >> 2;
    float m0 = 0.f, m1 = 0.f;
    const size_t stride = (size_t)gridDim.x * 256;
    size_t i = (size_t)blockIdx.x * 256 + tid;
    for (; p && i + stride < n4; i += 2 * stride) {
        const float4 u = p[i], v = p[i + stride];
        m0 = fmaxf(m0, fmaxf(fmaxf(fabsf(u.x), fabsf(u.y)), fmaxf(fabsf(u.z), fabsf(u.w))));
        m1 = fmaxf(m1, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    if (p && i < n4) { const float4 u = p[i]; m0 = fmaxf(m0, fmaxf(fmaxf(fabsf(u.x), fabsf(u.y)), fmaxf(fabsf(u.z), fabsf(u.w)))); }
    float mx = fmaxf(m0, m1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) sh[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) mx = fmaxf(mx, sh[w]);
        out[(size_t)which * gridDim.x + blockIdx.x] = __float_as_uint(mx);
    }
}

// saturate to the fp16 operand range but keep NaN (fminf / fmaxf would turn NaN into a finite bound): gradients are scaled
// from their measured maxima and never reach the bound; activations beyond it were replaced by NaN where they were produced
__device__ __forceinline__ float clamp_keepnan(float t) { return t > 60000.f ? 60000.f : (t < -60000.f ? -60000.f : t); }

__device__ __forceinline__ float gemm_op_scale(const GemmOp& o) {
    if (!o.absmax_bits) return o.fixed_scale;
    unsigned bits = 0u;
    for (int i = 0; i < o.n_absmax; ++i) { const unsigned b = o.absmax_bits[i]; bits = b > bits ? b : bits; }   // non-negative floats order like their bits
    if (bits == 0u) return 1.f;
    int e = 139 - (int)((bits >> 23) & 0xFFu);          // max |x| * 2^e in [2^12, 2^13)
    e = e > 100 ? 100 : (e < -100 ? -100 : e);
    return __uint_as_float((uint32_t)(127 + e) << 23);
}

// stage one 128 (m/n) x 64 (k) operand chunk as hi/lo fp16; 4 warps (tid4 = 0..127)
__device__ __forceinline__ void gemm_stage(const GemmOp& o, float scale, int mn0, int mn_total, int k0, int k_end,
                                           unsigned char* dst, int tid4) {
    const int warp = tid4 >> 5, lane = tid4 & 31;
    if (o.mode == 0) {
        // 16 lanes per row (64 floats), 2 rows per warp iteration
        const int cg = lane & 15, rsub = lane >> 4, chunk = cg >> 1, half8 = cg & 1;
        const int kk = k0 + 4 * cg;
        constexpr int U = 8;
        for (int i0 = 0; i0 < 16; i0 += U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = (warp + 4 * (i0 + u)) * 2 + rsub;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (mn0 + r < mn_total && kk < k_end) {
                    const float* src = o.p + (size_t)(mn0 + r) * o.ld + kk;
                    if (kk + 3 < k_end) v[u] = *reinterpret_cast<const float4*>(src);
                    else { v[u].x = src[0]; if (kk + 1 < k_end) v[u].y = src[1]; if (kk + 2 < k_end) v[u].z = src[2]; }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = (warp + 4 * (i0 + u)) * 2 + rsub;
                const float a0 = clamp_keepnan(v[u].x * scale), a1 = clamp_keepnan(v[u].y * scale);
                const float a2 = clamp_keepnan(v[u].z * scale), a3 = clamp_keepnan(v[u].w * scale);
                __half2 h01, l01, h23, l23;
                split2(a0, a1, h01, l01);
                split2(a2, a3, h23, l23);
                const uint32_t off = (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4) + half8 * 8);
                uint2 hv, lv;
                hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
                lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
                *reinterpret_cast<uint2*>(dst + off) = hv;
                *reinterpret_cast<uint2*>(dst + 16384 + off) = lv;
            }
        }
    } else {
        // one warp per k row (128 floats = two 64-element atoms)
        const int atom = lane >> 4, chunk = (lane & 15) >> 1, half8 = lane & 1;
        const int c = mn0 + 4 * lane;
        constexpr int U = 8;
        for (int i0 = 0; i0 < 16; i0 += U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = warp + 4 * (i0 + u);
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k0 + r < k_end && c < mn_total) {
                    const float* src = o.p + (size_t)(k0 + r) * o.ld + c;
                    if (c + 3 < mn_total) v[u] = *reinterpret_cast<const float4*>(src);
                    else { v[u].x = src[0]; if (c + 1 < mn_total) v[u].y = src[1]; if (c + 2 < mn_total) v[u].z = src[2]; }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = warp + 4 * (i0 + u);
                const float a0 = clamp_keepnan(v[u].x * scale), a1 = clamp_keepnan(v[u].y * scale);
                const float a2 = clamp_keepnan(v[u].z * scale), a3 = clamp_keepnan(v[u].w * scale);
                __half2 h01, l01, h23, l23;
                split2(a0, a1, h01, l01);
                split2(a2, a3, h23, l23);
                const uint32_t off = (uint32_t)(atom * 8192 + r * 128 + ((chunk ^ (r & 7)) << 4) + half8 * 8);
                uint2 hv, lv;
                hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
                lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
                *reinterpret_cast<uint2*>(dst + off) = hv;
                *reinterpret_cast<uint2*>(dst + 16384 + off) = lv;
            }
        }
    }
}

__global__ void __launch_bounds__(GM_THREADS, 1) k_gemm_tc(GemmParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    unsigned char* misc = smem + GM_OFF_MISC;
    const uint32_t bar0 = sbase + GM_OFF_MISC;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    // 0,1 full (256 loader threads) | 2,3 empty (MMA commit) | 4 done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 128);
    const int tid = (int)threadIdx.x, lane = tid & 31;
    const int warp = (int)warp_uniform((uint32_t)tid >> 5);     // provably warp-uniform (tc_ptx.cuh: elect_one)
    if (tid == 0) {
        mbar_init(BAR(0), 256); mbar_init(BAR(1), 256);
        mbar_init(BAR(2), 1); mbar_init(BAR(3), 1);
        mbar_init(BAR(4), 1);
        mbar_fence_init();
    }
    if (warp == 12) tmem_alloc<128>(smem_u32(tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = warp_uniform(*tmem_slot);

    const int n0 = (int)blockIdx.x * GM_T, m0 = (int)blockIdx.y * GM_T;
    const int k_begin = (int)blockIdx.z * p.kslice;
    const int k_end = (k_begin + p.kslice < p.K) ? k_begin + p.kslice : p.K;
    const int nchunks = (k_end - k_begin + GM_KC - 1) / GM_KC;
    const float sa = gemm_op_scale(p.A), sb = gemm_op_scale(p.B);

    if (warp >= 4 && warp < 12) {
        // ===================== loaders: warps 4-7 operand A, warps 8-11 operand B =====================
        const bool isB = warp >= 8;
        const int tid4 = tid - (isB ? 256 : 128);
        for (int i = 0; i < nchunks; ++i) {
            const int b = i & 1;
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            mbar_wait(BAR(2 + b), ph ^ 1);
            unsigned char* dst = smem + (b * 2 + (isB ? 1 : 0)) * GM_OP_BYTES;
            if (isB) gemm_stage(p.B, sb, n0, p.N, k_begin + i * GM_KC, k_end, dst, tid4);
            else gemm_stage(p.A, sa, m0, p.M, k_begin + i * GM_KC, k_end, dst, tid4);
            fence_proxy_async_smem();
            mbar_arrive(BAR(b));
        }
    } else if (warp == 12) {
        // ===================== MMA issuer =====================
        {   // the whole warp runs the loop (uniform operands), one elected lane issues: tc_ptx.cuh: elect_one
            const bool el = elect_one();
            const uint32_t idesc = idesc_f16(GM_T, GM_T) | ((uint32_t)(p.A.mode ? 1u : 0u) << 15) | ((uint32_t)(p.B.mode ? 1u : 0u) << 16);
            const uint32_t stepA = p.A.mode ? 2048u : 32u, stepB = p.B.mode ? 2048u : 32u;
            for (int i = 0; i < nchunks; ++i) {
                const int b = i & 1;
                const uint32_t ph = (uint32_t)(i >> 1) & 1u;
                mbar_wait(BAR(b), ph);
                tc_fence_after_sync();
                const uint32_t a = sbase + (b * 2 + 0) * GM_OP_BYTES, bb = sbase + (b * 2 + 1) * GM_OP_BYTES;
                const uint64_t da = p.A.mode ? desc_sw128_mnmajor(a, 8192) : desc_sw128_kmajor(a);
                const uint64_t db = p.B.mode ? desc_sw128_mnmajor(bb, 8192) : desc_sw128_kmajor(bb);
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t oa = (pass == 1) ? 16384u : 0u, ob = (pass == 2) ? 16384u : 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (el) mma_f16(tmem, da + ((oa + k * stepA) >> 4), db + ((ob + k * stepB) >> 4), idesc, (i | pass | k) ? 1u : 0u);
                }
                if (el) mma_commit(BAR(2 + b));
            }
            if (el) mma_commit(BAR(4));
        }
    } else if (warp < 4) {
        // ===================== epilogue: row m = TMEM lane =====================
        const int q = warp, m = m0 + q * 32 + lane;
        const float inv = 1.0f / (sa * sb);
        mbar_wait(BAR(4), 0);
        tc_fence_after_sync();
        float* out = p.C + (gridDim.z > 1 ? (size_t)blockIdx.z * p.M * p.N : 0) + (size_t)m * p.N + n0;
        const float* msk = (p.mask && gridDim.z == 1) ? p.mask + (size_t)m * p.N + n0 : nullptr;
        const bool vec = (p.N & 3) == 0;
#pragma unroll 1
        for (int c0 = 0; c0 < GM_T; c0 += 32) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (m < p.M) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int n = n0 + c0 + j;
                    if (n >= p.N) break;
                    float4 r = make_float4(v[j] * inv, v[j + 1] * inv, v[j + 2] * inv, v[j + 3] * inv);
                    if (vec && n + 3 < p.N) {
                        if (msk) {
                            const float4 mk = *reinterpret_cast<const float4*>(msk + c0 + j);
                            if (!(mk.x > 0.f)) r.x = 0.f;
                            if (!(mk.y > 0.f)) r.y = 0.f;
                            if (!(mk.z > 0.f)) r.z = 0.f;
                            if (!(mk.w > 0.f)) r.w = 0.f;
                        }
                        *reinterpret_cast<float4*>(out + c0 + j) = r;
                    } else {
                        const float rr[4] = {r.x, r.y, r.z, r.w};
                        for (int e = 0; e < 4 && n + e < p.N; ++e) {
                            float x = rr[e];
                            if (msk && !(msk[c0 + j + e] > 0.f)) x = 0.f;
                            out[c0 + j + e] = x;
                        }
                    }
                }
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 12) tmem_dealloc<128>(tmem);
}

inline void launch_gemm_tc(const GemmParams& p, int nsl, cudaStream_t s) {
    static int done[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!done[dev & 63]) {
        cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM_BYTES);
        done[dev & 63] = 1;
    }
    launch(k_gemm_tc, dim3((unsigned)idiv_up(p.N, GM_T), (unsigned)idiv_up(p.M, GM_T), (unsigned)nsl), dim3(GM_THREADS),
           (size_t)GM_SMEM_BYTES, s, p);
}

}}  // namespace pgpd::tc
