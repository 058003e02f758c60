// tc_accum.cuh -- tcgen05 GEMMs whose contraction runs over the POINTS of the batch (points on the MMA K axis):
//
//     D[m][n] = sum_P  Aop(P)[m] * Bop(P)[n]          m < 128, n < NB (128 or 64)
//
// Gram matrix of a2 (layer-3 backward collapse) and dW2 = dy2^T a1.  Each CTA walks a contiguous range of
// 128-point tiles, accumulating into ONE TMEM accumulator, and writes a per-CTA partial that a deterministic
// two-stage reduction sums afterwards.
//
// Operand tiles are staged exactly like the streaming kernels' B operand -- rows = points, 128 bytes
// (64 channels) per row per 64-channel atom, 128-byte swizzle -- and are consumed through MN-major UMMA
// descriptors (leading byte offset = atom pitch, stride byte offset = 1024 B per 8 points; validated by
// tests/tc_probe variants 11/13).  hi/lo 3-pass scheme as everywhere else.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_stream.cuh"

namespace pgpd { namespace tc {

constexpr int AC_NT = 128;        // points per tile (MMA K extent per tile)
constexpr int AC_THREADS = 448;      // 14 warps: (idle), MMA issuer, 4 epilogue, 8 producers

// MN-major SWIZZLE_128B descriptor: atoms of 64 M/N-elements (128 B) x 8 K-rows; atom pitch `lbo` bytes
__device__ __forceinline__ uint64_t desc_sw128_mnmajor(uint32_t saddr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_f16_mn(int M, int N) { return idesc_f16(M, N) | (1u << 15) | (1u << 16); }

// Traits T provides:
//   static constexpr int NB;          // columns of D: 128 or 64
//   static constexpr bool SAME;       // B operand == A operand (Gram)
//   struct Params { size_t M; int ntiles; float* part; ... };     // part [G][128*NB]
//   producer for A rows (128 channels, 32 lanes per row):  ProdA / RawA / fetchA / transformA
//   producer for B rows (64 channels, 16 lanes per row) :  ProdB / RawB / fetchB / transformB   (if !SAME)
//   __device__ static float out_scale(const Params&, int m);      // factor applied to row m of D on output
template <class T>
struct AccumCfg {
    static constexpr int NB = T::NB;
    static constexpr int A_BYTES = 2 * 2 * 16384;                         // [part][atom 0..1][128 rows][128 B]
    static constexpr int B_BYTES = T::SAME ? 0 : 2 * (NB / 64) * 16384;   // [part][atoms][128 rows][128 B]
    static constexpr int BUF_BYTES = A_BYTES + B_BYTES;
    static constexpr int OFF_MISC = 2 * BUF_BYTES;
    static constexpr int SMEM_BYTES = OFF_MISC + 1024 + 1024;
};

template <class T>
__global__ void __launch_bounds__(AC_THREADS, 1) k_accum_tc(typename T::Params p) {
    using Cfg = AccumCfg<T>;
    constexpr int NB = Cfg::NB;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    unsigned char* misc = smem + Cfg::OFF_MISC;
    const uint32_t bar0 = sbase + Cfg::OFF_MISC;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    // 0..1 full, 2..3 empty, 4 done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 128);

    const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(BAR(0), 256); mbar_init(BAR(1), 256);
        mbar_init(BAR(2), 1); mbar_init(BAR(3), 1);
        mbar_init(BAR(4), 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<128>(smem_u32(tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int G = (int)gridDim.x, cta = (int)blockIdx.x;
    const int t_begin = (int)(((long long)p.ntiles * cta) / G), t_end = (int)(((long long)p.ntiles * (cta + 1)) / G);

    if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t IDESC = idesc_f16_mn(128, NB);
            uint32_t phase = 0;
            int buf = 0;
            uint32_t first = 1;
            for (int t = t_begin; t < t_end; ++t) {
                mbar_wait(BAR(buf), phase);
                tc_fence_after_sync();
                const uint32_t a_hi = sbase + buf * Cfg::BUF_BYTES, a_lo = a_hi + 2 * 16384;
                const uint32_t b_hi = T::SAME ? a_hi : a_hi + Cfg::A_BYTES;
                const uint32_t b_lo = T::SAME ? a_lo : b_hi + (NB / 64) * 16384;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t wa = (pass == 1) ? a_lo : a_hi;
                    const uint32_t wb = (pass == 2) ? b_lo : b_hi;
#pragma unroll
                    for (int k = 0; k < AC_NT / 16; ++k) {
                        mma_f16(tmem, desc_sw128_mnmajor(wa + k * 2048, 16384), desc_sw128_mnmajor(wb + k * 2048, 16384), IDESC,
                                (first && pass == 0 && k == 0) ? 0u : 1u);
                    }
                }
                first = 0;
                mma_commit(BAR(2 + buf));
                if (++buf == 2) { buf = 0; phase ^= 1; }
            }
            mma_commit(BAR(4));
        }
    } else if (warp >= 2 && warp < 6) {
        // ===================== epilogue (once, at the end) =====================
        const int q = warp & 3, m = q * 32 + lane;
        float* out = p.part + (size_t)cta * 128 * NB + (size_t)m * NB;
        if (t_end > t_begin) {
            mbar_wait(BAR(4), 0);
            tc_fence_after_sync();
            const float sc = T::out_scale(p, m);
            for (int c0 = 0; c0 < NB; c0 += 32) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(out + c0 + j) = make_float4(v[j] * sc, v[j + 1] * sc, v[j + 2] * sc, v[j + 3] * sc);
            }
        } else {
            for (int c0 = 0; c0 < NB; ++c0) out[c0] = 0.f;
        }
    } else if (warp >= 6) {
        // ===================== operand producer =====================
        const int wp = warp - 6;
        typename T::ProdA pa;
        T::prodA_begin(pa, p, lane);
        typename T::ProdB pb;
        T::prodB_begin(pb, p, lane & 15);
        uint32_t phase = 0;
        int buf = 0;
        for (int t = t_begin; t < t_end; ++t) {
            const size_t P0 = (size_t)t * AC_NT;
            const int nvalid = (p.M - P0 < (size_t)AC_NT) ? (int)(p.M - P0) : AC_NT;
            if (wp == 0 && lane == 0 && t + 2 < t_end) {
                const size_t Pn = (size_t)(t + 2) * AC_NT;
                const int nvn = (p.M - Pn < (size_t)AC_NT) ? (int)(p.M - Pn) : AC_NT;
                T::prefetch(p, Pn, nvn);
            }
            mbar_wait(BAR(2 + buf), phase ^ 1);
            unsigned char* ab = smem + buf * Cfg::BUF_BYTES;
            // ---- A rows: 128 channels, one row per warp iteration
            {
                const int cg = lane, atom = cg >> 4, chunk = (cg & 15) >> 1, half8 = cg & 1;
                constexpr int U = 8;
                for (int i0 = 0; i0 < AC_NT / 8; i0 += U) {
                    typename T::RawA raw[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) { const int r = wp + 8 * (i0 + u); T::fetchA(pa, p, P0 + r, r < nvalid, cg, raw[u]); }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int r = wp + 8 * (i0 + u);
                        float v[4];
                        T::transformA(pa, p, r < nvalid, raw[u], v);
                        __half2 h01, l01, h23, l23;
                        split2(v[0], v[1], h01, l01);
                        split2(v[2], v[3], h23, l23);
                        const uint32_t off = (uint32_t)(atom * 16384 + r * 128 + ((chunk ^ (r & 7)) << 4) + half8 * 8);
                        uint2 hv, lv;
                        hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
                        lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
                        *reinterpret_cast<uint2*>(ab + off) = hv;
                        *reinterpret_cast<uint2*>(ab + 2 * 16384 + off) = lv;
                    }
                }
            }
            // ---- B rows: 64 channels, two rows per warp iteration
            if (!T::SAME) {
                unsigned char* bb = ab + Cfg::A_BYTES;
                const int cg = lane & 15, rsub = lane >> 4, chunk = cg >> 1, half8 = cg & 1;
                constexpr int U = 8;
                for (int i0 = 0; i0 < AC_NT / 16; i0 += U) {
                    typename T::RawB raw[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) { const int r = (wp + 8 * (i0 + u)) * 2 + rsub; T::fetchB(pb, p, P0 + r, r < nvalid, cg, raw[u]); }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int r = (wp + 8 * (i0 + u)) * 2 + rsub;
                        float v[4];
                        T::transformB(pb, p, r < nvalid, raw[u], v);
                        __half2 h01, l01, h23, l23;
                        split2(v[0], v[1], h01, l01);
                        split2(v[2], v[3], h23, l23);
                        const uint32_t off = (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4) + half8 * 8);
                        uint2 hv, lv;
                        hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
                        lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
                        *reinterpret_cast<uint2*>(bb + off) = hv;
                        *reinterpret_cast<uint2*>(bb + (NB / 64) * 16384 + off) = lv;
                    }
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(BAR(buf));
            if (++buf == 2) { buf = 0; phase ^= 1; }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<128>(tmem);
}

template <class T>
inline int launch_accum(const typename T::Params& p, int sms, cudaStream_t s) {
    static int done[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!done[dev & 63]) {
        cudaFuncSetAttribute(k_accum_tc<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, AccumCfg<T>::SMEM_BYTES);
        done[dev & 63] = 1;
    }
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    launch(k_accum_tc<T>, dim3(grid), dim3(AC_THREADS), (size_t)AccumCfg<T>::SMEM_BYTES, s, p);
    return grid;
}

// ==================================================================================================================
// Gram = sum_P a2[P] a2[P]^T,  a2 = relu(scale2*u2 + shift2)
// ==================================================================================================================
struct GramTC {
    static constexpr int NB = 128;
    static constexpr bool SAME = true;
    struct Params { size_t M; int ntiles; float* part; const float* Y2; const float* scale2; const float* shift2; };
    struct ProdA { float4 sc, sh; };
    struct RawA { float4 y; };
    __device__ static void prodA_begin(ProdA& s, const Params& p, int cg) {
        s.sc = *reinterpret_cast<const float4*>(p.scale2 + 4 * cg);
        s.sh = *reinterpret_cast<const float4*>(p.shift2 + 4 * cg);
        s.sc.x *= ACT_SCALE; s.sc.y *= ACT_SCALE; s.sc.z *= ACT_SCALE; s.sc.w *= ACT_SCALE;
        s.sh.x *= ACT_SCALE; s.sh.y *= ACT_SCALE; s.sh.z *= ACT_SCALE; s.sh.w *= ACT_SCALE;
    }
    __device__ static void prefetch(const Params& p, size_t P0, int nrows) { l2_prefetch(p.Y2 + P0 * C2, (uint32_t)nrows * C2 * 4u); }
    __device__ static void fetchA(ProdA&, const Params& p, size_t P, bool valid, int cg, RawA& r) {
        r.y = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) r.y = *reinterpret_cast<const float4*>(p.Y2 + P * C2 + 4 * cg);
    }
    __device__ static void transformA(ProdA& s, const Params&, bool valid, const RawA& r, float (&v)[4]) {
        v[0] = valid ? fminf(fmaxf(fmaf(s.sc.x, r.y.x, s.sh.x), 0.f), 60000.f) : 0.f;
        v[1] = valid ? fminf(fmaxf(fmaf(s.sc.y, r.y.y, s.sh.y), 0.f), 60000.f) : 0.f;
        v[2] = valid ? fminf(fmaxf(fmaf(s.sc.z, r.y.z, s.sh.z), 0.f), 60000.f) : 0.f;
        v[3] = valid ? fminf(fmaxf(fmaf(s.sc.w, r.y.w, s.sh.w), 0.f), 60000.f) : 0.f;
    }
    struct ProdB { int d; };
    struct RawB { int d; };
    __device__ static void prodB_begin(ProdB&, const Params&, int) {}
    __device__ static void fetchB(ProdB&, const Params&, size_t, bool, int, RawB&) {}
    __device__ static void transformB(ProdB&, const Params&, bool, const RawB&, float (&)[4]) {}
    __device__ static float out_scale(const Params&, int) { return 1.0f / (ACT_SCALE * ACT_SCALE); }
};

}}  // namespace pgpd::tc
