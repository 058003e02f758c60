// tc_kf.cuh -- tcgen05 layer-2 forward GEMM ("K_F"):  u2[P][c] = sum_k W2[c][k] a1[P][k]   (64 -> 128 channels)
// stores u2 (fp32) and accumulates the centred squares for the train-mode BatchNorm2 statistics.
//
// Same structure as tc_kb.cuh / tc_ka.cuh: persistent CTA, 64-point tiles; the loader bulk-copies the raw fp32 a1
// rows (16 KB per tile) into a FOUR-deep ring of operand buffers, 4 converter warps turn them into hi/lo fp16
// operand tiles in place, one thread issues the 12 MMAs of a tile (M = 128 channels, N = 64 points, K = 64),
// 8 epilogue warps (channel = TMEM lane) store u2 and accumulate (u2 - mean)^2.  HBM-bound: 256 B read + 512 B
// written per point.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace pgpd { namespace tc {

constexpr int KF_NT = 64;
constexpr int KF_NBUF = 4;
constexpr int KF_THREADS = 448;                             // 8 epilogue + 4 converter warps, loader, MMA issuer
constexpr int KF_W_BYTES = 32768;                           // W2 image [part][128 rows][128 B]  (tails.cuh: k_tower_pre)
constexpr int KF_OP_BYTES = 16384;                          // a1 tile  [part][64 rows][128 B]    (raw: [64][64] fp32)
constexpr int KF_OFF_BUF = KF_W_BYTES;
constexpr int KF_OFF_MISC = KF_OFF_BUF + KF_NBUF * KF_OP_BYTES;
constexpr int KF_SMEM_BYTES = KF_OFF_MISC + 512 + 1024;
constexpr int KF_EPI_ROWS = 2;

struct KfParams {
    const __half* Wimg; const float* inv; const float* mean_u2;     // mean_u2 == nullptr: no statistics (eval mode)
    const float* A1; float* Y2; float* css_part;                    // css_part [gridDim.x * 2][128]
    int B, N, tiles_per_cloud, ntiles;
};

__global__ void __launch_bounds__(KF_THREADS, 1) k_kf_tc(KfParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    unsigned char* misc = smem + KF_OFF_MISC;
    const uint32_t bar0 = sbase + KF_OFF_MISC;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    // 0 w_full | 1..4 full | 5..8 op_ready | 9..12 buf_empty (MMAs done reading) | 13,14 acc_full | 15,16 acc_empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 256);

    const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(BAR(0), 1);
        for (int b = 0; b < KF_NBUF; ++b) { mbar_init(BAR(1 + b), 1); mbar_init(BAR(5 + b), 128); mbar_init(BAR(9 + b), 1); }
        mbar_init(BAR(13), 1); mbar_init(BAR(14), 1);
        mbar_init(BAR(15), 256); mbar_init(BAR(16), 256);
        mbar_fence_init();
    }
    if (warp == 13) tmem_alloc<128>(smem_u32(tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int G = (int)gridDim.x, cta = (int)blockIdx.x;
    const int t_begin = (int)(((long long)p.ntiles * cta) / G), t_end = (int)(((long long)p.ntiles * (cta + 1)) / G);

    if (warp == 12) {
        // ===================== loader =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(BAR(0), KF_W_BYTES);
            bulk_g2s(sbase, p.Wimg, KF_W_BYTES, BAR(0));
            int i = 0;
            for (int t = t_begin; t < t_end; ++t, ++i) {
                const int b = i % KF_NBUF;
                const uint32_t ph = (uint32_t)(i / KF_NBUF) & 1u;
                {
                    const int tp = t + 2 * KF_NBUF;         // two ring lengths ahead -> L2
                    if (tp < t_end) {
                        const int cb = tp / p.tiles_per_cloud, tt = tp % p.tiles_per_cloud, n0 = tt * KF_NT;
                        const int nv = (p.N - n0 < KF_NT) ? p.N - n0 : KF_NT;
                        l2_prefetch(p.A1 + ((size_t)cb * p.N + n0) * C1, (uint32_t)nv * C1 * 4u);
                    }
                }
                const int cb = t / p.tiles_per_cloud, tt = t % p.tiles_per_cloud, n0 = tt * KF_NT;
                const int nv = (p.N - n0 < KF_NT) ? p.N - n0 : KF_NT;
                const size_t P0 = (size_t)cb * p.N + n0;
                mbar_wait(BAR(9 + b), ph ^ 1);
                mbar_arrive_expect_tx(BAR(1 + b), (uint32_t)nv * C1 * 4u);
                bulk_g2s(sbase + KF_OFF_BUF + b * KF_OP_BYTES, p.A1 + P0 * C1, (uint32_t)nv * C1 * 4u, BAR(1 + b));
            }
        }
    } else if (warp == 13) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t IDESC = idesc_f16(128, KF_NT);
            mbar_wait(BAR(0), 0);
            tc_fence_after_sync();
            int i = 0;
            for (int t = t_begin; t < t_end; ++t, ++i) {
                const int b = i % KF_NBUF, acc = i & 1;
                const uint32_t ph = (uint32_t)(i / KF_NBUF) & 1u, aph = (uint32_t)(i >> 1) & 1u;
                mbar_wait(BAR(5 + b), ph);                  // operand tile converted
                mbar_wait(BAR(15 + acc), aph ^ 1);          // accumulator drained
                tc_fence_after_sync();
                const uint32_t op = sbase + KF_OFF_BUF + b * KF_OP_BYTES;
                const uint32_t d = tmem + (uint32_t)(acc * KF_NT);
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t wa = (pass == 1) ? sbase + 16384 : sbase;
                    const uint32_t wb = (pass == 2) ? op + 8192 : op;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma_f16(d, desc_sw128_kmajor(wa + k * 32), desc_sw128_kmajor(wb + k * 32), IDESC, (pass | k) ? 1u : 0u);
                }
                mma_commit(BAR(9 + b));
                mma_commit(BAR(13 + acc));
            }
        }
    } else if (warp < 8) {
        // ===================== epilogue: channel c = TMEM lane, 32 of the tile's 64 points per warp =====================
        const int q = warp & 3, half = warp >> 2;
        const int c = q * 32 + lane;
        const float inv = p.inv[c];
        const bool stats = p.mean_u2 != nullptr;
        const float mu = stats ? p.mean_u2[c] : 0.f;
        float css = 0.f;
        int i = 0;
        for (int t = t_begin; t < t_end; ++t, ++i) {
            const int acc = i & 1;
            const uint32_t aph = (uint32_t)(i >> 1) & 1u;
            const int cb = t / p.tiles_per_cloud, tt = t % p.tiles_per_cloud, n0 = tt * KF_NT;
            const int nv = (p.N - n0 < KF_NT) ? p.N - n0 : KF_NT;
            float* yo = p.Y2 + ((size_t)cb * p.N + n0) * C2 + c;
            mbar_wait(BAR(13 + acc), aph);
            tc_fence_after_sync();
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * KF_NT + half * 32), v);
            tc_fence_before_sync();
            mbar_arrive(BAR(15 + acc));                     // values are in registers: the accumulator may be overwritten
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int pp = half * 32 + j;
                if (pp < nv) {
                    const float u = v[j] * inv;
                    yo[(size_t)pp * C2] = u;
                    const float d = u - mu;
                    css = fmaf(d, d, css);
                }
            }
        }
        if (stats) p.css_part[((size_t)cta * KF_EPI_ROWS + half) * C2 + c] = css;
    } else if (warp < 12) {
        // ===================== converters: raw a1 rows -> hi/lo operand tile, in place =====================
        const int cw = warp - 8;                            // 0..3
        const int cg = lane & 15, chunk = cg >> 1, half8 = cg & 1;
        int i = 0;
        for (int t = t_begin; t < t_end; ++t, ++i) {
            const int b = i % KF_NBUF;
            const uint32_t ph = (uint32_t)(i / KF_NBUF) & 1u;
            const int tt = t % p.tiles_per_cloud, n0 = tt * KF_NT;
            const int nv = (p.N - n0 < KF_NT) ? p.N - n0 : KF_NT;
            mbar_wait(BAR(1 + b), ph);
            unsigned char* opb = smem + KF_OFF_BUF + b * KF_OP_BYTES;
            float4 ra[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) ra[u] = *reinterpret_cast<const float4*>(opb + (cw * 16 + u * 2 + (lane >> 4)) * 256 + cg * 16);
            named_bar_sync(1, 128);                         // every converter thread has read its raw rows
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = cw * 16 + u * 2 + (lane >> 4);
                const bool ok = r < nv;
                const float a0 = ok ? fminf(ra[u].x * ACT_SCALE, 60000.f) : 0.f;
                const float a1v = ok ? fminf(ra[u].y * ACT_SCALE, 60000.f) : 0.f;
                const float a2 = ok ? fminf(ra[u].z * ACT_SCALE, 60000.f) : 0.f;
                const float a3 = ok ? fminf(ra[u].w * ACT_SCALE, 60000.f) : 0.f;
                __half2 h01, l01, h23, l23;
                split2(a0, a1v, h01, l01);
                split2(a2, a3, h23, l23);
                const uint32_t off = (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4) + half8 * 8);
                uint2 hv, lv;
                hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
                lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
                *reinterpret_cast<uint2*>(opb + off) = hv;
                *reinterpret_cast<uint2*>(opb + 8192 + off) = lv;
            }
            fence_proxy_async_smem();
            mbar_arrive(BAR(5 + b));
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 13) tmem_dealloc<128>(tmem);
}

inline int launch_kf(const KfParams& p, int sms, cudaStream_t s) {
    static int done[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!done[dev & 63]) {
        cudaFuncSetAttribute(k_kf_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, KF_SMEM_BYTES);
        done[dev & 63] = 1;
    }
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    launch(k_kf_tc, dim3(grid), dim3(KF_THREADS), (size_t)KF_SMEM_BYTES, s, p);
    return grid;
}

}}  // namespace pgpd::tc
