// tc_kf.cuh -- tcgen05 layer-1 + layer-2 forward of a train-mode tower ("K_F"):
//     a1[P][k] = relu(scale1_k (W1 T^T x_P)_k + shift1_k)   (CUDA cores, 3 FMAs per output, never written to memory)
//     u2[P][c] = sum_k W2[c][k] a1[P][k]                      (64 -> 128 channels, tcgen05)
// stores u2 (fp32) and accumulates what the train-mode BatchNorm2 statistics need: sum (u2 - c)^2 around a pilot centre c and the
// exact sum of u2 (var = sum (u2-c)^2 / M - (mean - c)^2), plus the exact sum of a1 (needed by the backward).
//
// Persistent CTA, 64-point tiles, a FOUR-deep ring of operand buffers: 8 "converter" warps compute the a1 rows of a tile from
// the 12 bytes of coordinates per point and write them as the hi/lo fp16 operand tile (round 1 wrote a1 to HBM in a separate
// kernel and bulk-copied it back: 512 B per point of traffic and one launch that no longer exist; the backward recomputes a1
// the same way, tc_kb.cuh), one thread issues the 12 MMAs of a tile (M = 128 channels, N = 64 points, K = 64), 8 epilogue warps
// (channel = TMEM lane) store u2 and accumulate the sums.  HBM: 12 B read + 512 B written per point.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace pgpd { namespace tc {

constexpr int KF_NT = 64;
constexpr int KF_NBUF = 4;
#ifndef PGPD_KF_CONV_WARPS
#define PGPD_KF_CONV_WARPS 8
#endif
constexpr int KF_CONV_WARPS = PGPD_KF_CONV_WARPS;                            // converter warps (each handles 64 / KF_CONV_WARPS rows of a tile)
constexpr int KF_CONV_THREADS = 32 * KF_CONV_WARPS;
constexpr int KF_WARP_LOAD = 8 + KF_CONV_WARPS, KF_WARP_MMA = KF_WARP_LOAD + 1;
constexpr int KF_THREADS = 32 * (KF_WARP_MMA + 1);          // 8 epilogue + converter warps, loader, MMA issuer
constexpr int KF_W_BYTES = 32768;                           // W2 image [part][128 rows][128 B]  (tails.cuh: k_tower_pre)
constexpr int KF_OP_BYTES = 16384;                          // a1 tile  [part][64 rows][128 B]    (raw: [64][64] fp32)
constexpr int KF_OFF_BUF = KF_W_BYTES;
constexpr int KF_OFF_X = KF_OFF_BUF + KF_NBUF * KF_OP_BYTES;   // [KF_NBUF][3][64] transformed coordinates of the tile
constexpr int KF_X_FLOATS = 2 * KF_CONV_WARPS * 64 > KF_NBUF * 3 * KF_NT ? 2 * KF_CONV_WARPS * 64 : KF_NBUF * 3 * KF_NT;   // also the a1-sum staging rows
constexpr int KF_OFF_MISC = KF_OFF_X + KF_X_FLOATS * 4;
constexpr int KF_SMEM_BYTES = KF_OFF_MISC + 512 + 1024;
constexpr int KF_EPI_ROWS = 2;

struct KfParams {
    const __half* Wimg; const float* inv; const float* centre;      // centre == nullptr: no statistics
    const float* x; const float* trans;                             // [B][3][N]; [B][9] or null
    const float* W1; const float* sc1; const float* sh1;            // conv1.weight [64][3], folded BatchNorm1
    float* Y2; float* css_part;                                     // css_part [gridDim.x * 2][128]: sum (u2-c)^2
    float* s1a_part;                                                // [gridDim.x][64] partial sums of a1 (or null)
    unsigned* bad;                                                  // [B] per-cloud flags: NaN / out-of-fp16-range activation
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

    const int tid = (int)threadIdx.x, lane = tid & 31;
    const int warp = (int)warp_uniform((uint32_t)tid >> 5);     // provably warp-uniform (tc_ptx.cuh: elect_one)
    if (tid == 0) {
        mbar_init(BAR(0), 1);
        for (int b = 0; b < KF_NBUF; ++b) { mbar_init(BAR(1 + b), 1); mbar_init(BAR(5 + b), KF_CONV_THREADS); mbar_init(BAR(9 + b), 1); }
        mbar_init(BAR(13), 1); mbar_init(BAR(14), 1);
        mbar_init(BAR(15), 256); mbar_init(BAR(16), 256);
        mbar_fence_init();
    }
    if (warp == KF_WARP_MMA) tmem_alloc<128>(smem_u32(tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = warp_uniform(*tmem_slot);

    const int G = (int)gridDim.x, cta = (int)blockIdx.x;
    const int t_begin = (int)(((long long)p.ntiles * cta) / G), t_end = (int)(((long long)p.ntiles * (cta + 1)) / G);

    if (warp == KF_WARP_LOAD) {
        // ===================== loader: the weight image, once =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(BAR(0), KF_W_BYTES);
            bulk_g2s(sbase, p.Wimg, KF_W_BYTES, BAR(0));
        }
    } else if (warp == KF_WARP_MMA) {
        // ===================== MMA issuer =====================
        {   // the whole warp runs the loop (uniform operands), one elected lane issues: tc_ptx.cuh: elect_one
            const bool el = elect_one();
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
                        if (el) mma_f16(d, desc_sw128_kmajor(wa + k * 32), desc_sw128_kmajor(wb + k * 32), IDESC, (pass | k) ? 1u : 0u);
                }
                if (el) mma_commit(BAR(9 + b));
                if (el) mma_commit(BAR(13 + acc));
            }
        }
    } else if (warp < 8) {
        // ===================== epilogue: channel c = TMEM lane, 32 of the tile's 64 points per warp =====================
        const int q = warp & 3, half = warp >> 2;
        const int c = q * 32 + lane;
        const float inv = p.inv[c];
        const bool stats = p.centre != nullptr;
        const float mu = stats ? p.centre[c] : 0.f;
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
            if (nv == KF_NT) {                              // full tile (the common case): no per-point bound checks
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float u = v[j] * inv;
#ifndef PGPD_DIAG_NOSTORE
                    yo[(size_t)(half * 32 + j) * C2] = u;
#endif
                    const float d = u - mu;
                    css = fmaf(d, d, css);
                }
            } else {
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
        }
        if (stats) {
            p.css_part[((size_t)cta * KF_EPI_ROWS + half) * C2 + c] = css;
        }
    } else if (warp < KF_WARP_LOAD) {
        // ===================== converters: coordinates -> a1 -> hi/lo operand tile =====================
        const int cw = warp - 8;                            // 0 .. KF_CONV_WARPS-1
        const int ctid = cw * 32 + lane;
        const int cg = lane & 15, chunk = cg >> 1, half8 = cg & 1;
        float* sx = reinterpret_cast<float*>(smem + KF_OFF_X);
        // my four channels 4 cg .. 4 cg + 3
        float w[4][3], sf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = 4 * cg + j;
            // BatchNorm1 scale and the 2^4 operand scale folded into the weights: a1 * 2^4 = relu(w' . x' + shift')
            const float s16 = p.sc1[c] * ACT_SCALE;
            w[j][0] = s16 * p.W1[c * 3 + 0]; w[j][1] = s16 * p.W1[c * 3 + 1]; w[j][2] = s16 * p.W1[c * 3 + 2];
            sf[j] = p.sh1[c] * ACT_SCALE;
        }
        float sa[4] = {0.f, 0.f, 0.f, 0.f};
        // the tile's points (threads 0..63, one point each) and the cloud's transform: the RAW loads of tile t + 1 are issued
        // while tile t is converted and first touched one iteration later (any arithmetic on them here would wait for them here)
        float rp[3], rT[9];
        auto load_raw = [&](int t) {
            rp[0] = rp[1] = rp[2] = 0.f;
#pragma unroll
            for (int e = 0; e < 9; ++e) rT[e] = (e % 4 == 0) ? 1.f : 0.f;
            if (t >= t_end || ctid >= KF_NT) return;
            const int cb = t / p.tiles_per_cloud, n = (t % p.tiles_per_cloud) * KF_NT + ctid;
            const float* xb = p.x + (size_t)cb * 3 * p.N + (n < p.N ? n : p.N - 1);
            rp[0] = __ldg(xb); rp[1] = __ldg(xb + p.N); rp[2] = __ldg(xb + 2 * (size_t)p.N);
            if (p.trans) {
#pragma unroll
                for (int e = 0; e < 9; ++e) rT[e] = __ldg(p.trans + (size_t)cb * 9 + e);
            }
        };
        load_raw(t_begin);
        int i = 0;
        for (int t = t_begin; t < t_end; ++t, ++i) {
            const int b = i % KF_NBUF;
            const uint32_t ph = (uint32_t)(i / KF_NBUF) & 1u;
            const int cb = t / p.tiles_per_cloud, tt = t % p.tiles_per_cloud, n0 = tt * KF_NT;
            const int nv = (p.N - n0 < KF_NT) ? p.N - n0 : KF_NT;
            const float t0 = rT[0] * rp[0] + rT[3] * rp[1] + rT[6] * rp[2];
            const float t1 = rT[1] * rp[0] + rT[4] * rp[1] + rT[7] * rp[2];
            const float t2 = rT[2] * rp[0] + rT[5] * rp[1] + rT[8] * rp[2];
            load_raw(t + 1);
            mbar_wait(BAR(9 + b), ph ^ 1);                  // the MMAs that read this buffer four tiles ago are done
            float* sxb = sx + b * (3 * KF_NT);
            if (ctid < KF_NT) { sxb[ctid] = t0; sxb[KF_NT + ctid] = t1; sxb[2 * KF_NT + ctid] = t2; }
            named_bar_sync(1, KF_CONV_THREADS);
            unsigned char* opb = smem + KF_OFF_BUF + b * KF_OP_BYTES;
            bool oor = false;
#pragma unroll
            for (int u = 0; u < 32 / KF_CONV_WARPS; ++u) {
                const int r = cw * (64 / KF_CONV_WARPS) + u * 2 + (lane >> 4);
                const bool ok = r < nv;
                const float q0 = sxb[r], q1 = sxb[KF_NT + r], q2 = sxb[2 * KF_NT + r];
                float a[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float v = ok ? relu_nan(fmaf(w[j][0], q0, fmaf(w[j][1], q1, fmaf(w[j][2], q2, sf[j])))) : 0.f;
                    oor = oor || !(v <= 60000.f);           // NaN / beyond the fp16 operand range: flagged per cloud, the pooled feature is poisoned
                    sa[j] += v;
                    a[j] = v;
                }
                __half2 h01, l01, h23, l23;
                split2(a[0], a[1], h01, l01);
                split2(a[2], a[3], h23, l23);
                const uint32_t off = (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4) + half8 * 8);
                uint2 hv, lv;
                hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
                lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
                *reinterpret_cast<uint2*>(opb + off) = hv;
                *reinterpret_cast<uint2*>(opb + 8192 + off) = lv;
            }
            if (oor) p.bad[cb] = 1u;
            fence_proxy_async_smem();
            mbar_arrive(BAR(5 + b));
        }
        if (p.s1a_part) {
            // partial sums of a1 -> one row per CTA [gridDim.x][64]: the (converter warp, half warp) partials of a column are
            // added through shared memory (the coordinate staging area is free: every tile has been converted)
            named_bar_sync(1, KF_CONV_THREADS);
            sx[(cw * 2 + (lane >> 4)) * C1 + 4 * cg + 0] = sa[0];
            sx[(cw * 2 + (lane >> 4)) * C1 + 4 * cg + 1] = sa[1];
            sx[(cw * 2 + (lane >> 4)) * C1 + 4 * cg + 2] = sa[2];
            sx[(cw * 2 + (lane >> 4)) * C1 + 4 * cg + 3] = sa[3];
            named_bar_sync(1, KF_CONV_THREADS);
            if (ctid < C1) {
                float tsum = 0.f;
#pragma unroll
                for (int g = 0; g < 2 * KF_CONV_WARPS; ++g) tsum += sx[g * C1 + ctid];
                p.s1a_part[(size_t)cta * C1 + ctid] = tsum;
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == KF_WARP_MMA) tmem_dealloc<128>(tmem);
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
