// tc_ka.cuh -- tcgen05 layer-2 backward pass 1 fused with the Gram matrix of a2 ("K_A").
//
//   d a2[P][c] = sparse[P][c] - u_c - sum_j Q[c][j] a2[P][j]        (layer-3 backward collapse, tower.cuh / DESIGN.md)
//   dz2 = d a2 . [a2 > 0]  (stored, fp32)      BatchNorm2-backward sums  sum dz2, sum dz2 yhat2,  max |dz2|
//   Gram = sum_P a2 a2^T   as  hi.hi  and  hi.lo  accumulators (Gram = hh + hl + hl^T: the third pass is the transpose
//   of the second)
//
// One persistent CTA per SM, 64-point tiles that never straddle a cloud, a FOUR-deep ring of 32 KB operand buffers: the loader
// bulk-copies the raw u2 rows of a tile into a buffer, 8 converter warps turn them into the a2 = relu(bn2(u2)) hi/lo fp16
// operand tile IN PLACE, one thread issues the tile's MMAs, and the buffer is free again as soon as those complete.  The
// epilogue (16 warps, channel = TMEM lane) does NOT touch the operand buffers: it re-reads its u2 values (L2 hits: the bulk copy
// just fetched them) and the sparse rows of d a2 straight from global memory, issued BEFORE it waits for the accumulator, so
// the loads hide behind the MMAs.  (Round 1 / early round 2 staged the sparse rows in shared memory and read a2 back from the
// operand tile: the buffer then lived through copy + conversion + MMAs + epilogue, only two fitted, and the kernel ran at the
// sum of its stages -- 5.3 k cycles per tile against an HBM floor of 2.6 k.)
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace pgpd { namespace tc {

constexpr int KA_NT = 64;
constexpr int KA_NBUF = 4;
constexpr int KA_THREADS = 832;                             // 16 epilogue + 8 converter warps, loader warp, MMA issuer
constexpr int KA_Q_BYTES = 65536;                           // Q image [kb][part][128 rows][128 B]
constexpr int KA_OP_BYTES = 32768;                          // a2 tile [part][kb][64 rows][128 B]   (raw: [64][128] fp32)
constexpr int KA_OFF_BUF = KA_Q_BYTES;
constexpr int KA_OFF_MISC = KA_OFF_BUF + KA_NBUF * KA_OP_BYTES;
constexpr int KA_SMEM_BYTES = KA_OFF_MISC + 256 + 1024;
constexpr int KA_EPI_ROWS = 4;                              // partial rows per CTA (four 16-column groups)

struct KaParams {
    const __half* Qimg; const float* inv;                   // pre-packed Q (tails.cuh: q_uvec_block, extra shift ACT_SHIFT) and its row scales
    const float* uvec; const float* scale2; const float* shift2; const float* gamma2; const float* beta2;
    const float* Y2; const float* da2s; const int* slot;
    int B, N, tiles_per_cloud, ntiles;
    float* DZ2;
    float* part;      // [gridDim.x * 4][2][128]   sum dz2, sum dz2*yhat2
    float* pmax;      // [gridDim.x * 4][2][128]   max |dz2|, (unused, 0)
    float* gpart;     // [gridDim.x][2][128*128]   Gram partials: hi.hi, hi.lo   (accumulator units: x 256)
};

__global__ void __launch_bounds__(KA_THREADS, 1) k_ka_tc(KaParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    unsigned char* misc = smem + KA_OFF_MISC;
    const uint32_t bar0 = sbase + KA_OFF_MISC;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    // 0 q_full | 1..4 full (bulk copy landed) | 5..8 op_ready (converted) | 9..12 buf_empty (the tile's MMAs complete)
    // 13,14 acc_full (d a2 accumulator complete) | 15,16 acc_empty (epilogue has it in registers) | 17 final
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 160);

    const int tid = (int)threadIdx.x, lane = tid & 31;
    const int warp = (int)warp_uniform((uint32_t)tid >> 5);     // provably warp-uniform (tc_ptx.cuh: elect_one)
    if (tid == 0) {
        mbar_init(BAR(0), 1);
        for (int b = 0; b < KA_NBUF; ++b) { mbar_init(BAR(1 + b), 1); mbar_init(BAR(5 + b), 256); mbar_init(BAR(9 + b), 1); }
        mbar_init(BAR(13), 1); mbar_init(BAR(14), 1);
        mbar_init(BAR(15), 512); mbar_init(BAR(16), 512);
        mbar_init(BAR(17), 1);
        mbar_fence_init();
    }
    if (warp == 25) tmem_alloc<512>(smem_u32(tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = warp_uniform(*tmem_slot);

    const int G = (int)gridDim.x, cta = (int)blockIdx.x;
    const int t_begin = (int)(((long long)p.ntiles * cta) / G), t_end = (int)(((long long)p.ntiles * (cta + 1)) / G);
    // tuning aid (pgpd_debug_stream_counters): 0 loader wait buf_empty | 2 converter work | 3 mma wait op_ready | 4 mma issue
    // 5 epilogue wait acc_full | 6 epilogue work | 7 total
    long long* const dbg = g_stream_dbg;
    long long dacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long tk0 = dbg ? clock64() : 0;
#define KA_T(slot, call) do { const long long _t0 = dbg ? clock64() : 0; call; if (dbg) dacc[slot] += clock64() - _t0; } while (0)

    if (warp == 24) {
        // ===================== loader =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(BAR(0), KA_Q_BYTES);
            bulk_g2s(sbase, p.Qimg, KA_Q_BYTES, BAR(0));
            int i = 0;
            for (int t = t_begin; t < t_end; ++t, ++i) {
                const int b = i % KA_NBUF;
                const uint32_t ph = (uint32_t)(i / KA_NBUF) & 1u;
                const int cb = t / p.tiles_per_cloud, tt = t % p.tiles_per_cloud, n0 = tt * KA_NT;
                const int nv = (p.N - n0 < KA_NT) ? p.N - n0 : KA_NT;
                const size_t P0 = (size_t)cb * p.N + n0;
                KA_T(0, mbar_wait(BAR(9 + b), ph ^ 1));
                mbar_arrive_expect_tx(BAR(1 + b), (uint32_t)nv * C2 * 4u);
                bulk_g2s(sbase + KA_OFF_BUF + b * KA_OP_BYTES, p.Y2 + P0 * C2, (uint32_t)nv * C2 * 4u, BAR(1 + b));
            }
            if (dbg) dbg[cta * 8 + 0] = dacc[0];
        }
    } else if (warp == 25) {
        // ===================== MMA issuer: the whole warp runs the loop, one elected lane issues (tc_ptx.cuh: elect_one) =====================
        {
            constexpr uint32_t IDESC_K = idesc_f16(128, KA_NT);
            constexpr uint32_t IDESC_MN = idesc_f16_mn(128, 128);
            mbar_wait(BAR(0), 0);
            tc_fence_after_sync();
            uint32_t first = 1;
            const uint64_t dQ = desc_sw128_kmajor(sbase);
            int i = 0;
            for (int t = t_begin; t < t_end; ++t, ++i) {
                const int b = i % KA_NBUF, acc = i & 1;
                const uint32_t ph = (uint32_t)(i / KA_NBUF) & 1u, aph = (uint32_t)(i >> 1) & 1u;
                KA_T(3, mbar_wait(BAR(5 + b), ph));
                mbar_wait(BAR(15 + acc), aph ^ 1);
                tc_fence_after_sync();
                const long long ti0 = dbg ? clock64() : 0;
                const uint32_t op = sbase + KA_OFF_BUF + b * KA_OP_BYTES;
                const uint32_t d1 = tmem + (uint32_t)(acc * KA_NT);
                const uint64_t kop = desc_sw128_kmajor(op), mop = desc_sw128_mnmajor(op, 8192);
                if (elect_one()) {
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass) {
                            const uint32_t oa = (uint32_t)((kb * 2 + (pass == 1 ? 1 : 0)) * 16384);
                            const uint32_t ob = (uint32_t)(((pass == 2 ? 1 : 0) * 2 + kb) * 8192);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                mma_f16(d1, dQ + ((oa + k * 32) >> 4), kop + ((ob + k * 32) >> 4), IDESC_K, (kb | pass | k) ? 1u : 0u);
                        }
                    }
                    mma_commit(BAR(13 + acc));
                    // Gram: hh += a2_hi^T a2_hi,  hl += a2_hi^T a2_lo   (K = the tile's 64 points; both operands MN-major,
                    // the two 64-channel atoms of a part are 8 KB apart)
#pragma unroll
                    for (int pass = 0; pass < 2; ++pass) {
                        const uint32_t dg = tmem + 128u + (uint32_t)(pass * 128);
                        const uint32_t ob = pass ? 16384u : 0u;
#pragma unroll
                        for (int k = 0; k < KA_NT / 16; ++k)
                            mma_f16(dg, mop + ((k * 2048) >> 4), mop + ((ob + k * 2048) >> 4), IDESC_MN, (first && k == 0) ? 0u : 1u);
                    }
                    mma_commit(BAR(9 + b));
                }
                __syncwarp();
                first = 0;
                if (dbg) dacc[4] += clock64() - ti0;
            }
            if (elect_one()) mma_commit(BAR(17));
            __syncwarp();
            if (dbg && lane == 0) { dbg[cta * 8 + 3] = dacc[3]; dbg[cta * 8 + 4] = dacc[4]; }
        }
    } else if (warp < 16) {
        // ===================== epilogue: channel c = TMEM lane, 16 of the tile's 64 points per warp =====================
        // The kernel is bound by instruction issue (16 epilogue warps x 16 points per tile), so this loop is written for few
        // instructions per point: constant-offset loads / stores from one base pointer on full tiles, sparse rows loaded
        // unconditionally (row 0 for a point that owns none) and discarded after the wait -- a select on the loaded value before
        // the wait would make the thread wait for the load at once --, folded constants.
        const int q = warp & 3, cgp = warp >> 2;
        const int c = q * 32 + lane;
        const float ninv = -p.inv[c], u = p.uvec[c];
        const float gm = p.gamma2[c], g2inv = gm != 0.f ? 1.0f / gm : 0.f, nbg = -p.beta2[c] * g2inv;   // yhat2 = a2 g2inv + nbg
        const float sc2 = p.scale2[c], sh2 = p.shift2[c];
        const float* sp_c = p.da2s + c;
        float s1 = 0.f, s2 = 0.f, mxdz = 0.f;
        int cb = t_begin / p.tiles_per_cloud, tt = t_begin % p.tiles_per_cloud;      // advanced incrementally (no divisions per tile)
        // slot index (row of the sparse part of d a2, or -1) of this warp's 16 points: lane j < 16 holds point cgp*16 + j;
        // fetched one tile ahead
        auto tile_slot = [&](int cbx, int ttx) -> int {
            if (cbx >= p.B || lane >= 16) return -1;
            const int n = ttx * KA_NT + cgp * 16 + lane;
            return n < p.N ? __ldg(p.slot + (size_t)cbx * p.N + n) : -1;
        };
        int myslot = (t_begin < t_end) ? tile_slot(cb, tt) : -1;
        int i = 0;
        for (int t = t_begin; t < t_end; ++t, ++i) {
            const int acc = i & 1;
            const uint32_t aph = (uint32_t)(i >> 1) & 1u;
            const int n0 = tt * KA_NT;
            const int nv = (p.N - n0 < KA_NT) ? p.N - n0 : KA_NT;
            const size_t Pw = (size_t)cb * p.N + n0 + cgp * 16;          // this warp's first point
            // u2 and the sparse rows of this tile: loads issued before the wait for the accumulator
            const float* yin = p.Y2 + Pw * C2 + c;
            float yv[16], sv[16];
            if (nv == KA_NT) {
#pragma unroll
                for (int j = 0; j < 16; ++j) yv[j] = __ldg(yin + j * C2);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int pp = cgp * 16 + j;
                    yv[j] = __ldg(yin + (ptrdiff_t)((pp < nv ? pp : nv - 1) - cgp * 16) * C2);
                }
            }
            const unsigned own = __ballot_sync(0xffffffffu, myslot >= 0);      // bit j: point j of this warp owns a sparse row
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int sl = __shfl_sync(0xffffffffu, myslot, j);
                // `own` is the same in every lane: a uniform predicate on the LOAD (not a select on its result, which would make the
                // thread wait for it here); points without a sparse row -- two out of three -- cost no L2 traffic
                sv[j] = ((own >> j) & 1u) ? __ldg(sp_c + (size_t)(unsigned)sl * C2) : 0.f;
            }
            int cbn = cb, ttn = tt + 1;
            if (ttn == p.tiles_per_cloud) { ttn = 0; ++cbn; }
            myslot = (t + 1 < t_end) ? tile_slot(cbn, ttn) : -1;
            KA_T(5, mbar_wait(BAR(13 + acc), aph));
            tc_fence_after_sync();
            const long long te0 = dbg ? clock64() : 0;
            float v[16];
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * KA_NT + cgp * 16), v);
            tc_fence_before_sync();
            mbar_arrive(BAR(15 + acc));                     // the accumulator is in registers
            float* dzo = p.DZ2 + Pw * C2 + c;
            const int nj = nv - cgp * 16;                   // valid points of this warp's 16 (>= 16 on full tiles)
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (nv == KA_NT || j < nj) {
                    const float a = fmaxf(fmaf(sc2, yv[j], sh2), 0.f);      // a2 (same sign decision as the operand tile's)
                    const float da2 = fmaf(v[j], ninv, sv[j] - u);
                    const float dz = a > 0.f ? da2 : 0.f;
#ifndef PGPD_DIAG_NOSTORE
                    dzo[j * C2] = dz;
#endif
                    s1 += dz;
                    s2 = fmaf(dz, fmaf(a, g2inv, nbg), s2);                 // yhat2 only matters where dz != 0
                    mxdz = fmaxf(mxdz, fabsf(dz));
                }
            }
            cb = cbn; tt = ttn;
            if (dbg) dacc[6] += clock64() - te0;
        }
        if (dbg && warp == 0 && lane == 0) { dbg[cta * 8 + 5] = dacc[5]; dbg[cta * 8 + 6] = dacc[6]; dbg[cta * 8 + 7] = clock64() - tk0; }
        const size_t row = (size_t)cta * KA_EPI_ROWS + cgp;
        p.part[row * 2 * C2 + c] = s1; p.part[row * 2 * C2 + C2 + c] = s2;
        p.pmax[row * 2 * C2 + c] = mxdz; p.pmax[row * 2 * C2 + C2 + c] = 0.f;
    } else {
        // ===================== converters: u2 rows -> a2 = relu(bn2(u2)) hi/lo operand tile, in place =====================
        const int cw = warp - 16;                           // 0..7
        float4 sc = *reinterpret_cast<const float4*>(p.scale2 + 4 * lane);
        float4 sh = *reinterpret_cast<const float4*>(p.shift2 + 4 * lane);
        sc.x *= ACT_SCALE; sc.y *= ACT_SCALE; sc.z *= ACT_SCALE; sc.w *= ACT_SCALE;
        sh.x *= ACT_SCALE; sh.y *= ACT_SCALE; sh.z *= ACT_SCALE; sh.w *= ACT_SCALE;
        const int kb = lane >> 4, chunk = (lane & 15) >> 1, half8 = lane & 1;
        int i = 0;
        for (int t = t_begin; t < t_end; ++t, ++i) {
            const int b = i % KA_NBUF;
            const uint32_t ph = (uint32_t)(i / KA_NBUF) & 1u;
            const int tt = t % p.tiles_per_cloud, n0 = tt * KA_NT;
            const int nv = (p.N - n0 < KA_NT) ? p.N - n0 : KA_NT;
            mbar_wait(BAR(1 + b), ph);
            const long long tc0 = dbg ? clock64() : 0;
            unsigned char* opb = smem + KA_OFF_BUF + b * KA_OP_BYTES;
            float4 ry[8];
#pragma unroll
            for (int uu = 0; uu < 8; ++uu) ry[uu] = *reinterpret_cast<const float4*>(opb + (cw * 8 + uu) * 512 + lane * 16);
            named_bar_sync(1, 256);                         // every converter thread has read its raw rows
#pragma unroll
            for (int uu = 0; uu < 8; ++uu) {
                const int r = cw * 8 + uu;
                const bool ok = r < nv;
                const float a0 = ok ? fminf(fmaxf(fmaf(sc.x, ry[uu].x, sh.x), 0.f), 60000.f) : 0.f;
                const float a1 = ok ? fminf(fmaxf(fmaf(sc.y, ry[uu].y, sh.y), 0.f), 60000.f) : 0.f;
                const float a2 = ok ? fminf(fmaxf(fmaf(sc.z, ry[uu].z, sh.z), 0.f), 60000.f) : 0.f;
                const float a3 = ok ? fminf(fmaxf(fmaf(sc.w, ry[uu].w, sh.w), 0.f), 60000.f) : 0.f;
                __half2 h01, l01, h23, l23;
                split2(a0, a1, h01, l01);
                split2(a2, a3, h23, l23);
                const uint32_t off = (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4) + half8 * 8);
                uint2 hv, lv;
                hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
                lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
                *reinterpret_cast<uint2*>(opb + (0 * 2 + kb) * 8192 + off) = hv;
                *reinterpret_cast<uint2*>(opb + (1 * 2 + kb) * 8192 + off) = lv;
            }
            fence_proxy_async_smem();
            mbar_arrive(BAR(5 + b));
            if (dbg) dacc[2] += clock64() - tc0;
        }
        if (dbg && cw == 0 && lane == 0) dbg[cta * 8 + 2] = dacc[2];
    }
#undef KA_T

    // ===================== read-out of the Gram accumulators (warps 0..15) =====================
    if (warp < 16) {
        const int q = warp & 3, cg = warp >> 2, row = q * 32 + lane;
        mbar_wait(BAR(17), 0);
        tc_fence_after_sync();
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
            float* out = p.gpart + ((size_t)cta * 2 + part) * (C2 * C2) + (size_t)row * C2 + cg * 32;
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 128u + (uint32_t)(part * 128 + cg * 32), v);
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(out + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 25) tmem_dealloc<512>(tmem);
}

inline int launch_ka(const KaParams& p, int sms, cudaStream_t s) {
    static int done[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!done[dev & 63]) {
        cudaFuncSetAttribute(k_ka_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, KA_SMEM_BYTES);
        done[dev & 63] = 1;
    }
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    launch(k_ka_tc, dim3(grid), dim3(KA_THREADS), (size_t)KA_SMEM_BYTES, s, p);
    return grid;
}

}}  // namespace pgpd::tc
