// tc_kb.cuh -- tcgen05 version of the fused layer-2 / layer-1 backward pass (math and outputs: l2bwd.cuh).
//
// One persistent CTA per SM walks 64-point tiles (never straddling a cloud).  Per tile:
//   loader     two bulk async copies (UBLKCP) bring the raw fp32 rows of dz2 [64][128] and a1 [64][64] straight
//              into the operand buffers (48 KB; nothing passes through registers on the way in, so the bytes in
//              flight per SM are bounded by shared memory, not by the register file);
//   converters 16 warps read the raw rows, synchronise among themselves, and overwrite them IN PLACE with the
//              hi/lo fp16 operand tiles (128-byte swizzle), dz2 scaled per channel by a power of two;
//   MMA        D1[k][P]  = sum_c A1op[k][c] dz[P][c] + sum_k' A2op[k][k'] a1[P][k']      (d a1 without the constant)
//              D2[c][k] += sum_P dz[P][c] a1[P][k]                                       (C, for dW2; MN-major operands)
//              D3[m][k] += sum_P [a1_hi ; a1_lo][P][m] a1[P][k]                           (Gram of a1 in two passes)
//   epilogue   8 warps: d a1 -> ReLU mask (a1 read back from the operand tile) -> BatchNorm1-backward sums and the
//              per-cloud sums H = sum dz1 x_j; nothing is written per point.
//              (8 warps: only TMEM lanes 0-63 of the zero-padded M = 128 accumulator hold features);
// Accumulators: D1 double-buffered (2 x 64 TMEM columns), D2 and D3 persistent (64 columns each).
// Measured (profiles/README.md): per 64-point tile the stages take ~1.8 k (bulk copy), ~2.7 k (conversion), ~3.4 k (56 MMAs
// of ~49 cycles: an N <= 96 MMA costs the same as N = 64) and ~2.5 k cycles (epilogue) with two buffers in flight.
// Same fp32-grade 3-pass hi/lo scheme as the other tensor-core kernels.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace pgpd { namespace tc {

constexpr int KB_NT = 64;                                   // points per tile
constexpr int KB_THREADS = 832;                             // 26 warps: 8 epilogue + 16 converter, loader, MMA issuer
constexpr int KB_CONV_THREADS = 512;
constexpr int KB_NBUF = 3;                                  // operand buffers in flight
// The two A-operand images hold only their 64 real rows (features): the M = 128 MMAs read rows 64..127 from whatever follows
// in shared memory (the next 8 KB block), which only fills TMEM lanes 64..127 of the d a1 accumulator -- lanes nobody reads.
// Zero-padding the images to 128 rows (round 1) cost 48 KB, i.e. the third operand buffer.
constexpr int KB_A1_BYTES = 32768;                          // W2^T image  [kb][part][64 rows][128 B]
constexpr int KB_A2_BYTES = 16384;                          // -K image        [part][64 rows][128 B]
constexpr int KB_IMG_BLOCK = 8192;                          // bytes per (kb, part) block of an image
constexpr int KB_DZ_BYTES = 32768;                          // dz2 tile    [part][kb][64 rows][128 B]  (raw: [64][128] fp32)
constexpr int KB_A1T_BYTES = 16384;                         // a1 tile         [part][64 rows][128 B]  (raw: [64][64] fp32)
constexpr int KB_BUF_BYTES = KB_DZ_BYTES + KB_A1T_BYTES;
constexpr int KB_OFF_A2 = KB_A1_BYTES;
constexpr int KB_OFF_BUF = KB_A1_BYTES + KB_A2_BYTES;
constexpr int KB_OFF_X = KB_OFF_BUF + KB_NBUF * KB_BUF_BYTES;   // [KB_NBUF][3][64] floats
constexpr int KB_OFF_MISC = KB_OFF_X + KB_NBUF * 3 * KB_NT * 4;
constexpr int KB_SMEM_BYTES = KB_OFF_MISC + 256 + 1024;
constexpr int KB_EPI_GROUPS = 4;                            // column groups of 16 points: partial rows per tile / per CTA

// The per-channel scale of dz2 (esc / einv) and the two A-operand images of this pass are produced by tails.cuh: kb_prep_row.
struct KbParams {
    const __half* A1img; const __half* A2img; const float* ginv;
    const float* cvec; const float* gamma1; const float* beta1;
    const float* esc; const float* einv;
    const float* DZ2; const float* x; const float* trans;           // a1 is recomputed from the coordinates (tc_kf.cuh)
    const float* W1; const float* sc1; const float* sh1;
    int B, N, tiles_per_cloud, ntiles;
    float* Cpart;     // [gridDim.x][128*64]
    float* G1part;    // [gridDim.x][64*64]
    float* bnpart;    // [gridDim.x * 4][2][64]
    float* Hpart;     // [ntiles * 4][64*3]
};

__global__ void __launch_bounds__(KB_THREADS, 1) k_kb_tc(KbParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    unsigned char* misc = smem + KB_OFF_MISC;
    const uint32_t bar0 = sbase + KB_OFF_MISC;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    // 0 a_full | 1..3 full (bulk copies landed) | 4..6 op_ready (converted) | 7..9 buf_empty (all MMAs reading the buffer complete +
    // epilogue done with it) | 10,11 acc_full (D1 complete) | 12,13 acc_empty (epilogue has D1 in registers) | 14 final
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 128);
    float* sx = reinterpret_cast<float*>(smem + KB_OFF_X);           // [KB_NBUF][3][64]

    const int tid = (int)threadIdx.x, lane = tid & 31;
    const int warp = (int)warp_uniform((uint32_t)tid >> 5);     // provably warp-uniform (tc_ptx.cuh: elect_one)
    if (tid == 0) {
        mbar_init(BAR(0), 1);
        for (int b = 0; b < KB_NBUF; ++b) { mbar_init(BAR(1 + b), 1); mbar_init(BAR(4 + b), KB_CONV_THREADS); mbar_init(BAR(7 + b), 257); }
        mbar_init(BAR(10), 1); mbar_init(BAR(11), 1);
        mbar_init(BAR(12), 256); mbar_init(BAR(13), 256);
        mbar_init(BAR(14), 1);
        mbar_fence_init();
    }
    if (warp == 25) tmem_alloc<256>(smem_u32(tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = warp_uniform(*tmem_slot);

    const int G = (int)gridDim.x, cta = (int)blockIdx.x;
    const int t_begin = (int)(((long long)p.ntiles * cta) / G), t_end = (int)(((long long)p.ntiles * (cta + 1)) / G);
    // tuning aid (pgpd_debug_stream_counters): 0 loader wait buf_empty | 1 load latency (issue -> landed, seen by a converter)
    // 2 converter work | 3 mma wait op_ready | 4 mma issue | 5 epilogue wait acc_full | 6 epilogue work | 7 total
    long long* const dbg = g_stream_dbg ? g_stream_dbg + 256 * 8 : nullptr;      // rows 256.. of the debug buffer
    long long dacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    volatile long long* s_clk = reinterpret_cast<volatile long long*>(misc + 160);      // [KB_NBUF] issue time of the buffer's loads
    const long long tk0 = dbg ? clock64() : 0;
#define KB_T(slot, call) do { const long long _t0 = dbg ? clock64() : 0; call; if (dbg) dacc[slot] += clock64() - _t0; } while (0)

    if (warp == 24) {
        // ===================== loader =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(BAR(0), KB_A1_BYTES + KB_A2_BYTES);
            bulk_g2s(sbase, p.A1img, KB_A1_BYTES, BAR(0));
            bulk_g2s(sbase + KB_OFF_A2, p.A2img, KB_A2_BYTES, BAR(0));
            int i = 0;
            for (int t = t_begin; t < t_end; ++t, ++i) {
                const int b = i % KB_NBUF;
                const uint32_t ph = (uint32_t)(i / KB_NBUF) & 1u;
                {   // tile t+4 -> L2
                    const int tp = t + 4;
                    if (tp < t_end) {
                        const int cb = tp / p.tiles_per_cloud, tt = tp % p.tiles_per_cloud, n0 = tt * KB_NT;
                        const int nv = (p.N - n0 < KB_NT) ? p.N - n0 : KB_NT;
                        const size_t P0 = (size_t)cb * p.N + n0;
                        l2_prefetch(p.DZ2 + P0 * C2, (uint32_t)nv * C2 * 4u);
                    }
                }
                const int cb = t / p.tiles_per_cloud, tt = t % p.tiles_per_cloud, n0 = tt * KB_NT;
                const int nv = (p.N - n0 < KB_NT) ? p.N - n0 : KB_NT;
                const size_t P0 = (size_t)cb * p.N + n0;
                KB_T(0, mbar_wait(BAR(7 + b), ph ^ 1));
                if (dbg) s_clk[b] = clock64();
                mbar_arrive_expect_tx(BAR(1 + b), (uint32_t)nv * C2 * 4u);
                const uint32_t dst = sbase + KB_OFF_BUF + b * KB_BUF_BYTES;
                bulk_g2s(dst, p.DZ2 + P0 * C2, (uint32_t)nv * C2 * 4u, BAR(1 + b));
            }
            if (dbg) dbg[cta * 8 + 0] = dacc[0];
        }
    } else if (warp == 25) {
        // ===================== MMA issuer =====================
        {   // the whole warp runs the loop (uniform operands), one elected lane issues: tc_ptx.cuh: elect_one
            const bool el = elect_one();
            constexpr uint32_t IDESC_K = idesc_f16(128, KB_NT);
            constexpr uint32_t IDESC_MN = idesc_f16_mn(128, KB_NT);
            mbar_wait(BAR(0), 0);
            tc_fence_after_sync();
            uint32_t first = 1;
            const uint64_t dA1 = desc_sw128_kmajor(sbase), dA2 = desc_sw128_kmajor(sbase + KB_OFF_A2);
            int i = 0;
            for (int t = t_begin; t < t_end; ++t, ++i) {
                const int b = i % KB_NBUF, acc = i & 1;
                const uint32_t ph = (uint32_t)(i / KB_NBUF) & 1u, aph = (uint32_t)(i >> 1) & 1u;
                KB_T(3, mbar_wait(BAR(4 + b), ph));
                mbar_wait(BAR(12 + acc), aph ^ 1);
                tc_fence_after_sync();
                const long long ti0 = dbg ? clock64() : 0;
                const uint32_t dz = sbase + KB_OFF_BUF + b * KB_BUF_BYTES, a1 = dz + KB_DZ_BYTES;
                const uint32_t d1 = tmem + (uint32_t)(acc * KB_NT);
                // descriptors: one base per operand, every other one is base + (byte offset >> 4) in the address field
                const uint64_t kA1 = dA1, kA2 = dA2;
                const uint64_t kdz = desc_sw128_kmajor(dz), ka1 = desc_sw128_kmajor(a1);
                const uint64_t mdz = desc_sw128_mnmajor(dz, 8192), ma1 = desc_sw128_mnmajor(a1, 8192);
                // ---- D1 = A1op x dz (K = 128 channels)
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t oa = (uint32_t)((kb * 2 + (pass == 1 ? 1 : 0)) * KB_IMG_BLOCK);
                        const uint32_t ob = (uint32_t)(((pass == 2 ? 1 : 0) * 2 + kb) * 8192);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (el) mma_f16(d1, kA1 + ((oa + k * 32) >> 4), kdz + ((ob + k * 32) >> 4), IDESC_K, (kb | pass | k) ? 1u : 0u);
                    }
                }
                // ---- D1 += A2op x a1 (K = 64 channels)
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t oa = (pass == 1) ? (uint32_t)KB_IMG_BLOCK : 0u, ob = (pass == 2) ? 8192u : 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (el) mma_f16(d1, kA2 + ((oa + k * 32) >> 4), ka1 + ((ob + k * 32) >> 4), IDESC_K, 1u);
                }
                if (el) mma_commit(BAR(10 + acc));                  // d a1 of this tile complete -> epilogue
                // ---- D2 += dz^T a1 (K = 64 points, MN-major operands; A atoms = the two 64-channel blocks, 8 KB apart)
                {
                    const uint32_t d2 = tmem + 128u;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t oa = (pass == 1) ? 16384u : 0u, ob = (pass == 2) ? 8192u : 0u;
#pragma unroll
                        for (int k = 0; k < KB_NT / 16; ++k)
                            if (el) mma_f16(d2, mdz + ((oa + k * 2048) >> 4), ma1 + ((ob + k * 2048) >> 4), IDESC_MN,
                                    (first && pass == 0 && k == 0) ? 0u : 1u);
                    }
                }
                // ---- D3 += [a1_hi ; a1_lo]^T a1_hi, then ... a1_lo  (rows 0-63: hi.hi + hi.lo, rows 64-127: lo.hi + lo.lo)
                {
                    const uint32_t d3 = tmem + 192u;
#pragma unroll
                    for (int pass = 0; pass < 2; ++pass) {
                        const uint32_t ob = pass ? 8192u : 0u;
#pragma unroll
                        for (int k = 0; k < KB_NT / 16; ++k)
                            if (el) mma_f16(d3, ma1 + ((k * 2048) >> 4), ma1 + ((ob + k * 2048) >> 4), IDESC_MN,
                                    (first && pass == 0 && k == 0) ? 0u : 1u);
                    }
                }
                first = 0;
                if (el) mma_commit(BAR(7 + b));                     // the tensor core is done reading this buffer
                if (dbg) dacc[4] += clock64() - ti0;
            }
            if (el) mma_commit(BAR(14));
            if (dbg) { dbg[cta * 8 + 3] = dacc[3]; dbg[cta * 8 + 4] = dacc[4]; }
        }
    } else if (warp < 16 && (warp & 3) < 2) {
        // ===================== epilogue: feature k = TMEM lane, 16 of the tile's 64 points per warp =====================
        const int q = warp & 3, cgp = warp >> 2;
        const int k = q * 32 + lane;
        const float ginv = p.ginv[k], cv = p.cvec[k], be = p.beta1[k];
        const float gm = p.gamma1[k], g1inv = gm != 0.f ? 1.0f / gm : 0.f;
        const uint32_t koff = (uint32_t)((k & 7) * 2), kchunk = (uint32_t)(k >> 3);
        float s1 = 0.f, s2 = 0.f;
        int i = 0;
        for (int t = t_begin; t < t_end; ++t, ++i) {
            const int b = i % KB_NBUF, acc = i & 1;
            const uint32_t aph = (uint32_t)(i >> 1) & 1u;
            const int tt = t % p.tiles_per_cloud, n0 = tt * KB_NT;
            const int nv = (p.N - n0 < KB_NT) ? p.N - n0 : KB_NT;
            KB_T(5, mbar_wait(BAR(10 + acc), aph));
            tc_fence_after_sync();
            const long long te0 = dbg ? clock64() : 0;
            float v[16];
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * KB_NT + cgp * 16), v);
            tc_fence_before_sync();
            mbar_arrive(BAR(12 + acc));                     // the accumulator is in registers
            const unsigned char* a1b = smem + KB_OFF_BUF + b * KB_BUF_BYTES + KB_DZ_BYTES;
            const float* xb = sx + b * (3 * KB_NT);
            float h0 = 0.f, h1 = 0.f, h2 = 0.f;
            float av[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {                  // a1 of the 16 points, read back from the operand tile
                const int pp = cgp * 16 + j;
                const uint32_t off = (uint32_t)pp * 128u + ((kchunk ^ (uint32_t)(pp & 7)) << 4) + koff;
                av[j] = (__half2float(*reinterpret_cast<const __half*>(a1b + off)) +
                         __half2float(*reinterpret_cast<const __half*>(a1b + 8192 + off))) * (1.0f / ACT_SCALE);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int pp = cgp * 16 + j;
                const float a = av[j];                      // rows >= nv of the operand tile are zero -> dz1 = 0
                const float da1 = fmaf(v[j], ginv, cv);
                const float dz1 = a > 0.f ? da1 : 0.f;
                const float yh = (a - be) * g1inv;          // only used where dz1 != 0
                s1 += dz1;
                s2 = fmaf(dz1, yh, s2);
                h0 = fmaf(dz1, xb[pp], h0); h1 = fmaf(dz1, xb[KB_NT + pp], h1); h2 = fmaf(dz1, xb[2 * KB_NT + pp], h2);
            }
            (void)nv;
            float* h = p.Hpart + ((size_t)t * KB_EPI_GROUPS + cgp) * (C1 * 3) + k * 3;
            h[0] = h0; h[1] = h1; h[2] = h2;
            tc_fence_before_sync();
            mbar_arrive(BAR(7 + b));
            if (dbg) dacc[6] += clock64() - te0;
        }
        if (dbg && warp == 0 && lane == 0) { dbg[cta * 8 + 5] = dacc[5]; dbg[cta * 8 + 6] = dacc[6]; dbg[cta * 8 + 7] = clock64() - tk0; }
        float* o = p.bnpart + ((size_t)cta * KB_EPI_GROUPS + cgp) * 2 * C1;
        o[k] = s1; o[C1 + k] = s2;
    } else {
        // ===================== converters: raw fp32 rows -> hi/lo fp16 operand tiles, in place =====================
        const int cw = warp < 16 ? (warp >> 2) * 2 + (warp & 1) : warp - 8;        // 0..15
        const int ctid = cw * 32 + lane;
        const float4 e4 = *reinterpret_cast<const float4*>(p.esc + 4 * lane);
        // a1 is recomputed from the coordinates: my four channels 4 cg .. 4 cg + 3 (cg = lane & 15)
        float w1[4][3], sf1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = 4 * (lane & 15) + j;
            const float s16 = p.sc1[c] * ACT_SCALE;         // folded exactly as in tc_kf.cuh: the same a1 bit for bit
            w1[j][0] = s16 * p.W1[c * 3 + 0]; w1[j][1] = s16 * p.W1[c * 3 + 1]; w1[j][2] = s16 * p.W1[c * 3 + 2];
            sf1[j] = p.sh1[c] * ACT_SCALE;
        }
        int i = 0;
        for (int t = t_begin; t < t_end; ++t, ++i) {
            const int b = i % KB_NBUF;
            const uint32_t ph = (uint32_t)(i / KB_NBUF) & 1u;
            const int cb = t / p.tiles_per_cloud, tt = t % p.tiles_per_cloud, n0 = tt * KB_NT;
            const int nv = (p.N - n0 < KB_NT) ? p.N - n0 : KB_NT;
            float xv = 0.f;
            if (ctid < 3 * KB_NT) {
                const int j = ctid >> 6, pp = ctid & 63;
                if (pp < nv) xv = __ldg(p.x + (size_t)cb * 3 * p.N + (size_t)j * p.N + n0 + pp);
            }
            mbar_wait(BAR(1 + b), ph);
            const long long tc0 = dbg ? clock64() : 0;
            if (dbg) dacc[1] += tc0 - s_clk[b];
            unsigned char* dzb = smem + KB_OFF_BUF + b * KB_BUF_BYTES;
            unsigned char* a1b = dzb + KB_DZ_BYTES;
            float4 rdz[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) rdz[u] = *reinterpret_cast<const float4*>(dzb + (cw * 4 + u) * 512 + lane * 16);
            if (ctid < 3 * KB_NT) sx[b * (3 * KB_NT) + ctid] = xv;      // raw coordinates (the epilogue needs them too)
            named_bar_sync(1, KB_CONV_THREADS);             // every converter thread has read its raw rows; the coordinates are staged
            float T[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
            if (p.trans) {
#pragma unroll
                for (int e = 0; e < 9; ++e) T[e] = __ldg(p.trans + (size_t)cb * 9 + e);
            }
            {
                const int kb = lane >> 4, chunk = (lane & 15) >> 1, half8 = lane & 1;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = cw * 4 + u;
                    const bool ok = r < nv;
                    const float d0 = ok ? fminf(fmaxf(rdz[u].x * e4.x, -60000.f), 60000.f) : 0.f;
                    const float d1 = ok ? fminf(fmaxf(rdz[u].y * e4.y, -60000.f), 60000.f) : 0.f;
                    const float d2 = ok ? fminf(fmaxf(rdz[u].z * e4.z, -60000.f), 60000.f) : 0.f;
                    const float d3 = ok ? fminf(fmaxf(rdz[u].w * e4.w, -60000.f), 60000.f) : 0.f;
                    __half2 h01, l01, h23, l23;
                    split2(d0, d1, h01, l01);
                    split2(d2, d3, h23, l23);
                    const uint32_t off = (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4) + half8 * 8);
                    uint2 hv, lv;
                    hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
                    lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
                    *reinterpret_cast<uint2*>(dzb + (0 * 2 + kb) * 8192 + off) = hv;
                    *reinterpret_cast<uint2*>(dzb + (1 * 2 + kb) * 8192 + off) = lv;
                }
            }
            {
                const int cg = lane & 15, chunk = cg >> 1, half8 = cg & 1;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int r = cw * 4 + u * 2 + (lane >> 4);
                    const bool ok = r < nv;
                    // a1 * 2^4 = relu(w' . (T^T x) + shift'), exactly as the forward computed it (tc_kf.cuh)
                    const float* xb = sx + b * (3 * KB_NT);
                    const float p0 = xb[r], p1 = xb[KB_NT + r], p2 = xb[2 * KB_NT + r];
                    const float q0 = T[0] * p0 + T[3] * p1 + T[6] * p2, q1 = T[1] * p0 + T[4] * p1 + T[7] * p2,
                                q2 = T[2] * p0 + T[5] * p1 + T[8] * p2;
                    float av4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        av4[j] = ok ? fminf(relu_nan(fmaf(w1[j][0], q0, fmaf(w1[j][1], q1, fmaf(w1[j][2], q2, sf1[j])))), 60000.f) : 0.f;
                    }
                    __half2 h01, l01, h23, l23;
                    split2(av4[0], av4[1], h01, l01);
                    split2(av4[2], av4[3], h23, l23);
                    const uint32_t off = (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4) + half8 * 8);
                    uint2 hv, lv;
                    hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
                    lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
                    *reinterpret_cast<uint2*>(a1b + off) = hv;
                    *reinterpret_cast<uint2*>(a1b + 8192 + off) = lv;
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(BAR(4 + b));
            if (dbg) dacc[2] += clock64() - tc0;
        }
        if (dbg && cw == 0 && lane == 0) { dbg[cta * 8 + 1] = dacc[1]; dbg[cta * 8 + 2] = dacc[2]; }
    }
#undef KB_T

    // ===================== read-out of the persistent accumulators (warps 0..15: all four TMEM lane quadrants) ==========
    if (warp < 16) {
        const int q = warp & 3, cg = warp >> 2, row = q * 32 + lane;
        mbar_wait(BAR(14), 0);
        tc_fence_after_sync();
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + 128u + (uint32_t)(cg * 16), v);
        {
            const float sc = p.einv[row] * (1.0f / ACT_SCALE);
            float* out = p.Cpart + (size_t)cta * (C2 * C1) + (size_t)row * C1 + cg * 16;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(out + j) = make_float4(v[j] * sc, v[j + 1] * sc, v[j + 2] * sc, v[j + 3] * sc);
        }
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + 192u + (uint32_t)(cg * 16), v);
        float* scr = reinterpret_cast<float*>(smem + KB_OFF_BUF);      // the operand buffers are dead: [64][64] scratch
        if (q >= 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) scr[(row - 64) * C1 + cg * 16 + j] = v[j];
        }
        named_bar_sync(2, 512);
        if (q < 2) {
            float* out = p.G1part + (size_t)cta * (C1 * C1) + (size_t)row * C1 + cg * 16;
            constexpr float sc = 1.0f / (ACT_SCALE * ACT_SCALE);
#pragma unroll
            for (int j = 0; j < 16; ++j) out[j] = (v[j] + scr[row * C1 + cg * 16 + j]) * sc;
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 25) tmem_dealloc<256>(tmem);
}

inline int launch_kb(const KbParams& p, int sms, cudaStream_t s) {
    static int done[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!done[dev & 63]) {
        cudaFuncSetAttribute(k_kb_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, KB_SMEM_BYTES);
        done[dev & 63] = 1;
    }
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    launch(k_kb_tc, dim3(grid), dim3(KB_THREADS), (size_t)KB_SMEM_BYTES, s, p);
    return grid;
}

}}  // namespace pgpd::tc
