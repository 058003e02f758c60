// tc_l3.cuh -- tcgen05 kernel for the dominant op of the path (94 % of the tower flops):
//   u3[c][n] = sum_k W3[c][k] * a2[n][k]      (128 -> 1024 channels, every point of every cloud)
//   fused with: the BatchNorm2+ReLU prologue (a2 = relu(scale2*u2 + shift2) computed while staging),
//               the global max-pool (+ first arg-max) over the points of each cloud,
//               the centred sum of squares needed for the train-mode BatchNorm3 statistics.
// The 1024-wide activation lives only in TMEM.
//
// Numerics: fp32-grade.  Both operands are split x = hi + lo into two fp16 values (22 significant
// bits) and the product is formed as hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (3 MMAs
// at the 16-bit rate).  Operands are pre-scaled by powers of two (exact) so that the lo parts stay
// in fp16's normal range: every W3 row by 2^e_c with max|w_c|*2^e_c in [2^13,2^14), activations
// by 2^4.  Max/arg-max are invariant under the positive scaling; sums are rescaled on output.
//
// Orientation: channels on the MMA M axis (TMEM lanes), points on N (columns): the max / sum over
// points is then a per-thread serial reduction over the columns each epilogue thread loads.
//
// CTA PAIRS (cta_group::2), 26 warps per CTA, persistent over tiles of 256 points (128 per CTA):
//   warp 0      W3 producer: tensor-map (TMA) copies of 32 KB stages (this CTA's 128 channels x 64 k, hi+lo) of the pre-swizzled
//               weight image from L2 into a 3-deep ring; both CTAs' copies complete on the leader's barrier;
//   warp 1      MMA issuer of the leader CTA (whole warp runs the loop, one elected lane issues: tc_ptx.cuh): per 256-channel
//               block 2 k-blocks x 3 passes x 4 MMAs of 256 x 256 x 16 into one of two 256-column TMEM accumulators;
//   warps 2-17  epilogue: tcgen05.ld the accumulator, max / arg-max / centred squares per channel;
//   warps 18-25 a2 producers: u2 tile -> BN2+ReLU -> hi/lo fp16 -> swizzled shared memory (double-buffered).
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace pgpd { namespace tc {

// Tensor map over a pre-swizzled operand image seen as rows of 128 bytes: box = 256 rows = one 32 KB weight stage (hi + lo of a
// 128-channel x 64-k block).  Encoding is a host-side computation (no allocation, no synchronisation); the driver entry point is
// looked up once through the runtime, so the library has no link-time dependency on libcuda.
inline PFN_cuTensorMapEncodeTiled tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled enc = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
        else cudaGetLastError();
    }
    return enc;
}
inline bool make_image_map(CUtensorMap* m, const void* img, size_t bytes) {
    const PFN_cuTensorMapEncodeTiled enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t gdim[2] = {128, (cuuint64_t)(bytes / 128)};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {128, 256};
    const cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(img), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr int L3_NT = 256;                        // points per tile (MMA N)
constexpr int L3_STAGES = 3;
constexpr int L3_STAGE_BYTES = 2 * 128 * 128;     // hi + lo, 128 rows x 128 B
constexpr int L3_A2_PART = L3_NT * 128;           // one (part,kblock) sub-tile: 256 rows x 128 B = 32 KB
constexpr int L3_A2_BYTES = 4 * L3_A2_PART;       // hi/lo x 2 k-blocks = 128 KB
constexpr int L3_SMEM_W = L3_A2_BYTES;
constexpr int L3_SMEM_MISC = L3_SMEM_W + L3_STAGES * L3_STAGE_BYTES;
constexpr int L3_SMEM_BYTES = L3_SMEM_MISC + 2048 + 1024;   // + slack to align the base to 1024 B
constexpr int L3_THREADS = 448;                  // v2 kernel: 14 warps: W producer, MMA issuer, 4 epilogue, 8 a2 producers
#ifndef PGPD_L3_NP
#define PGPD_L3_NP 16
#endif
constexpr int L3A_NP = PGPD_L3_NP;               // a2 producer warps of the v1 kernel (8 or 16)
constexpr int L3A_THREADS = 320 + 32 * L3A_NP;   // v1 kernel: W producer, MMA issuer, 8 epilogue, L3A_NP a2 producer warps
constexpr float L3_ACT_SCALE = 16.0f;             // 2^4
constexpr size_t L3_WIMG_BYTES = (size_t)8 * 2 * L3_STAGE_BYTES;   // 512 KB

// The weight image (W3 [1024][128] fp32 -> swizzled hi/lo fp16, rows scaled by 2^e_c, sign(gamma3) folded in; index
// ((mt*2 + kb)*2 + part) * 16 KB + r*128 + ((chunk ^ (r&7)) << 4) + within*2) and inv[c] = sign(gamma3[c]) * 2^-(e_c + 4) are
// produced by tails.cuh: k_tower_pre.
struct L3Params {
    const float* Y2;          // [M][128] layer-2 pre-activation
    const float* scale2;      // [128]
    const float* shift2;      // [128]
    const __half* Wimg;       // pre-packed W3 image
    const float* inv;         // [1024]
    const float* mu_s;        // [1024] centre of the sums of squares (a pilot estimate of mean(u3)) in accumulator units, or nullptr (no statistics)
    unsigned long long* keys; // [B][1024] (ordered max value, ~arg-max)
    float* css_part;          // [ntiles][1024]
    int B, N, tiles_per_cloud, ntiles;
    long long* dbg;           // optional [gridDim.x][8] cycle counters (see PGPD_L3_DEBUG), or nullptr
    float* s1_part;           // optional partial sums over the points of a2 * 2^4: v1 [gridDim.x][128], v3 [gridDim.x * 8][128]
    unsigned* bad;            // [B] set to 1 for a cloud with a NaN / out-of-fp16-range activation (version 3)
};

// ====================================================================================================================
// CTA PAIRS with cta_group::2 MMAs (M = 256 channels x N = 256 points per instruction).
//   * each CTA of a pair holds HALF of the W3 block of a stage (its 128 of the 256 channels: 32 KB) and HALF of the a2
//     operand tile (its 128 of the pair's 256 points: 64 KB): the W3 stream per SM is halved and the a2 operand fits twice,
//     so staging the next tile overlaps the MMAs of the current one;
//   * the leader CTA (cluster rank 0) issues every MMA; tcgen05.commit arrives on BOTH CTAs' barriers (multicast);
//     barriers the leader waits on that depend on the peer (operand tile staged, accumulator drained) receive the peer's
//     arrivals through DSMEM (mapa + mbarrier.arrive.release.cluster); the weight stages need no relay: the peer's TMA copy
//     signals the leader's barrier directly (.cta_group::2);
//   * each CTA's epilogue drains its own 128 channels x 256 points from its own TMEM and hands the accumulator back as soon
//     as its last tcgen05.ld has landed, before the arithmetic.
// What round 2's measurements say about this kernel (profiles/r2/README.md, "layer-3 kernel: what bounds it"): the weight ring
// never runs dry (the issuing thread's "wait" cycles were MMA-queue back-pressure); knocking out the accumulator drain saves 21 %,
// the operand staging 5 %, the weight stream 0 %; and issuing the MMAs under `if (lane == 0)` made every tcgen05.mma cost ~130
// cycles of issue (per-instruction R2UR waterfall) -- as long as its own 128.6 cycles of execution.
// ====================================================================================================================
constexpr int L3C_NH = 128;                        // points staged per CTA (half of the pair's tile)
constexpr int L3C_A2_PART = L3C_NH * 128;          // 16 KB: one (part, k-block) sub-tile
constexpr int L3C_A2_BUF = 4 * L3C_A2_PART;        // 64 KB
constexpr int L3C_SMEM_W = 2 * L3C_A2_BUF;         // 128 KB
constexpr int L3C_SMEM_MISC = L3C_SMEM_W + L3_STAGES * L3_STAGE_BYTES;
constexpr int L3C_SMEM_BYTES = L3C_SMEM_MISC + 2048 + 1024;
constexpr int L3C_NPROD = 8;                       // a2 producer warps (measured with 4 / 2 and both tcgen05.ld of an epilogue warp in flight: the
                                                   // register allocator still spills at 80 / 96 registers, 2.6-2.9x slower; profiles/r2)
constexpr int L3C_THREADS = (2 + 16 + L3C_NPROD) * 32;   // W producer, MMA issuer, 16 epilogue, L3C_NPROD a2 producer warps
constexpr int L3C_EPI_ROWS = 4;                    // partial rows of centred squares per tile (one per 64-column quarter)

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(L3C_THREADS, 1) k_l3_fwd_tc3(L3Params p, const __grid_constant__ CUtensorMap wmap) {
    constexpr int NSUB = L3_STAGES;                        // ring slots
    constexpr int SUB_BYTES = L3_STAGE_BYTES;
    constexpr int W_FULL = 0, W_EMPTY = 12, A2_FULL = 18, A2_EMPTY = 20, TM_FULL = 22, TM_EMPTY = 24;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    unsigned char* misc = smem + L3C_SMEM_MISC;
    const uint32_t bar0 = sbase + L3C_SMEM_MISC;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    // W_FULL.. (leader: both CTAs' tensor-map copies of a weight stage have landed) | W_EMPTY.. (commit, both CTAs)
    // A2_FULL (leader: 16 producer warps of both CTAs) | A2_EMPTY (commit, both) | TM_FULL (commit, both)
    // TM_EMPTY (leader: 2 x 16 epilogue warps of both CTAs)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 240);
    float* s_scale = reinterpret_cast<float*>(misc + 256);
    float* s_shift = s_scale + 128;

    const int tid = (int)threadIdx.x, lane = tid & 31;
    const int warp = (int)warp_uniform((uint32_t)tid >> 5);     // provably warp-uniform: role dispatch without divergence
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    long long gt0 = 0, ck0 = 0;
    if (p.dbg && tid == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0)); ck0 = clock64(); }

    if (tid == 0) {
        for (int i = 0; i < NSUB; ++i) { mbar_init(BAR(W_FULL + i), 1); mbar_init(BAR(W_EMPTY + i), 1); }
        mbar_init(BAR(A2_FULL), 2 * L3C_NPROD); mbar_init(BAR(A2_FULL + 1), 2 * L3C_NPROD);
        mbar_init(BAR(A2_EMPTY), 1); mbar_init(BAR(A2_EMPTY + 1), 1);
        mbar_init(BAR(TM_FULL), 1); mbar_init(BAR(TM_FULL + 1), 1);
        mbar_init(BAR(TM_EMPTY), 32); mbar_init(BAR(TM_EMPTY + 1), 32);
        mbar_fence_init();
    }
    if (tid < 128) {
        s_scale[tid] = p.scale2[tid] * L3_ACT_SCALE; s_shift[tid] = p.shift2[tid] * L3_ACT_SCALE;
    }
    if (warp == 1) tmem_alloc_pair<512>(smem_u32(tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                 // both CTAs' barriers exist before anything is signalled across
    tc_fence_after_sync();
    const uint32_t tmem = warp_uniform(*tmem_slot);

    // tiles of this PAIR (256 points each); both CTAs walk the same tiles
    const int npairs = (int)gridDim.x >> 1, pair = (int)blockIdx.x >> 1;
    const int T0 = (int)(((long long)p.ntiles * pair) / npairs), T1 = (int)(((long long)p.ntiles * (pair + 1)) / npairs);

    if (warp == 0) {
        // ===================== W3 producer: this CTA's 128 of the 256 channels of every stage =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = T0; t < T1; ++t)
                for (int mt4 = 0; mt4 < 4; ++mt4)
                    for (int kb = 0; kb < 2; ++kb) {
                        const int blk = ((((mt4 + pair) & 3) * 2 + (int)rank) * 2 + kb);      // pairs walk the channel blocks in different rotations
                        mbar_wait(BAR(W_EMPTY + stage), phase ^ 1);
                        // tensor-map copy of my 32 KB half of the stage; BOTH CTAs' copies complete on the LEADER's barrier, which
                        // expects the whole 64 KB: the MMA issuer waits on one local barrier, nothing is relayed through the peer
                        // (round 1 / early round 2: plain bulk copies + a relay thread in the peer: +4 % kernel time)
                        if (leader) mbar_arrive_expect_tx(BAR(W_FULL + stage), 2 * SUB_BYTES);
                        tma2d_g2s_pair_leaderbar(sbase + L3C_SMEM_W + stage * SUB_BYTES, &wmap, 0, blk * 256, BAR(W_FULL + stage));
                        if (++stage == NSUB) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 1) {
        if (leader) {
            // ===================== leader: MMA issuer for the pair (WHOLE WARP, one elected lane issues; tc_ptx.cuh: elect_one) =====
            // (Measured and dropped, profiles/r2/README.md: testing the next stage's barriers one unit ahead, a 16 KB sub-stage ring,
            // four 128-column accumulators -- twice the MMA instructions.)
            constexpr uint32_t IDESC = idesc_f16(256, L3_NT);
            constexpr uint32_t OB_LO = (uint32_t)(2 * L3C_A2_PART);      // the a2 lo part
            int stage = 0; uint32_t wphase = 0;
            int acc = 0; uint32_t aphase = 0;
            int buf = 0; uint32_t bphase = 0;
            long long w_a2 = 0, w_acc = 0, w_w = 0;
            const long long tl0 = p.dbg ? clock64() : 0;
            for (int t = T0; t < T1; ++t) {
                { const long long _t = p.dbg ? clock64() : 0; mbar_wait_cluster(BAR(A2_FULL + buf), bphase); if (p.dbg) w_a2 += clock64() - _t; }   // both halves staged
                tc_fence_after_sync();
                const uint32_t a2b = sbase + buf * L3C_A2_BUF;
                for (int mt4 = 0; mt4 < 4; ++mt4) {
                    { const long long _t = p.dbg ? clock64() : 0; mbar_wait_cluster(BAR(TM_EMPTY + acc), aphase ^ 1); if (p.dbg) w_acc += clock64() - _t; }   // drained
                    tc_fence_after_sync();
                    const uint32_t d = tmem + (uint32_t)(acc * L3_NT);
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint64_t db = desc_sw128_kmajor(a2b + kb * L3C_A2_PART);
                        { const long long _t = p.dbg ? clock64() : 0; mbar_wait(BAR(W_FULL + stage), wphase); if (p.dbg) w_w += clock64() - _t; }   // weight slot (both halves)
                        tc_fence_after_sync();
                        const uint64_t dw = desc_sw128_kmajor(sbase + L3C_SMEM_W + stage * L3_STAGE_BYTES);
                        if (elect_one()) {
#pragma unroll
                            for (int pass = 0; pass < 3; ++pass) {
                                const uint32_t oa = (pass == 1) ? 16384u : 0u;
                                const uint32_t ob = (pass == 2) ? OB_LO : 0u;
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    mma_f16_pair(d, dw + ((oa + k * 32) >> 4), db + ((ob + k * 32) >> 4), IDESC, (kb | pass | k) ? 1u : 0u);
                            }
                            mma_commit_pair(BAR(W_EMPTY + stage), (uint16_t)0x3);      // slot free in both CTAs
                            if (kb == 1) mma_commit_pair(BAR(TM_FULL + acc), (uint16_t)0x3);                 // accumulator complete in both CTAs
                            if (kb == 1 && mt4 == 3) mma_commit_pair(BAR(A2_EMPTY + buf), (uint16_t)0x3);   // operand buffer free in both CTAs
                        }
                        __syncwarp();
                        if (++stage == L3_STAGES) { stage = 0; wphase ^= 1u; }
                    }
                    if (++acc == 2) { acc = 0; aphase ^= 1; }
                }
                if (++buf == 2) { buf = 0; bphase ^= 1; }
            }
            if (p.dbg && lane == 0) {
                long long* o = p.dbg + (size_t)blockIdx.x * 8;
                o[0] = w_a2; o[1] = w_acc; o[2] = w_w; o[3] = clock64() - tl0; o[4] = 0; o[5] = 0; o[6] = 0; o[7] = 0;
            }
        }
    } else if (warp < 18) {
        // ===================== epilogue (16 warps: TMEM lane quadrant x 64-column quarter): my 128 channels of every
        // 256-channel block, all 256 points of the tile =====================
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;                   // 0..3: which quarter of the 256 columns
        const int row = q * 32 + lane;
        const bool stats = p.mu_s != nullptr;
        int acc = 0; uint32_t aphase = 0;
        float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f, cs3 = 0.f;      // centred squares of my channel of block mt4 = 0..3, summed over my tiles
        for (int t = T0; t < T1; ++t) {
            const int b = t / p.tiles_per_cloud, tt = t % p.tiles_per_cloud;
            const int n0 = tt * L3_NT;
            const int nvalid = (p.N - n0 < L3_NT) ? p.N - n0 : L3_NT;
            for (int mt4 = 0; mt4 < 4; ++mt4) {
                const int ch = (((mt4 + pair) & 3) * 2 + (int)rank) * 128 + row;
                const float mu = stats ? p.mu_s[ch] : 0.f;
                const uint64_t nmu2 = f2_pack(-mu, -mu);
                float best = -INFINITY; int bidx = 0; float css = 0.f;
                // my two chunks of 32 columns (points half*64 + u*32 of the tile) of this block's accumulator
                auto chunk = [&](const float (&v)[32], int pb) {
                    if (pb + 32 <= nvalid) {
                        float m0 = v[0], m1 = v[1], m2 = v[2], m3 = v[3];
#pragma unroll
                        for (int j = 4; j < 32; j += 4) {
                            m0 = fmaxf(m0, v[j]); m1 = fmaxf(m1, v[j + 1]); m2 = fmaxf(m2, v[j + 2]); m3 = fmaxf(m3, v[j + 3]);
                        }
                        const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                        if (stats) {
                            uint64_t q0 = 0ull, q1 = 0ull;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const uint64_t d0 = f2_add(f2_pack(v[j], v[j + 1]), nmu2), d1 = f2_add(f2_pack(v[j + 2], v[j + 3]), nmu2);
                                q0 = f2_fma(d0, d0, q0); q1 = f2_fma(d1, d1, q1);
                            }
                            float c0s, c1s, c2s, c3s;
                            f2_unpack(q0, c0s, c1s); f2_unpack(q1, c2s, c3s);
                            css += (c0s + c1s) + (c2s + c3s);
                        }
                        if (m > best) {
                            best = m;
                            int jj = 31;
#pragma unroll
                            for (int j = 30; j >= 0; --j) if (v[j] == m) jj = j;
                            bidx = n0 + pb + jj;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (pb + j < nvalid) {
                                if (stats) { const float dlt = v[j] - mu; css = fmaf(dlt, dlt, css); }
                                if (v[j] > best) { best = v[j]; bidx = n0 + pb + j; }
                            }
                        }
                    }
                };
                auto release = [&]() {
                    // everything I need of this accumulator is in registers: hand it back BEFORE the arithmetic (drain time and MMA
                    // time per block are nearly equal: every cycle between the last tcgen05.ld and this arrival is on the critical path)
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) { if (leader) mbar_arrive(BAR(TM_EMPTY + acc)); else mbar_arrive_cluster(BAR(TM_EMPTY + acc), 0u); }
                };
                const uint32_t tcol = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * L3_NT + half * 64);
                const int pb0 = half * 64;
                mbar_wait(BAR(TM_FULL + acc), aphase);
                tc_fence_after_sync();
                {
                    float v[32];
                    if (pb0 < nvalid) tmem_ld32(tcol, v);
                    const bool more = pb0 + 32 < nvalid;            // warp-uniform
                    if (!more) release();
                    if (pb0 < nvalid) chunk(v, pb0);
                    if (more) {
                        tmem_ld32(tcol + 32u, v);
                        release();
                        chunk(v, pb0 + 32);
                    }
                }
                if (++acc == 2) { acc = 0; aphase ^= 1; }
                const unsigned long long key = ((unsigned long long)ord_encode(best) << 32) |
                                               (unsigned long long)(0xFFFFFFFFu - (unsigned)bidx);
                atomicMax(&p.keys[(size_t)b * C3 + ch], key);
                if (mt4 == 0) cs0 += css; else if (mt4 == 1) cs1 += css; else if (mt4 == 2) cs2 += css; else cs3 += css;
            }
        }
        if (stats) {
            // four partial rows per PAIR (one per 64-column quarter); each CTA of the pair fills its 512 channels of them
#pragma unroll
            for (int mt4 = 0; mt4 < 4; ++mt4) {
                const int ch = (((mt4 + pair) & 3) * 2 + (int)rank) * 128 + row;
                const float iv = p.inv[ch];
                const float cs = mt4 == 0 ? cs0 : (mt4 == 1 ? cs1 : (mt4 == 2 ? cs2 : cs3));
                p.css_part[((size_t)pair * L3C_EPI_ROWS + half) * C3 + ch] = cs * iv * iv;
            }
        }
    } else {
        // ===================== a2 producer: my 128 of the tile's 256 points, double-buffered =====================
        const int wp = warp - 18;                           // 0 .. L3C_NPROD-1
        const int kb = lane >> 4, chunk = (lane & 15) >> 1, half8 = lane & 1;
        const float sc0 = s_scale[4 * lane + 0], sc1 = s_scale[4 * lane + 1], sc2 = s_scale[4 * lane + 2], sc3 = s_scale[4 * lane + 3];
        const float sh0 = s_shift[4 * lane + 0], sh1 = s_shift[4 * lane + 1], sh2 = s_shift[4 * lane + 2], sh3 = s_shift[4 * lane + 3];
        float sa0 = 0.f, sa1 = 0.f, sa2 = 0.f, sa3 = 0.f;
        bool oor = false;
        int buf = 0; uint32_t bphase = 0;
        for (int t = T0; t < T1; ++t) {
            const int b = t / p.tiles_per_cloud, tt = t % p.tiles_per_cloud;
            const int n0 = tt * L3_NT + (int)rank * L3C_NH;
            int nvalid = p.N - n0;                          // valid rows of MY half (<= 0: none)
            nvalid = nvalid < 0 ? 0 : (nvalid > L3C_NH ? L3C_NH : nvalid);
            const float* src = p.Y2 + ((size_t)b * p.N + n0) * C2 + 4 * lane;
            if (wp == 0 && lane == 0 && t + 2 < T1) {
                const int t2 = t + 2;                       // two tiles ahead -> L2
                const int b2 = t2 / p.tiles_per_cloud, tt2 = t2 % p.tiles_per_cloud;
                const int m0 = tt2 * L3_NT + (int)rank * L3C_NH;
                int nv2 = p.N - m0;
                nv2 = nv2 < 0 ? 0 : (nv2 > L3C_NH ? L3C_NH : nv2);
                if (nv2 > 0) l2_prefetch(p.Y2 + ((size_t)b2 * p.N + m0) * C2, (uint32_t)nv2 * C2 * 4u);
            }
            mbar_wait(BAR(A2_EMPTY + buf), bphase ^ 1);     // the MMAs that read this buffer two tiles ago are done
            unsigned char* a2b = smem + buf * L3C_A2_BUF;
            constexpr int U = 8;
            for (int i0 = 0; i0 < L3C_NH / L3C_NPROD; i0 += U) {
                float4 y[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int r = wp + L3C_NPROD * (i0 + u);
                    y[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < nvalid) y[u] = *reinterpret_cast<const float4*>(src + (size_t)r * C2);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int r = wp + L3C_NPROD * (i0 + u);
                    const bool ok = r < nvalid;
                    // relu that keeps NaN; values beyond the fp16 operand range (or NaN) flag the cloud (its pooled feature
                    // becomes NaN in k_tail_l3) instead of being clamped silently
                    float a0 = ok ? relu_nan(fmaf(sc0, y[u].x, sh0)) : 0.f;
                    float a1 = ok ? relu_nan(fmaf(sc1, y[u].y, sh1)) : 0.f;
                    float a2 = ok ? relu_nan(fmaf(sc2, y[u].z, sh2)) : 0.f;
                    float a3 = ok ? relu_nan(fmaf(sc3, y[u].w, sh3)) : 0.f;
                    // all four are >= 0 (or NaN): their sum within the fp16 range implies each one is, and NaN fails the comparison
                    oor = oor || !((a0 + a1) + (a2 + a3) <= 60000.f);
                    sa0 += a0; sa1 += a1; sa2 += a2; sa3 += a3;
                    __half2 h01, l01, h23, l23;
                    split2(a0, a1, h01, l01);
                    split2(a2, a3, h23, l23);
                    const uint32_t off = (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4) + half8 * 8);
                    uint2 hv, lv;
                    hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
                    lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
                    *reinterpret_cast<uint2*>(a2b + (0 * 2 + kb) * L3C_A2_PART + off) = hv;
                    *reinterpret_cast<uint2*>(a2b + (1 * 2 + kb) * L3C_A2_PART + off) = lv;
                }
            }
            if (oor) { p.bad[b] = 1u; oor = false; }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {                                // one arrival per warp on the LEADER's barrier
                if (leader) mbar_arrive(BAR(A2_FULL + buf)); else mbar_arrive_cluster(BAR(A2_FULL + buf), 0u);
            }
            if (++buf == 2) { buf = 0; bphase ^= 1; }
        }
        if (p.s1_part) {
            // sums of a2 (x 2^4): ONE partial row per CTA, [gridDim.x][128].  The 8 producer warps add their rows in warp order
            // through the first operand buffer, which is free once the MMAs of the last two tiles have released both buffers.
            for (int e = 0; e < 2; ++e) {
                mbar_wait(BAR(A2_EMPTY + buf), bphase ^ 1);
                if (++buf == 2) { buf = 0; bphase ^= 1; }
            }
            float* red = reinterpret_cast<float*>(smem);
            float* o = red + wp * C2 + 4 * lane;
            o[0] = sa0; o[1] = sa1; o[2] = sa2; o[3] = sa3;
            named_bar_sync(3, L3C_NPROD * 32);
            if (wp == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float t = red[4 * lane + i];
#pragma unroll
                    for (int w2 = 1; w2 < L3C_NPROD; ++w2) t += red[w2 * C2 + 4 * lane + i];
                    p.s1_part[(size_t)blockIdx.x * C2 + 4 * lane + i] = t;
                }
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                 // nobody exits while the peer may still signal this CTA
    if (p.dbg && tid == 0) {
        long long gt1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        p.dbg[(size_t)(200 + (blockIdx.x & 31)) * 8 + 0] = gt1 - gt0;
        p.dbg[(size_t)(200 + (blockIdx.x & 31)) * 8 + 1] = clock64() - ck0;
    }
    if (warp == 1) tmem_dealloc_pair<512>(tmem);
}

// per-device one-time setup: is this an sm_100 part, and can the kernel have its shared memory?
struct DevInfo { int state = 0; int sms = 148; };   // state: 0 unknown, 1 usable, -1 not usable
inline DevInfo& dev_info() {
    static DevInfo info[64];
    int dev = 0;
    cudaGetDevice(&dev);
    DevInfo& d = info[dev & 63];
    if (d.state == 0) {
        int major = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
        if (d.sms <= 0) d.sms = 148;
        if (d.sms > 256) d.sms = 256;      // per-CTA partial buffers are sized for <= 256 CTAs (plan_tower_scratch)
        cudaError_t e = cudaFuncSetAttribute(k_l3_fwd_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, L3C_SMEM_BYTES);
        d.state = (major == 10 && e == cudaSuccess && tensor_map_encoder() != nullptr) ? 1 : -1;   // the weight stream needs the driver's tensor-map encoder
        if (e != cudaSuccess) cudaGetLastError();
    }
    return d;
}
inline bool available() { return dev_info().state == 1; }

// tuning aid, only in builds with -DPGPD_DEBUG (never in the product library: the ABI promises no allocation, no global
// state): a lazily cudaMalloc'ed [512][8] int64 buffer (rows 256.. : the layer-2/1 backward kernel) of pipeline cycle counters, used when PGPD_L3_DEBUG is set
#ifdef PGPD_DEBUG
inline long long* l3_debug_buffer() {
    static long long* buf = nullptr;
    if (!buf) { cudaMalloc(&buf, 512 * 8 * sizeof(long long)); cudaMemset(buf, 0, 512 * 8 * sizeof(long long)); }
    return buf;
}
inline long long* l3_debug_buffer_if_enabled() {
    static const bool on = getenv("PGPD_L3_DEBUG") != nullptr;
    return on ? l3_debug_buffer() : nullptr;
}
#else
inline long long* l3_debug_buffer_if_enabled() { return nullptr; }
#endif

}}  // namespace pgpd::tc
