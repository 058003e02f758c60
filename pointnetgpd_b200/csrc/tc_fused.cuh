// tc_fused.cuh -- the WHOLE tower 3 -> 64 -> 128 -> 1024 + global max-pool of an eval-mode forward in ONE kernel:
//
//     x [B][3][N]  --fp32 FMA-->  a1 (64)  --tcgen05-->  u2 (128)  --BN2+ReLU-->  a2  --tcgen05-->  u3 (1024)  --max over points
//
// (STN3d.forward lines 29-33 / PointNetfeat.forward lines 140-149 of PointNetGPD/model/pointnet.py with the BatchNorms folded
// into per-channel scale/shift from the running statistics.)  Neither a1 nor u2 / a2 nor the 1024-wide activation ever
// leaves the SM: per point the kernel reads 12 bytes of coordinates, per (cloud, channel) it writes one 8-byte key.
//
// Structure = the layer-3 kernel of tc_l3.cuh (CTA pairs, cta_group::2 MMAs of 256 x 256 x 16, W3 image streamed through a
// 3-stage bulk-copy ring, two 256-column TMEM accumulators, 16 epilogue warps) with the a2 operand produced on chip:
//   * the 8 "producer" warps of each CTA compute a1 for the CTA's 128 points of the NEXT tile on the CUDA cores (3 FMAs per
//     output, transform and BatchNorm1 folded in) and write it as the hi/lo fp16 operand tile [128 points][64 k];
//   * the issuing thread runs layer 2 as one more accumulator use per tile: 12 MMAs of 256 points x 128 channels x 16 (A = the
//     pair's a1 tiles, B = the W2 image, half of its rows resident in each CTA) into whichever accumulator is free;
//   * the producer warps drain that accumulator (tcgen05.ld: lane = point, columns = channels), apply BatchNorm2 + ReLU, split
//     to hi/lo fp16 and write the a2 operand tile in place -- the tile layer 3 then consumes exactly as in tc_l3.cuh.
// Accumulator uses per tile: L2, blk0, blk1, blk2, blk3 (5: the two TMEM slots keep alternating); a slot is released by
// whoever drained it (producers after an L2 use, epilogue warps after a layer-3 block).  The a2 tile is single-buffered
// (shared memory: 64 KB a2 + 32 KB a1 + 16 KB W2 half + 96 KB W3 ring), so the tensor pipe idles while the L2 result of the
// next tile is drained (~3 k of ~21 k cycles per tile) -- the price for not writing 768 B per point to HBM and reading it back.
// (Tried: layer 2 of the next tile ahead of block 3 with the a2 tile handed over per k-block.  With two accumulator slots block 3
// then lands in the slot block 2 has just filled and waits for its drain; measured 5 % slower, profiles/r2/README.md.)
// Numerics as everywhere: fp32-grade 3-pass hi/lo fp16 split, power-of-two operand scales.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_l3.cuh"

namespace pgpd { namespace tc {

constexpr int FZ_A1 = L3C_A2_BUF;                          // a2 tile at 0 (64 KB), then the a1 tile [part][128 rows][128 B]
constexpr int FZ_W2 = FZ_A1 + 32768;                       // this CTA's 64 rows of the W2 image [part][64 rows][128 B]
constexpr int FZ_W = FZ_W2 + 16384;                        // W3 ring
constexpr int FZ_CONST = FZ_W + L3_STAGES * L3_STAGE_BYTES;    // W1 (192), scale1 (64), shift1 (64), s2' (128), h2' (128) floats
constexpr int FZ_MISC = FZ_CONST + 3072;
constexpr int FZ_SMEM_BYTES = FZ_MISC + 1024 + 1024;       // + slack to align the base to 1024 B
constexpr int FZ_THREADS = 832;                               // W producer, MMA issuer, 16 epilogue, 8 producer warps

struct FusedParams {
    const float* x; const float* trans;       // [B][3][N]; [B][9] or null (identity)
    const float* W1; const float* sc1; const float* sh1;      // conv1.weight [64][3], folded BatchNorm1
    const __half* W2img; const float* inv2;   // conv2 image (tails.cuh: k_tower_pre) and its per-row inverse scale (incl. 2^-4)
    const float* sc2; const float* sh2;       // folded BatchNorm2
    const __half* W3img; const float* inv3;   // conv3 image, inv[c] = sign(gamma3) 2^-(e_c + 4)
    unsigned long long* keys;                 // [B][1024] (ordered max value, ~arg-max), zeroed by the caller
    int B, N, tiles_per_cloud, ntiles;
    unsigned* bad;                            // [B] flags: NaN / out-of-fp16-range activation in this cloud
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FZ_THREADS, 1) k_tower_fused_eval(FusedParams p, const __grid_constant__ CUtensorMap wmap) {
    constexpr int W_FULL = 0, W_EMPTY = 6, A1_FULL = 9, A2_FULL = 10, TM2_FULL = 11, TM2_EMPTY = 12, TM_FULL = 14, TM_EMPTY = 16;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    unsigned char* misc = smem + FZ_MISC;
    const uint32_t bar0 = sbase + FZ_MISC;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 240);
    float* s_w1 = reinterpret_cast<float*>(smem + FZ_CONST);       // [64][3]
    float* s_sc1 = s_w1 + 192;
    float* s_sh1 = s_sc1 + 64;
    float* s_s2 = s_sh1 + 64;                                      // scale2 * 16 * inv2: accumulator of layer 2 -> a2 * 16
    float* s_h2 = s_s2 + 128;

    const int tid = (int)threadIdx.x, lane = tid & 31;
    const int warp = (int)warp_uniform((uint32_t)tid >> 5);     // provably warp-uniform (tc_ptx.cuh: elect_one)
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (tid == 0) {
        for (int i = 0; i < 3; ++i) { mbar_init(BAR(W_FULL + i), 1); mbar_init(BAR(W_EMPTY + i), 1); }
        mbar_init(BAR(A1_FULL), 16); mbar_init(BAR(A2_FULL), 16);
        mbar_init(BAR(TM2_FULL), 1);
        mbar_init(BAR(TM2_EMPTY), 16); mbar_init(BAR(TM2_EMPTY + 1), 16);
        mbar_init(BAR(TM_FULL), 1); mbar_init(BAR(TM_FULL + 1), 1);
        mbar_init(BAR(TM_EMPTY), 32); mbar_init(BAR(TM_EMPTY + 1), 32);
        mbar_fence_init();
    }
    if (tid < 192) s_w1[tid] = p.W1[tid];
    if (tid < 64) { s_sc1[tid] = p.sc1[tid]; s_sh1[tid] = p.sh1[tid]; }
    if (tid < 128) { s_s2[tid] = p.sc2[tid] * ACT_SCALE * p.inv2[tid]; s_h2[tid] = p.sh2[tid] * ACT_SCALE; }
    {
        // my 64 rows (channels 64*rank ..) of the W2 image: hi rows at 0, lo rows at 16 KB of the global image
        const uint4* src_hi = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.W2img) + (size_t)rank * 8192);
        const uint4* src_lo = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.W2img) + 16384 + (size_t)rank * 8192);
        uint4* dst = reinterpret_cast<uint4*>(smem + FZ_W2);
        for (int i = tid; i < 512; i += FZ_THREADS) { dst[i] = src_hi[i]; dst[512 + i] = src_lo[i]; }
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc_pair<512>(smem_u32(tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                 // both CTAs' barriers and W2 halves exist before anything is signalled across
    tc_fence_after_sync();
    const uint32_t tmem = warp_uniform(*tmem_slot);

    const int npairs = (int)gridDim.x >> 1, pair = (int)blockIdx.x >> 1;
    const int T0 = (int)(((long long)p.ntiles * pair) / npairs), T1 = (int)(((long long)p.ntiles * (pair + 1)) / npairs);

    if (warp == 0) {
        // ===================== W3 producer: this CTA's 128 of the 256 channels of every stage =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = T0; t < T1; ++t)
                for (int mt4 = 0; mt4 < 4; ++mt4)
                    for (int kb = 0; kb < 2; ++kb) {
                        const int blk = ((((mt4 + pair) & 3) * 2 + (int)rank) * 2 + kb);
                        mbar_wait(BAR(W_EMPTY + stage), phase ^ 1);
                        // as in tc_l3.cuh: both CTAs' tensor-map copies complete on the LEADER's barrier
                        if (leader) mbar_arrive_expect_tx(BAR(W_FULL + stage), 2 * L3_STAGE_BYTES);
                        tma2d_g2s_pair_leaderbar(sbase + FZ_W + stage * L3_STAGE_BYTES, &wmap, 0, blk * 256, BAR(W_FULL + stage));
                        if (++stage == 3) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 1) {
        if (leader) {
            // ===================== leader: MMA issuer for the pair (whole warp, one elected lane issues: tc_ptx.cuh) =====================
            constexpr uint32_t IDESC3 = idesc_f16(256, L3_NT);
            constexpr uint32_t IDESC2 = idesc_f16(256, 128);
            constexpr uint32_t OB_LO = (uint32_t)(2 * L3C_A2_PART);
            int stage = 0; uint32_t wphase = 0;
            uint32_t ph_e[2] = {0u, 0u}, ph_p[2] = {0u, 0u};       // next completion parity of TM_EMPTY[s] / TM2_EMPTY[s]
            uint32_t ph_a1 = 0u, ph_a2 = 0u;
            long long use = 0;                                     // accumulator uses so far: slot = use & 1
            auto wait_slot = [&]() {
                if (use < 2) return;
                const int s = (int)(use & 1);
                if ((use - 2) % 5 == 0) { mbar_wait_cluster(BAR(TM2_EMPTY + s), ph_p[s]); ph_p[s] ^= 1u; }   // drained by the producers
                else { mbar_wait_cluster(BAR(TM_EMPTY + s), ph_e[s]); ph_e[s] ^= 1u; }                      // drained by the epilogue
                tc_fence_after_sync();
            };
            const uint32_t a1s = sbase + FZ_A1, w2s = sbase + FZ_W2, a2b = sbase;
            for (int t = T0; t < T1; ++t) {
                // ---- layer 2 of this tile: D[256 points][128 channels] = a1 (hi, lo) x W2 (hi, lo), K = 64
                mbar_wait_cluster(BAR(A1_FULL), ph_a1); ph_a1 ^= 1u;
                tc_fence_after_sync();
                wait_slot();
                {
                    const uint32_t d2 = tmem + (uint32_t)((use & 1) * L3_NT);
                    const uint64_t da = desc_sw128_kmajor(a1s), db = desc_sw128_kmajor(w2s);
                    if (elect_one()) {
#pragma unroll
                        for (int pass = 0; pass < 3; ++pass) {
                            const uint32_t oa = (pass == 1) ? 16384u : 0u;      // a1 lo
                            const uint32_t ob = (pass == 2) ? 8192u : 0u;       // W2 lo
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                mma_f16_pair(d2, da + ((oa + k * 32) >> 4), db + ((ob + k * 32) >> 4), IDESC2, (pass | k) ? 1u : 0u);
                        }
                        mma_commit_pair(BAR(TM2_FULL), (uint16_t)0x3);
                    }
                    __syncwarp();
                    ++use;
                }
                // ---- layer 3: four 256-channel blocks
                mbar_wait_cluster(BAR(A2_FULL), ph_a2); ph_a2 ^= 1u;
                tc_fence_after_sync();
                for (int mt4 = 0; mt4 < 4; ++mt4) {
                    wait_slot();
                    const int s = (int)(use & 1);
                    const uint32_t d = tmem + (uint32_t)(s * L3_NT);
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint64_t dbk = desc_sw128_kmajor(a2b + kb * L3C_A2_PART);
                        mbar_wait(BAR(W_FULL + stage), wphase);
                        tc_fence_after_sync();
                        const uint64_t dw = desc_sw128_kmajor(sbase + FZ_W + stage * L3_STAGE_BYTES);
                        if (elect_one()) {
#pragma unroll
                            for (int pass = 0; pass < 3; ++pass) {
                                const uint32_t oa = (pass == 1) ? 16384u : 0u;
                                const uint32_t ob = (pass == 2) ? OB_LO : 0u;
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    mma_f16_pair(d, dw + ((oa + k * 32) >> 4), dbk + ((ob + k * 32) >> 4), IDESC3, (kb | pass | k) ? 1u : 0u);
                            }
                            mma_commit_pair(BAR(W_EMPTY + stage), (uint16_t)0x3);
                            if (kb == 1) mma_commit_pair(BAR(TM_FULL + s), (uint16_t)0x3);
                        }
                        __syncwarp();
                        if (++stage == 3) { stage = 0; wphase ^= 1u; }
                    }
                    ++use;
                }
            }
        }
    } else if (warp < 18) {
        // ===================== epilogue (16 warps): max / arg-max of my 128 channels of every 256-channel block =====================
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        uint32_t fe[2] = {0u, 0u};
        long long use = 0;
        for (int t = T0; t < T1; ++t) {
            const int b = t / p.tiles_per_cloud, tt = t % p.tiles_per_cloud;
            const int n0 = tt * L3_NT;
            const int nvalid = (p.N - n0 < L3_NT) ? p.N - n0 : L3_NT;
            ++use;                                                   // the layer-2 use of this tile is not mine
            for (int mt4 = 0; mt4 < 4; ++mt4, ++use) {
                const int s = (int)(use & 1);
                const int ch = (((mt4 + pair) & 3) * 2 + (int)rank) * 128 + row;
                mbar_wait(BAR(TM_FULL + s), fe[s]); fe[s] ^= 1u;
                tc_fence_after_sync();
                float best = -INFINITY; int bidx = 0;
                bool released = false;
                const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * L3_NT);
                for (int c0 = half * (L3_NT / 4); c0 < (half + 1) * (L3_NT / 4); c0 += 32) {
                    if (c0 >= nvalid) break;                // warp-uniform
                    float v[32];
                    tmem_ld32(tbase + (uint32_t)c0, v);
                    if (c0 + 32 >= (half + 1) * (L3_NT / 4) || c0 + 32 >= nvalid) {
                        // my last chunk is in registers: hand the accumulator back before the arithmetic on it (tc_l3.cuh)
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) { if (leader) mbar_arrive(BAR(TM_EMPTY + s)); else mbar_arrive_cluster(BAR(TM_EMPTY + s), 0u); }
                        released = true;
                    }
                    if (c0 + 32 <= nvalid) {
                        float m0 = v[0], m1 = v[1], m2 = v[2], m3 = v[3];
#pragma unroll
                        for (int j = 4; j < 32; j += 4) {
                            m0 = fmaxf(m0, v[j]); m1 = fmaxf(m1, v[j + 1]); m2 = fmaxf(m2, v[j + 2]); m3 = fmaxf(m3, v[j + 3]);
                        }
                        const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                        if (m > best) {
                            best = m;
                            int jj = 31;
#pragma unroll
                            for (int j = 30; j >= 0; --j) if (v[j] == m) jj = j;
                            bidx = n0 + c0 + jj;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c0 + j < nvalid && v[j] > best) { best = v[j]; bidx = n0 + c0 + j; }
                    }
                }
                if (!released) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) {                        // one arrival per warp on the LEADER's barrier
                        if (leader) mbar_arrive(BAR(TM_EMPTY + s)); else mbar_arrive_cluster(BAR(TM_EMPTY + s), 0u);
                    }
                }
                const unsigned long long key = ((unsigned long long)ord_encode(best) << 32) |
                                               (unsigned long long)(0xFFFFFFFFu - (unsigned)bidx);
                atomicMax(&p.keys[(size_t)b * C3 + ch], key);
            }
        }
    } else {
        // ===================== producers (8 warps): layer 1 on the CUDA cores, BatchNorm2 + ReLU of the layer-2 accumulator =====================
        const int wp = warp - 18;                           // 0..7
        const int tidp = wp * 32 + lane;                    // 0..255
        const int pt = tidp & 127, chh = tidp >> 7;         // layer 1: point row, channels chh*32 .. +32
        const int q = warp & 3, hh = wp >> 2;               // drain: TMEM quadrant of this warp, k-block (64 channels) hh
        const int r2 = q * 32 + lane;                       // drain: point row
        unsigned char* a1s = smem + FZ_A1;
        auto stage_a1 = [&](int t) {
            const int b = t / p.tiles_per_cloud, tt = t % p.tiles_per_cloud;
            const int n0 = tt * L3_NT + (int)rank * L3C_NH;
            const bool valid = n0 + pt < p.N;
            float t0 = 0.f, t1 = 0.f, t2 = 0.f;
            if (valid) {
                const float* xb = p.x + (size_t)b * 3 * p.N + n0 + pt;
                const float p0 = xb[0], p1 = xb[p.N], p2 = xb[2 * (size_t)p.N];
                t0 = p0; t1 = p1; t2 = p2;
                if (p.trans) {
                    const float* T = p.trans + (size_t)b * 9;
                    t0 = T[0] * p0 + T[3] * p1 + T[6] * p2;
                    t1 = T[1] * p0 + T[4] * p1 + T[7] * p2;
                    t2 = T[2] * p0 + T[5] * p1 + T[8] * p2;
                }
            }
            bool oor = false;
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
                const int c0 = chh * 32 + g8 * 8;
                float a[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = c0 + j;
                    const float u = s_w1[c * 3 + 0] * t0 + s_w1[c * 3 + 1] * t1 + s_w1[c * 3 + 2] * t2;
                    float v = relu_nan(s_sc1[c] * u + s_sh1[c]);
                    oor = oor || !(v <= TC_ACT_LIMIT);
                    a[j] = valid ? v * ACT_SCALE : 0.f;
                }
                __half2 h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) split2(a[2 * j], a[2 * j + 1], h[j], l[j]);
                const uint32_t off = (uint32_t)(pt * 128 + (((c0 >> 3) ^ (pt & 7)) << 4));
                uint4 hv, lv;
                hv.x = *reinterpret_cast<uint32_t*>(&h[0]); hv.y = *reinterpret_cast<uint32_t*>(&h[1]);
                hv.z = *reinterpret_cast<uint32_t*>(&h[2]); hv.w = *reinterpret_cast<uint32_t*>(&h[3]);
                lv.x = *reinterpret_cast<uint32_t*>(&l[0]); lv.y = *reinterpret_cast<uint32_t*>(&l[1]);
                lv.z = *reinterpret_cast<uint32_t*>(&l[2]); lv.w = *reinterpret_cast<uint32_t*>(&l[3]);
                *reinterpret_cast<uint4*>(a1s + off) = hv;
                *reinterpret_cast<uint4*>(a1s + 16384 + off) = lv;
            }
            if (valid && oor) p.bad[b] = 1u;
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(BAR(A1_FULL)); else mbar_arrive_cluster(BAR(A1_FULL), 0u);
            }
        };
        if (T0 < T1) stage_a1(T0);
        uint32_t ph2 = 0u;
        long long use = 0;
        for (int t = T0; t < T1; ++t, use += 5) {
            const int b = t / p.tiles_per_cloud, tt = t % p.tiles_per_cloud;
            const int n0 = tt * L3_NT + (int)rank * L3C_NH;
            const bool valid = n0 + r2 < p.N;
            const int s = (int)(use & 1);
            mbar_wait(BAR(TM2_FULL), ph2); ph2 ^= 1u;       // layer 2 of this tile is complete (and with it every earlier MMA:
            tc_fence_after_sync();                          // the a2 tile and the a1 tile are free to be overwritten)
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * L3_NT + hh * 64);
            bool oor = false;
#pragma unroll 1
            for (int h2 = 0; h2 < 2; ++h2) {
                float v[32];
                tmem_ld32(taddr + (uint32_t)(h2 * 32), v);
                if (h2 == 1) {
                    // both halves are in registers / consumed: the accumulator slot may be overwritten
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) {
                        if (leader) mbar_arrive(BAR(TM2_EMPTY + s)); else mbar_arrive_cluster(BAR(TM2_EMPTY + s), 0u);
                    }
                }
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    const int cl = h2 * 32 + g8 * 8;                 // channel within the k-block
                    float a[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = hh * 64 + cl + j;
                        const float z = relu_nan(fmaf(s_s2[c], v[g8 * 8 + j], s_h2[c]));
                        oor = oor || !(z <= 60000.f);
                        a[j] = valid ? z : 0.f;
                    }
                    __half2 hq[4], lq[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) split2(a[2 * j], a[2 * j + 1], hq[j], lq[j]);
                    const uint32_t off = (uint32_t)(r2 * 128 + (((cl >> 3) ^ (r2 & 7)) << 4));
                    uint4 hv, lv;
                    hv.x = *reinterpret_cast<uint32_t*>(&hq[0]); hv.y = *reinterpret_cast<uint32_t*>(&hq[1]);
                    hv.z = *reinterpret_cast<uint32_t*>(&hq[2]); hv.w = *reinterpret_cast<uint32_t*>(&hq[3]);
                    lv.x = *reinterpret_cast<uint32_t*>(&lq[0]); lv.y = *reinterpret_cast<uint32_t*>(&lq[1]);
                    lv.z = *reinterpret_cast<uint32_t*>(&lq[2]); lv.w = *reinterpret_cast<uint32_t*>(&lq[3]);
                    *reinterpret_cast<uint4*>(smem + (0 * 2 + hh) * L3C_A2_PART + off) = hv;
                    *reinterpret_cast<uint4*>(smem + (1 * 2 + hh) * L3C_A2_PART + off) = lv;
                }
            }
            if (valid && oor) p.bad[b] = 1u;
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(BAR(A2_FULL)); else mbar_arrive_cluster(BAR(A2_FULL), 0u);
            }
            if (t + 1 < T1) stage_a1(t + 1);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                 // nobody exits while the peer may still signal this CTA
    if (warp == 1) tmem_dealloc_pair<512>(tmem);
}

inline bool fused_configure() {
    static int done[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!done[dev & 63]) {
        const cudaError_t e = cudaFuncSetAttribute(k_tower_fused_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, FZ_SMEM_BYTES);
        if (e != cudaSuccess) cudaGetLastError();
        done[dev & 63] = (e == cudaSuccess) ? 1 : -1;
    }
    return done[dev & 63] == 1;
}

}}  // namespace pgpd::tc
