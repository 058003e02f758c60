// tower.cuh -- the per-point shared-MLP tower 3->64->128->1024 + global max-pool, forward and
// backward, CUDA-core (fp32) kernels + host orchestration.
//
// Reference semantics: STN3d.forward lines 29-33 (relu_last) and PointNetfeat.forward lines
// 140-149 (transform, no ReLU before the pool) of PointNetGPD/model/pointnet.py.
//
// Design (DESIGN.md has the derivations):
//  * the 1024-wide activation is NEVER written to memory.  BatchNorm3(+ReLU) is monotone per
//    channel, so max_n post(u3) = post(max_n sgn*u3): the GEMM epilogue keeps a per-(cloud,
//    channel) running (max, first arg-max) of the raw pre-activation and BN3 is applied to the
//    pooled [B,1024] values only;
//  * train-mode batch statistics are exact two-pass: the mean of every pre-activation is linear in
//    the mean of its input (mean(W a) = W mean(a)), so it is known BEFORE the GEMM runs and the
//    epilogue accumulates centred squares;
//  * the 64- and 128-wide activations (a1, u2) are kept in HBM (768 B/point) for the backward;
//  * the layer-3 backward never forms the dense 1024-wide gradient: with the Gram matrix of a2 the
//    dense BatchNorm terms collapse to 128x128 algebra plus a sparse scatter of the <=1024
//    arg-max rows per cloud ("layer-3 backward collapse").
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"
#include "tails.cuh"
#include "l2bwd.cuh"
#ifndef PGPD_EMU
#include "tc_l3.cuh"
#include "tc_kb.cuh"
#include "tc_ka.cuh"
#include "tc_kf.cuh"
#include "tc_fused.cuh"
#endif

namespace pgpd {

// ================================================================================================
// workspace
// ================================================================================================
struct TowerKeep {
    // ---- kept from forward to backward -------------------------------------------------------
    float* A1;        // [M][64]   a1 = relu(bn1(conv1(x')))
    float* Y2;        // [M][128]  u2 = W2 a1 (bias-free pre-activation of layer 2)
    float* uext;      // [B][1024] raw u3 at the arg-max (sign restored)
    int* idx;         // [B][1024] arg-max point index within the cloud
    BnState bn[3];
    float* sgn;       // [1024] +1 / -1 : sign of gamma3 (max vs min selection)
    double* S1;       // [128]  sum over points of a2
    double* S1a;      // [64]   sum over points of a1
    double* xmom;     // [B][12] per-cloud raw-coordinate moments: sum x (3), sum x x^T (3x3)
};

struct TowerScratch {
    double* moments;  // [B][9]
    double* dpart;    // double partials, max(nb_a1*64, nb_a2*128)
    double* rtmp;     // [REDUCE_MAX_SLICES][16384] stage-1 output of the two-stage reductions
    float* fpart;     // float partials: max over users (see plan_tower_scratch)
    unsigned long long* keys;  // [B][1024]   -- keys, bad and counters are contiguous: ONE memset of zero_bytes per forward
    unsigned* bad;             // [B+1] per-cloud "NaN / out-of-range activation" flags (+ [B]: a conv3 weight is not finite)
    unsigned* counters;        // [16] tickets of the fused tail kernels (self-resetting)
    size_t zero_bytes;
    void* wimg;       // 512 KB: pre-swizzled hi/lo fp16 image of W3 for the tcgen05 kernel
    float* mu_s;      // [1024] mean of u3 in accumulator units (tcgen05 kernel)
    float* centre2;   // [128] pilot estimate of mean(u2) the tcgen05 layer-2 kernel centres its squares on (train)
    float* s1part;    // [256 * 8][128] partial sums of a2 * 2^4 written by the tcgen05 layer-3 kernel
    void* wimg_kb;    // 48 KB: the two A-operand images of the fused layer-2/1 backward pass (tc_kb.cuh)
    void* wimg_s;     // 64 KB: image of the resident weight matrix of a streaming tcgen05 GEMM
    float* inv_s;     // [128] its per-row inverse scales
    float* pmax;      // [1024][2][128] per-epilogue-row maxima of |dz2|, |yhat2|
    float* pmx;       // [128] per-channel max |dz2| (reduced from pmax)
    float* esc;       // [128] per-channel power-of-two scale of dy2 (tcgen05 dW2)
    float* einv;      // [128] its inverse
    // backward scratch
    float* coef;      // [B][1024]
    float* dvec;      // [1024]
    float* evec;      // [1024]
    float* gram;      // [128*128]
    float* gram2;     // [2][128*128] reduced hi.hi / hi.lo Gram accumulators of the fused pass-1 kernel
    float* ka_part;   // [1024][2][128] BatchNorm2-backward partial sums of the tcgen05 pass-1 kernels
    float* WG;        // [1024*128]
    float* Q;         // [128*128]
    float* uvec;      // [128]
    float* da2s;      // [B*1024][128] compact rows of the sparse part of d a2
    int* slot;        // [M]  row in da2s or -1
    float* DZ2;       // [M][128]
    float* m1_2; float* m2_2;   // [128]
    float* m1_1; float* m2_1;   // [64]
    // fused layer-2/1 backward (l2bwd.cuh, tc_kb.cuh)
    float* Kmat;      // [64*64]
    float* cvec;      // [64]
    float* kbC;       // [128*64]  C = sum dz2 a1^T
    float* kbG1;      // [64*64]   Gram of a1
    float* kb_Cpart;  // [KB_MAX_PART][128*64]
    float* kb_G1part; // [KB_MAX_PART][64*64]
    float* kb_bn;     // [kb_rows][2][64]
    float* kb_H;      // [kb_rows][64*3]
    int kb_rows;
    // sizes
    int nb_a1, nb_a2, nb_l2, nb_gram, tiles_per_cloud;
    size_t fpart_elems;
};

struct TowerWs : TowerKeep, TowerScratch {};

constexpr int GRAM_CHUNK = 1024;   // points per split-K block of the Gram GEMM
constexpr int KB_MAX_PART = KB_REF_MAX_BLOCKS;   // per-block partial slots of the fused layer-2/1 backward pass

inline void plan_tower(Carver& c, TowerKeep& w, int B, int N) {
    const size_t M = (size_t)B * N;
    w.A1 = c.take<float>(M * C1);
    w.Y2 = c.take<float>(M * C2);
    w.uext = c.take<float>((size_t)B * C3);
    w.idx = c.take<int>((size_t)B * C3);
    w.bn[0].carve(c, C1); w.bn[1].carve(c, C2); w.bn[2].carve(c, C3);
    w.sgn = c.take<float>(C3);
    w.S1 = c.take<double>(C2);
    w.S1a = c.take<double>(C1);
    w.xmom = c.take<double>((size_t)B * 12);
}

inline void plan_tower_scratch(Carver& c, TowerScratch& w, int B, int N, bool backward) {
    const size_t M = (size_t)B * N;
    w.tiles_per_cloud = idiv_up(N, 128);
    w.nb_a1 = B * idiv_up(N, A1_CHUNK * A1_CPB);                              // one partial row per k_a1 block
    w.nb_a2 = (int)std::min<size_t>(8192, (M + 15) / 16);
    w.nb_l2 = (int)((M + 127) / 128);
    w.nb_gram = (int)((M + GRAM_CHUNK - 1) / GRAM_CHUNK);
    w.moments = c.take<double>((size_t)B * 12);
    w.rtmp = c.take<double>((size_t)TL2_BLOCKS * C2 + 256);      // scratch rows of the fused tails (k_tail_l2 partials, BatchNorm-backward sums)
    w.dpart = c.take<double>((size_t)std::max(w.nb_a1 * C1, w.nb_a2 * C2));
    // persistent tcgen05 kernels write one partial row per CTA (<= TC_MAX_CTAS rows)
    constexpr size_t TC_MAX_CTAS = 256;
    size_t fp = (size_t)w.nb_l2 * C2;                                        // css2 partials
    fp = std::max(fp, 2 * TC_MAX_CTAS * 2 * C2);
    fp = std::max(fp, (size_t)B * w.tiles_per_cloud * C3);                   // css3 partials (CUDA-core kernel, 128-point tiles)
    fp = std::max(fp, (size_t)B * 4 * idiv_up(N, 256) * C3);                  // css3 partials (tcgen05 kernels: up to 4 rows per 256-point tile)
    if (backward) {
        fp = std::max(fp, (size_t)w.nb_gram * C2 * C2);                      // Gram partials
        fp = std::max(fp, (size_t)w.nb_l2 * 2 * C2);                         // BN2 backward partials
        fp = std::max(fp, (size_t)B * (C1 * 3));                             // dW1 partials
        fp = std::max(fp, 2 * TC_MAX_CTAS * C2 * C2);                        // per-CTA Gram (hi.hi, hi.lo) / BN-backward partials
    }
    w.fpart_elems = fp;
    w.fpart = c.take<float>(fp);
    {
        // [B*1024 u64 keys][B+1 flags][16 tickets], zeroed together at the start of every forward
        const size_t key_bytes = (size_t)B * C3 * sizeof(unsigned long long);
        const size_t flag_elems = (size_t)B + 1 + 16;
        unsigned char* z = c.take<unsigned char>(key_bytes + flag_elems * sizeof(unsigned));
        w.keys = reinterpret_cast<unsigned long long*>(z);
        w.bad = reinterpret_cast<unsigned*>(z + key_bytes);
        w.counters = w.bad + (B + 1);
        w.zero_bytes = key_bytes + flag_elems * sizeof(unsigned);
    }
    w.wimg = c.take<unsigned char>((size_t)512 * 1024);
    w.mu_s = c.take<float>(C3);
    w.centre2 = c.take<float>(C2);
    w.s1part = c.take<float>((size_t)256 * 8 * C2);
    w.wimg_kb = c.take<unsigned char>((size_t)48 * 1024);
    w.wimg_s = c.take<unsigned char>((size_t)64 * 1024);
    w.inv_s = c.take<float>(C2);
    w.pmax = c.take<float>((size_t)1024 * 2 * C2);
    w.pmx = c.take<float>(C2);
    w.esc = c.take<float>(C2);
    w.einv = c.take<float>(C2);
    if (backward) {
        w.coef = c.take<float>((size_t)B * C3);
        w.dvec = c.take<float>(C3);
        w.evec = c.take<float>(C3);
        w.gram = c.take<float>(C2 * C2);
        w.gram2 = c.take<float>(2 * C2 * C2);
        w.ka_part = c.take<float>((size_t)std::max(1024, w.nb_l2) * 2 * C2);
        w.WG = c.take<float>((size_t)C3 * C2);
        w.Q = c.take<float>(C2 * C2);
        w.uvec = c.take<float>(C2);
        w.da2s = c.take<float>((size_t)B * C3 * C2);
        w.slot = c.take<int>(M);
        w.DZ2 = c.take<float>(M * C2);
        w.m1_2 = c.take<float>(C2); w.m2_2 = c.take<float>(C2);
        w.m1_1 = c.take<float>(C1); w.m2_1 = c.take<float>(C1);
        w.Kmat = c.take<float>(C1 * C1);
        w.cvec = c.take<float>(C1);
        w.kbC = c.take<float>(C2 * C1);
        w.kbG1 = c.take<float>(C1 * C1);
        w.kb_Cpart = c.take<float>((size_t)KB_MAX_PART * C2 * C1);
        w.kb_G1part = c.take<float>((size_t)KB_MAX_PART * C1 * C1);
        w.kb_rows = B * idiv_up(N, 64) * 4;             // >= the rows either version of the pass writes
        w.kb_bn = c.take<float>((size_t)w.kb_rows * 2 * C1);
        w.kb_H = c.take<float>((size_t)w.kb_rows * C1 * 3);
    }
}

// ================================================================================================
// forward kernels
// ================================================================================================

// per-block sums over points of a2 = relu(scale2*u2 + shift2); block = 128 channels x 2 slots.
// pstride > 1: only every pstride-th point (Ms = ceil(M / pstride) samples) -- the pilot estimate of mean(a2).
__global__ void k_a2_sum(const float* __restrict__ Y2, size_t Ms, size_t pstride, BnState st, double* __restrict__ part) {
    __shared__ double sh[256];
    const int tid = (int)threadIdx.x, k = tid & 127, q = tid >> 7;
    const float sc = st.scale[k], sf = st.shift[k];
    // four independent partial sums so that the loads of consecutive iterations overlap
    float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
    const size_t stride = (size_t)gridDim.x * 2;
    const size_t rs = pstride * C2;                 // floats between consecutive samples
    size_t P = (size_t)blockIdx.x * 2 + q;
    for (; P + 3 * stride < Ms; P += 4 * stride) {
        const float y0 = Y2[P * rs + k], y1 = Y2[(P + stride) * rs + k], y2 = Y2[(P + 2 * stride) * rs + k], y3 = Y2[(P + 3 * stride) * rs + k];
        f0 += fmaxf(sc * y0 + sf, 0.f); f1 += fmaxf(sc * y1 + sf, 0.f); f2 += fmaxf(sc * y2 + sf, 0.f); f3 += fmaxf(sc * y3 + sf, 0.f);
    }
    for (; P < Ms; P += stride) f0 += fmaxf(sc * Y2[P * rs + k] + sf, 0.f);
    const double acc = ((double)f0 + (double)f1) + ((double)f2 + (double)f3);
    sh[tid] = acc;
    __syncthreads();
    if (tid < 128) part[(size_t)blockIdx.x * C2 + tid] = sh[tid] + sh[tid + 128];
}

// ---- layer 2 forward: u2[P][c] = sum_k a1[P][k] W2[c][k] -----------------------------------------
struct ProbL2Fwd {
    static constexpr bool A_KFAST = true, B_NFAST = false;
    static constexpr int SCRATCH = 0;
    using Cfg = CfgBig;
    const float* A1; const float* W2; float* Y2; const float* mean_u2; float* css_part; size_t M;
    struct Blk { int m0, n0, k0, k1; };
    __device__ void setup(Blk& b) const { b.m0 = (int)blockIdx.x * Cfg::BM; b.n0 = 0; b.k0 = 0; b.k1 = C1; }
    __device__ void prologue(const Blk&, float*) const {}
    __device__ float loadA(const Blk&, const float*, int m, int k) const { return (size_t)m < M ? A1[(size_t)m * C1 + k] : 0.f; }
    __device__ float loadB(const Blk&, const float*, int k, int n) const { return W2[n * C1 + k]; }
    __device__ void epilogue(const Blk& b, const float*, float (&acc)[Cfg::TM][Cfg::TN], int ty, int tx, void* red) const {
        float v[Cfg::TN];
#pragma unroll
        for (int j = 0; j < Cfg::TN; ++j) v[j] = 0.f;
#pragma unroll
        for (int i = 0; i < Cfg::TM; ++i) {
            size_t P = (size_t)b.m0 + Cfg::row_of(ty, i);
            if (P < M) {
#pragma unroll
                for (int c4 = 0; c4 < Cfg::TN / 4; ++c4) {
                    int col = Cfg::col_of(tx, c4 * 4);
                    *reinterpret_cast<float4*>(&Y2[P * C2 + col]) =
                        make_float4(acc[i][c4 * 4 + 0], acc[i][c4 * 4 + 1], acc[i][c4 * 4 + 2], acc[i][c4 * 4 + 3]);
                }
                if (mean_u2) {
#pragma unroll
                    for (int j = 0; j < Cfg::TN; ++j) {
                        float d = acc[i][j] - mean_u2[Cfg::col_of(tx, j)];
                        v[j] = fmaf(d, d, v[j]);
                    }
                }
            }
        }
        if (mean_u2) {
            float s = block_col_sum<Cfg>(v, ty, tx, (float*)red);
            if ((int)threadIdx.x < Cfg::BN) css_part[(size_t)blockIdx.x * C2 + threadIdx.x] = s;
        }
    }
};

// ---- layer 3 forward + max-pool: u3[c][n] = sum_k W3[c][k] a2[n][k], per cloud ------------------
struct ProbL3Fwd {
    static constexpr bool A_KFAST = true, B_NFAST = false;
    static constexpr int SCRATCH = 2 * C2;
    using Cfg = CfgBig;
    const float* W3; const float* Y2; const float* scale2; const float* shift2; const float* sgn;
    const float* mean_u3; unsigned long long* keys; float* css_part; int N; int tiles;
    struct Blk { int m0, n0, k0, k1, b; };
    __device__ void setup(Blk& b) const {
        b.m0 = (int)blockIdx.y * Cfg::BM; b.n0 = (int)blockIdx.x * Cfg::BN; b.b = (int)blockIdx.z; b.k0 = 0; b.k1 = C2;
    }
    __device__ void prologue(const Blk&, float* s) const {
        for (int i = (int)threadIdx.x; i < C2; i += Cfg::NT) { s[i] = scale2[i]; s[C2 + i] = shift2[i]; }
        __syncthreads();
    }
    __device__ float loadA(const Blk&, const float*, int m, int k) const { return W3[(size_t)m * C2 + k]; }
    __device__ float loadB(const Blk& b, const float* s, int k, int n) const {
        if (n >= N) return 0.f;
        return relu_nan(s[k] * Y2[((size_t)b.b * N + n) * C2 + k] + s[C2 + k]);
    }
    __device__ void epilogue(const Blk& b, const float*, float (&acc)[Cfg::TM][Cfg::TN], int ty, int tx, void* red) const {
        unsigned long long best[Cfg::TM];
        float v[Cfg::TM];
#pragma unroll
        for (int i = 0; i < Cfg::TM; ++i) {
            const int c = b.m0 + Cfg::row_of(ty, i);
            const float sg = sgn[c];
            const float mu = mean_u3 ? mean_u3[c] : 0.f;
            unsigned long long bk = 0ull;
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < Cfg::TN; ++j) {
                const int n = b.n0 + Cfg::col_of(tx, j);
                if (n < N) {
                    unsigned long long key = ((unsigned long long)ord_encode(sg * acc[i][j]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)n);
                    bk = key > bk ? key : bk;
                    float d = acc[i][j] - mu;
                    s = fmaf(d, d, s);
                }
            }
            best[i] = bk;
            v[i] = s;
        }
        unsigned long long km = block_row_max_u64<Cfg>(best, ty, tx, (unsigned long long*)red);
        if ((int)threadIdx.x < Cfg::BM) atomicMax(&keys[(size_t)b.b * C3 + b.m0 + threadIdx.x], km);
        if (mean_u3) {
            float s = block_row_sum<Cfg>(v, ty, tx, (float*)red);
            if ((int)threadIdx.x < Cfg::BM)
                css_part[((size_t)b.b * tiles + blockIdx.x) * C3 + b.m0 + threadIdx.x] = s;
        }
    }
};

// ================================================================================================
// backward kernels
// ================================================================================================

// BatchNorm3 / max-pool backward on the pooled values.  block = 32 channels x 32 cloud lanes.
__global__ void k_pool_bwd(const float* __restrict__ dG, const float* __restrict__ uext, int B, int relu_last,
                           double count, const float* __restrict__ gamma, BnState st,
                           float* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta,
                           float* __restrict__ dvec, float* __restrict__ evec) {
    __shared__ double sh1[32][33], sh2[32][33];
    const int tid = (int)threadIdx.x, cx = tid & 31, ry = tid >> 5;
    const int c = (int)blockIdx.x * 32 + cx;
    const float sc = st.scale[c], sf = st.shift[c], mu = st.mean[c], r = st.rstd[c];
    double sdz = 0.0, sdzy = 0.0;
#pragma unroll 4
    for (int b = ry; b < B; b += 32) {
        const size_t i = (size_t)b * C3 + c;
        const float u = uext[i];
        float dz = dG[i];
        if (relu_last && !(sc * u + sf > 0.f)) dz = 0.f;
        const float yhat = (u - mu) * r;
        sdz += (double)dz;
        sdzy += (double)dz * (double)yhat;
        coef[i] = sc * dz;
    }
    sh1[ry][cx] = sdz; sh2[ry][cx] = sdzy;
    __syncthreads();
    if (ry == 0) {
        double t1 = 0.0, t2 = 0.0;
#pragma unroll
        for (int q = 0; q < 32; ++q) { t1 += sh1[q][cx]; t2 += sh2[q][cx]; }
        dgamma[c] = (float)t2;
        dbeta[c] = (float)t1;
        const double m1 = t1 / count, m2 = t2 / count;
        const double d = (double)sc * m2 * (double)r;
        dvec[c] = (float)d;
        evec[c] = (float)((double)sc * m1 - d * (double)mu);
    }
    (void)gamma;
}

// Gram = sum_P a2[P] a2[P]^T, split over point chunks
struct ProbGram {
    static constexpr bool A_KFAST = false, B_NFAST = true;
    static constexpr int SCRATCH = 2 * C2;
    using Cfg = CfgBig;
    const float* Y2; const float* scale2; const float* shift2; float* part; size_t M;
    struct Blk { int m0, n0, k0, k1; };
    __device__ void setup(Blk& b) const {
        b.m0 = 0; b.n0 = 0;
        size_t k0 = (size_t)blockIdx.x * GRAM_CHUNK, k1 = k0 + GRAM_CHUNK;
        b.k0 = (int)k0; b.k1 = (int)(k1 < M ? k1 : M);
    }
    __device__ void prologue(const Blk&, float* s) const {
        for (int i = (int)threadIdx.x; i < C2; i += Cfg::NT) { s[i] = scale2[i]; s[C2 + i] = shift2[i]; }
        __syncthreads();
    }
    __device__ float a2(const float* s, int P, int ch) const { return fmaxf(s[ch] * Y2[(size_t)P * C2 + ch] + s[C2 + ch], 0.f); }
    __device__ float loadA(const Blk&, const float* s, int m, int k) const { return a2(s, k, m); }
    __device__ float loadB(const Blk&, const float* s, int k, int n) const { return a2(s, k, n); }
    __device__ void epilogue(const Blk&, const float*, float (&acc)[Cfg::TM][Cfg::TN], int ty, int tx, void*) const {
        float* out = part + (size_t)blockIdx.x * C2 * C2;
#pragma unroll
        for (int i = 0; i < Cfg::TM; ++i)
#pragma unroll
            for (int j = 0; j < Cfg::TN; ++j) out[Cfg::row_of(ty, i) * C2 + Cfg::col_of(tx, j)] = acc[i][j];
    }
};

// dW3[c][k] = sum_b coef[b][c] a2[argmax(b,c)][k]  -  d[c] * (W3 Gram)[c][k]  -  e[c] * S1[k]
// grid = 1024 channels (+ extra blocks, see below), block = 512 = 16 cloud lanes (warps) x 32 lanes of 4 consecutive k (one 16-byte load per lane and
// arg-max row, four rows in flight per warp); the 16 cloud lanes are summed in a fixed order.  The row c of W3 * Gram
// (128 x 128, L2-resident) is formed here by warp 0 instead of by a separate 1024 x 128 x 128 GEMM launch.
// Blocks [1024, gridDim.x) are a different job that merely shares the launch: rows of the precompute of the fused layer-2/1
// backward pass (tails.cuh: kb_prep_row), which like dW3 only depends on the tail of pass A.
__global__ void __launch_bounds__(512) k_dw3(const float* __restrict__ coef, const int* __restrict__ idx, const float* __restrict__ Y2, BnState st2,
                      int B, int N, const float* __restrict__ dvec, const float* __restrict__ evec, const float* __restrict__ W3,
                      const float* __restrict__ gram, const double* __restrict__ S1, float* __restrict__ dW3, float* __restrict__ db3,
                      KbPrepParams kp) {
    if ((int)blockIdx.x >= C3) { kb_prep_row(kp, (int)blockIdx.x - C3); return; }
    __shared__ float4 sh[16][32];
    __shared__ float s_cf[512];
    __shared__ int s_ix[512];
    __shared__ float s_w[C2];
    const int c = (int)blockIdx.x, tid = (int)threadIdx.x, lane = tid & 31, q = tid >> 5;
    const float4 sc = *reinterpret_cast<const float4*>(st2.scale + 4 * lane);
    const float4 sf = *reinterpret_cast<const float4*>(st2.shift + 4 * lane);
    if (tid < C2) s_w[tid] = W3[(size_t)c * C2 + tid];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b0 = 0; b0 < B; b0 += 512) {
        const int nb = (B - b0 < 512) ? B - b0 : 512;
        // stage this channel's coefficients / arg-max indices first, so the row loads below are independent
        if (tid < nb) { s_cf[tid] = coef[(size_t)(b0 + tid) * C3 + c]; s_ix[tid] = idx[(size_t)(b0 + tid) * C3 + c]; }
        __syncthreads();
#pragma unroll 8
        for (int bb = q; bb < nb; bb += 16) {
            const float cf = s_cf[bb];
            const size_t P = (size_t)(b0 + bb) * N + s_ix[bb];
            const float4 y = *reinterpret_cast<const float4*>(Y2 + P * C2 + 4 * lane);
            acc.x = fmaf(cf, fmaxf(fmaf(sc.x, y.x, sf.x), 0.f), acc.x);
            acc.y = fmaf(cf, fmaxf(fmaf(sc.y, y.y, sf.y), 0.f), acc.y);
            acc.z = fmaf(cf, fmaxf(fmaf(sc.z, y.z, sf.z), 0.f), acc.z);
            acc.w = fmaf(cf, fmaxf(fmaf(sc.w, y.w, sf.w), 0.f), acc.w);
        }
        __syncthreads();
    }
    // (W3 Gram)[c][4*lane .. 4*lane+3]: every warp takes 8 of the 128 terms (all loads in flight at once; one warp doing all 128 in
    // sequence was a 7 us serial tail per block: ncu, profiles/r2) and folds -d * its partial into its gather sum
    {
        float4 wg = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int jj = 0; jj < C2 / 16; ++jj) {
            const int j = q * (C2 / 16) + jj;
            const float wj = s_w[j];
            const float4 gr = *reinterpret_cast<const float4*>(gram + (size_t)j * C2 + 4 * lane);
            wg.x = fmaf(wj, gr.x, wg.x); wg.y = fmaf(wj, gr.y, wg.y); wg.z = fmaf(wj, gr.z, wg.z); wg.w = fmaf(wj, gr.w, wg.w);
        }
        const float d = dvec[c];
        acc.x = fmaf(-d, wg.x, acc.x); acc.y = fmaf(-d, wg.y, acc.y); acc.z = fmaf(-d, wg.z, acc.z); acc.w = fmaf(-d, wg.w, acc.w);
    }
    sh[q][lane] = acc;
    __syncthreads();
    if (q == 0) {
        float4 t = sh[0][lane];
#pragma unroll
        for (int w = 1; w < 16; ++w) { const float4 u = sh[w][lane]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
        const float e = evec[c];
        float4 o;
        o.x = t.x - e * (float)S1[4 * lane + 0];
        o.y = t.y - e * (float)S1[4 * lane + 1];
        o.z = t.z - e * (float)S1[4 * lane + 2];
        o.w = t.w - e * (float)S1[4 * lane + 3];
        *reinterpret_cast<float4*>(dW3 + (size_t)c * C2 + 4 * lane) = o;
        if (lane == 0 && db3) db3[c] = 0.f;   // bias feeding a train-mode BatchNorm: gradient is identically zero
    }
}

// sparse part of d a2: rows  sum_{c : argmax(b,c)=p} coef[b][c] W3[c][:]  for the arg-max points of one cloud.
// block = 1024 threads (one per channel) = one cloud.  The (point, channel) pairs are sorted, then the 1024 entries are processed
// ENTRY-parallel: warp w takes the sorted entries 32w .. 32w+31 (lane = 4 consecutive output channels), loads the W3 rows and
// coefficients of eight entries at a time (all in flight together) and accumulates runs of equal points in sorted order.  A row that
// starts in warp w and continues into later warps is finished by warp w, which adds the later warps' "head" partials from shared
// memory in warp order -- every row is a fixed-order sum, the result is reproducible.  (Round 1 walked the ROWS, one warp per row,
// one dependent L2 round trip per two entries: 79 us per launch with `long scoreboard` as the top stall; ncu, profiles/r2.)
__global__ void __launch_bounds__(1024, 1) k_da2_sparse(const float* __restrict__ coef, const int* __restrict__ idx, const float* __restrict__ W3, int N,
                             float* __restrict__ da2s, int* __restrict__ slot) {
    __shared__ unsigned sk[C3];
    __shared__ int scan[2][C3];
    __shared__ int rowpos[C3 + 1];
    __shared__ float4 headp[32][32];
    __shared__ int nvalid_s;
    const int b = (int)blockIdx.x, tid = (int)threadIdx.x;
    // bitonic sort of 1024 keys, one per thread IN A REGISTER: partners less than 32 apart are exchanged by warp shuffle, only the
    // 15 steps with stride >= 32 go through shared memory (two alternating buffers: one barrier per step instead of the 55 barriers
    // of the all-shared-memory version)
    unsigned key;
    {
        const float cf = coef[(size_t)b * C3 + tid];
        key = (cf != 0.f) ? (((unsigned)idx[(size_t)b * C3 + tid] << 10) | (unsigned)tid) : 0xFFFFFFFFu;
    }
    {
        unsigned* xbuf[2] = {sk, reinterpret_cast<unsigned*>(scan[0])};
        int xb = 0;
        for (int size = 2; size <= C3; size <<= 1) {
            const bool up = ((tid & size) == 0);
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                unsigned other;
                if (stride >= 32) {
                    xbuf[xb][tid] = key;
                    __syncthreads();
                    other = xbuf[xb][tid ^ stride];
                    xb ^= 1;
                } else {
                    other = __shfl_xor_sync(0xffffffffu, key, stride);
                }
                const bool lower = (tid & stride) == 0;         // I hold the lower index of the pair
                const unsigned mn = key < other ? key : other, mx = key < other ? other : key;
                key = (lower == up) ? mn : mx;
            }
        }
        __syncthreads();                                        // the last readers of the exchange buffers are done
        sk[tid] = key;
        __syncthreads();
    }
    // row starts + inclusive scan: entry e belongs to row incl[e] - 1
    const bool valid = key != 0xFFFFFFFFu;
    const bool start = valid && (tid == 0 || (sk[tid - 1] >> 10) != (key >> 10));
    scan[0][tid] = start ? 1 : 0;
    __syncthreads();
    int cb = 0;
    for (int off = 1; off < C3; off <<= 1) {
        int v = scan[cb][tid];
        if (tid >= off) v += scan[cb][tid - off];
        scan[cb ^ 1][tid] = v;
        cb ^= 1;
        __syncthreads();
    }
    const int* incl = scan[cb];
    if (start) rowpos[incl[tid] - 1] = tid;
    if (tid == 0) {
        int lo = 0, hi = C3;                        // first invalid key (keys are sorted, invalid = max)
        while (lo < hi) { int mid = (lo + hi) >> 1; if (sk[mid] != 0xFFFFFFFFu) lo = mid + 1; else hi = mid; }
        nvalid_s = lo;
    }
    __syncthreads();
    const int nvalid = nvalid_s;
    const int wrp = tid >> 5, lane = tid & 31;
    const int wbeg = wrp * 32, wend = (wbeg + 32 < nvalid) ? wbeg + 32 : nvalid;
    const float4* W3v = reinterpret_cast<const float4*>(W3);
    float4* out = reinterpret_cast<float4*>(da2s);
    auto store_row = [&](int r, const float4& a) {
        const size_t row = (size_t)b * C3 + r;
        out[row * (C2 / 4) + lane] = a;
        if (lane == 0) slot[(size_t)b * N + (sk[rowpos[r]] >> 10)] = (int)row;
    };
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int cur = -1;
    bool own = false;
    if (wbeg < wend) {
        cur = incl[wbeg] - 1;
        const bool head_open = rowpos[cur] < wbeg;          // my first row began in an earlier warp: its part here is a head partial
        bool first_row = true;
        constexpr int U = 8;
        for (int e0 = wbeg; e0 < wend; e0 += U) {
            float4 wv[U]; float cf[U]; int rw[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u;
                if (e < wend) {
                    const int c = (int)(sk[e] & 1023u);
                    cf[u] = coef[(size_t)b * C3 + c];
                    wv[u] = W3v[(size_t)c * (C2 / 4) + lane];
                    rw[u] = incl[e] - 1;
                } else { cf[u] = 0.f; wv[u] = make_float4(0.f, 0.f, 0.f, 0.f); rw[u] = -1; }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (rw[u] < 0) continue;
                if (rw[u] != cur) {                             // the run of `cur` ends inside my segment
                    if (first_row && head_open) headp[wrp][lane] = acc; else store_row(cur, acc);
                    first_row = false;
                    acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    cur = rw[u];
                }
                acc.x = fmaf(cf[u], wv[u].x, acc.x); acc.y = fmaf(cf[u], wv[u].y, acc.y);
                acc.z = fmaf(cf[u], wv[u].z, acc.z); acc.w = fmaf(cf[u], wv[u].w, acc.w);
            }
        }
        if (first_row && head_open) headp[wrp][lane] = acc;     // my whole segment lies inside a row that began earlier
        else own = true;                                        // my last row began in my segment: I finish it
    }
    __syncthreads();
    if (own) {
        for (int w2 = wrp + 1; w2 < 32 && w2 * 32 < nvalid && incl[w2 * 32] - 1 == cur; ++w2) {
            const float4 h = headp[w2][lane];
            acc.x += h.x; acc.y += h.y; acc.z += h.z; acc.w += h.w;
        }
        store_row(cur, acc);
    }
}

// Q = W3^T diag(d) W3, uvec = W3^T e (+ the operand image of Q): tails.cuh: q_uvec_block
__global__ void __launch_bounds__(1024) k_q_uvec(QuParams p) { q_uvec_block(p); }

// ---- layer 2 backward, pass 1: d a2 -> dz2 (stored) + BN2 backward sums ---------------------------
struct ProbL2BwdA {
    static constexpr bool A_KFAST = true, B_NFAST = true;
    static constexpr int SCRATCH = 2 * C2;
    using Cfg = CfgBig;
    const float* Y2; BnState st2; const float* Q; const float* uvec; const float* da2s; const int* slot;
    float* DZ2; float* part; size_t M;
    struct Blk { int m0, n0, k0, k1; };
    __device__ void setup(Blk& b) const { b.m0 = (int)blockIdx.x * Cfg::BM; b.n0 = 0; b.k0 = 0; b.k1 = C2; }
    __device__ void prologue(const Blk&, float* s) const {
        for (int i = (int)threadIdx.x; i < C2; i += Cfg::NT) { s[i] = st2.scale[i]; s[C2 + i] = st2.shift[i]; }
        __syncthreads();
    }
    __device__ float loadA(const Blk&, const float* s, int m, int k) const {
        return (size_t)m < M ? fmaxf(s[k] * Y2[(size_t)m * C2 + k] + s[C2 + k], 0.f) : 0.f;
    }
    __device__ float loadB(const Blk&, const float*, int k, int n) const { return Q[k * C2 + n]; }
    __device__ void epilogue(const Blk& b, const float* s, float (&acc)[Cfg::TM][Cfg::TN], int ty, int tx, void* red) const {
        float v1[Cfg::TN], v2[Cfg::TN];
#pragma unroll
        for (int j = 0; j < Cfg::TN; ++j) { v1[j] = 0.f; v2[j] = 0.f; }
#pragma unroll
        for (int i = 0; i < Cfg::TM; ++i) {
            const size_t P = (size_t)b.m0 + Cfg::row_of(ty, i);
            if (P >= M) continue;
            const int sl = slot[P];
#pragma unroll
            for (int j = 0; j < Cfg::TN; ++j) {
                const int c = Cfg::col_of(tx, j);
                const float y = Y2[P * C2 + c];
                const float z = s[c] * y + s[C2 + c];
                float da2 = -acc[i][j] - uvec[c];
                if (sl >= 0) da2 += da2s[(size_t)sl * C2 + c];
                const float dz = z > 0.f ? da2 : 0.f;
                DZ2[P * C2 + c] = dz;
                const float yhat = (y - st2.mean[c]) * st2.rstd[c];
                v1[j] += dz;
                v2[j] = fmaf(dz, yhat, v2[j]);
            }
        }
        float s1 = block_col_sum<Cfg>(v1, ty, tx, (float*)red);
        float s2 = block_col_sum<Cfg>(v2, ty, tx, (float*)red);
        if ((int)threadIdx.x < Cfg::BN) {
            part[((size_t)blockIdx.x * 2 + 0) * C2 + threadIdx.x] = s1;
            part[((size_t)blockIdx.x * 2 + 1) * C2 + threadIdx.x] = s2;
        }
    }
};

// ================================================================================================
// host orchestration
// ================================================================================================
struct TowerArgs {
    const pgpd_tower* t;
    const float* x;
    const float* trans;   // or nullptr
    int B, N;
    bool relu_last, train, save;
    cudaStream_t stream;
    bool use_tc;          // tcgen05 kernels where available (never in the emulator build)
};

inline void tower_forward(const TowerArgs& a, TowerWs& w, float* pooled) {
    const pgpd_tower& t = *a.t;
    const size_t M = (size_t)a.B * a.N;
    cudaStream_t s = a.stream;
    const double count = (double)M;
    bool tcp = false;               // tensor-core path (tcgen05 kernels) for this call
#ifndef PGPD_EMU
    tcp = a.use_tc;
#endif
    const float act_limit = tcp ? TC_ACT_LIMIT : INFINITY;

    // max-pool keys, per-cloud "bad activation" flags and the tickets of the fused tail kernels: one memset
    cudaMemsetAsync(w.keys, 0, w.zero_bytes, s);

    // ---- F1: cloud moments + BatchNorm1 statistics (train) / folded running statistics (eval) + weight images -------------
    {
        PreParams p{};
        p.x = a.x; p.trans = a.trans; p.B = a.B; p.N = a.N;
        p.moments = w.moments; p.rawmom = a.save ? w.xmom : nullptr;
        for (int L = 0; L < 3; ++L) { p.conv[L] = t.conv[L]; p.bn[L] = t.bn[L]; p.st[L] = w.bn[L]; }
        p.train = a.train ? 1 : 0; p.count = count;
        p.counter = w.counters + 0; p.bad = w.bad;
        p.n_mom = a.train ? a.B : 1;
        p.n_w2 = tcp ? C2 / 4 : 0; p.n_w3 = tcp ? C3 / 2 : 0;
        p.wimg2 = w.wimg_s; p.inv2 = w.inv_s; p.wimg3 = w.wimg; p.sgn = w.sgn;
#ifndef PGPD_EMU
        p.act_shift = tc::ACT_SHIFT;
        p.centre2 = (tcp && a.train) ? w.centre2 : nullptr;
#endif
        const int n_sign = tcp ? 0 : C3 / PRE_THREADS;
        launch(k_tower_pre, dim3(p.n_mom + p.n_w2 + p.n_w3 + n_sign), dim3(PRE_THREADS), 0, s, p);
    }

#ifndef PGPD_EMU
    // ---- eval mode on the tensor-core path: the whole tower in ONE kernel (tc_fused.cuh); a1 / u2 / a2 never leave the SM --------
    if (tcp && !a.train && tc::fused_configure()) {
        const int tpc = idiv_up(a.N, tc::L3_NT), ntiles = a.B * tpc;
        tc::FusedParams p{a.x, a.trans, t.conv[0].w, w.bn[0].scale, w.bn[0].shift, (const __half*)w.wimg_s, w.inv_s,
                          w.bn[1].scale, w.bn[1].shift, (const __half*)w.wimg, w.sgn, w.keys, a.B, a.N, tpc, ntiles, w.bad};
        const int sms = tc::dev_info().sms;
        const int pairs = ntiles < sms / 2 ? ntiles : sms / 2;
        CUtensorMap wmap;
        tc::make_image_map(&wmap, w.wimg, tc::L3_WIMG_BYTES);      // host-side encoding of the weight image's tensor map (no driver call that allocates or synchronises)
        profiler().begin(s);
        launch(tc::k_tower_fused_eval, dim3(2 * pairs), dim3(tc::FZ_THREADS), (size_t)tc::FZ_SMEM_BYTES, s, p, wmap);
        profiler().end(s);
        TailL3Params q{};
        q.B = a.B; q.relu_last = a.relu_last ? 1 : 0; q.train = 0;
        q.keys = w.keys; q.sgn = w.sgn; q.st3 = w.bn[2];
        q.pooled = pooled; q.uext = nullptr; q.idx = nullptr;
        q.bad = w.bad; q.limit = act_limit;
        launch(k_tail_l3, dim3(C3 / TL3_CH), dim3(1024), 0, s, q);
        return;
    }
#endif

    // ---- layers 1 + 2 ---------------------------------------------------------------------------------------------------------
    // Tensor-core path: ONE kernel; a1 is formed from the coordinates inside the layer-2 kernel's operand producers and never
    // written (the backward recomputes it the same way, tc_kb.cuh).  CUDA-core path: k_a1 writes a1, a tiled GEMM reads it.
    int n_css2 = 0, n_s1a = 0;
#ifndef PGPD_EMU
    if (tcp) {
        const int tpc = idiv_up(a.N, tc::KF_NT), ntiles = a.B * tpc;
        tc::KfParams p{(const __half*)w.wimg_s, w.inv_s, a.train ? (const float*)w.centre2 : (const float*)nullptr,
                       a.x, a.trans, t.conv[0].w, w.bn[0].scale, w.bn[0].shift, w.Y2, w.fpart,
                       a.train ? w.pmax : (float*)nullptr, w.bad, a.B, a.N, tpc, ntiles};
        const int grid = tc::launch_kf(p, tc::dev_info().sms, s);
        n_css2 = tc::KF_EPI_ROWS * grid;
        n_s1a = grid;
    } else
#endif
    {
        A1Params q{a.x, a.trans, a.B, a.N, t.conv[0].w, w.bn[0], w.A1, a.train ? w.dpart : (double*)nullptr,
                   w.counters + 1, t.conv[1].w, count, w.bn[1].mean, w.S1a, w.bad, act_limit};
        launch(k_a1, dim3(idiv_up(a.N, A1_CHUNK * A1_CPB), a.B), dim3(A1_THREADS), 0, s, q);
        ProbL2Fwd p{w.A1, t.conv[1].w, w.Y2, a.train ? w.bn[1].mean : nullptr, w.fpart, M};
        launch_gemm<ProbL2Fwd::Cfg>(p, dim3(w.nb_l2), s);
        n_css2 = w.nb_l2;
    }

    // ---- T3 (train): BatchNorm2 statistics, (pilot) mean of a2 -> mean of the layer-3 pre-activation -----------------------
    // Tensor-core path with many points: only a PILOT estimate of mean(a2) from TL2_SAMPLE points is computed here (the
    // centre of the layer-3 kernel's sum of squares); that kernel's operand producers accumulate the exact sum of a2 on the
    // way and k_tail_l3 corrects the statistics (var = sum (u-c)^2 / M - (mean - c)^2, an identity).
    bool l3_pilot = false;
    if (a.train) {
        TailL2Params p{};
        p.css = w.fpart; p.n_css = n_css2;
        p.s1a_part = tcp ? w.pmax : nullptr; p.n_s1a = n_s1a; p.S1a = w.S1a; p.s1a_scale = 1.0 / 16.0; p.W2 = t.conv[1].w;
        p.mean_u2 = tcp ? w.centre2 : w.bn[1].mean; p.count = count; p.bias2 = t.conv[1].b; p.bn2 = t.bn[1]; p.st2 = w.bn[1];
        p.Y2 = w.Y2; p.part = w.rtmp; p.counter = w.counters + 2;
        p.W3 = t.conv[2].w; p.mean_u3 = w.bn[2].mean; p.S1 = w.S1; p.inv3 = w.sgn; p.mu_s = tcp ? w.mu_s : nullptr;
        const size_t total_samples = (size_t)TL2_BLOCKS * TL2_SPB;
        p.npoints = M; p.pstride = M > total_samples ? M / total_samples : 1; p.nsample = M;
        launch(k_tail_l2, dim3(TL2_BLOCKS), dim3(1024), 0, s, p);
        if (tcp) {
            l3_pilot = true;        // the centres are estimates: the layer-3 kernel accumulates the exact sum of a2, k_tail_l3 corrects
        } else {
            // CUDA-core path: its layer-3 kernel centres the squares on the exact mean, so the sum of a2 over ALL points is
            // taken now (BatchNorm2 is final) and the tail runs once more for the mean propagation only
            const int nb = (int)std::min<size_t>((size_t)w.nb_a2, (M + 15) / 16);
            launch(k_a2_sum, dim3(nb), dim3(256), 0, s, (const float*)w.Y2, M, (size_t)1, w.bn[1], w.dpart);
            p.bn_done = 1; p.a2part = w.dpart; p.n_a2part = nb; p.nsample = M; p.pstride = 1;
            launch(k_tail_l2, dim3(TL2_BLOCKS), dim3(1024), 0, s, p);
        }
    }

    // ---- layer 3 + max-pool ------------------------------------------------------------------------------------------------
    int n_css = 0, n_s1 = 0;
#ifndef PGPD_EMU
    if (tcp) {
        const int tpc = idiv_up(a.N, tc::L3_NT), ntiles = a.B * tpc;
        tc::L3Params p{w.Y2, w.bn[1].scale, w.bn[1].shift, (const __half*)w.wimg, w.sgn, a.train ? w.mu_s : nullptr,
                       w.keys, w.fpart, a.B, a.N, tpc, ntiles, tc::l3_debug_buffer_if_enabled(),
                       l3_pilot ? w.s1part : (float*)nullptr, w.bad};
        const int sms = tc::dev_info().sms;
        const int pairs = ntiles < sms / 2 ? ntiles : sms / 2;
        CUtensorMap wmap;
        tc::make_image_map(&wmap, w.wimg, tc::L3_WIMG_BYTES);      // host-side encoding of the weight image's tensor map (no driver call that allocates or synchronises)
        profiler().begin(s);
        launch(tc::k_l3_fwd_tc3, dim3(2 * pairs), dim3(tc::L3C_THREADS), (size_t)tc::L3C_SMEM_BYTES, s, p, wmap);
        profiler().end(s);
        n_css = pairs * tc::L3C_EPI_ROWS;             // partial rows of centred squares: four per CTA pair
        n_s1 = 2 * pairs;                             // partial rows of the sum of a2: one per CTA
    } else
#endif
    {
        ProbL3Fwd p{t.conv[2].w, w.Y2, w.bn[1].scale, w.bn[1].shift, w.sgn, a.train ? w.bn[2].mean : nullptr,
                    w.keys, w.fpart, a.N, w.tiles_per_cloud};
        profiler().begin(s);
        launch_gemm<ProbL3Fwd::Cfg>(p, dim3(w.tiles_per_cloud, C3 / 128, a.B), s);
        profiler().end(s);
        n_css = a.B * w.tiles_per_cloud;
    }

    // ---- T4: BatchNorm3 statistics (train) + pooled values -------------------------------------------------------------------
    {
        TailL3Params p{};
        p.B = a.B; p.relu_last = a.relu_last ? 1 : 0; p.train = a.train ? 1 : 0;
        p.keys = w.keys; p.sgn = w.sgn; p.st3 = w.bn[2];
        p.pooled = pooled; p.uext = w.uext; p.idx = w.idx;
        p.bad = w.bad; p.limit = act_limit;
        p.css = w.fpart; p.n_css = n_css;
        p.s1part = l3_pilot ? w.s1part : nullptr; p.n_s1 = n_s1; p.s1scale = 1.0 / 16.0;   // the kernel sums a2 * 2^4 (L3_ACT_SCALE)
        p.W3 = t.conv[2].w; p.bias3 = t.conv[2].b; p.bn3 = t.bn[2]; p.count = count; p.S1 = w.S1;
        launch(k_tail_l3, dim3(C3 / TL3_CH), dim3(1024), 0, s, p);
    }
}

// dpooled [B][1024] -> parameter gradients (+ d trans).  Train-mode statistics only.
inline void tower_backward(const TowerArgs& a, TowerWs& w, const pgpd_tower_grad& g, const float* dpooled, float* dtrans_out) {
    const pgpd_tower& t = *a.t;
    const size_t M = (size_t)a.B * a.N;
    cudaStream_t s = a.stream;
    const double count = (double)M;
    bool tcp = false;
#ifndef PGPD_EMU
    tcp = a.use_tc;
#endif

    // ---- BN3 / max-pool on the pooled values ---------------------------------------------------------
    launch(k_pool_bwd, dim3(C3 / 32), dim3(1024), 0, s, dpooled, (const float*)w.uext, a.B, a.relu_last ? 1 : 0, count,
           t.bn[2].gamma, w.bn[2], w.coef, g.bn[2].dgamma, g.bn[2].dbeta, w.dvec, w.evec);

    // ---- Q = W3^T diag(d) W3, uvec = W3^T e (+ the operand image of Q for the tensor-core pass A) ---------------
    {
        QuParams p{t.conv[2].w, w.dvec, w.evec, w.Q, w.uvec, tcp ? w.wimg_s : nullptr, w.inv_s, 0};
#ifndef PGPD_EMU
        p.act_shift = tc::ACT_SHIFT;
#endif
        launch(k_q_uvec, dim3(C2 / 4), dim3(1024), 0, s, p);
    }

    // ---- sparse part of d a2 -----------------------------------------------------------------------
    cudaMemsetAsync(w.slot, 0xFF, M * sizeof(int), s);
    launch(k_da2_sparse, dim3(a.B), dim3(C3), 0, s, (const float*)w.coef, (const int*)w.idx, t.conv[2].w, a.N, w.da2s, w.slot);

    // ---- pass A: d a2 -> dz2 (stored), BatchNorm2 backward sums, Gram matrix of a2 --------------------------------------------
    TailKaParams tk{};
    tk.g2 = w.gram2; tk.gram = w.gram; tk.count = count; tk.bnsum = w.rtmp; tk.pmx = w.pmx;
    tk.dgamma = g.bn[1].dgamma; tk.dbeta = g.bn[1].dbeta; tk.m1 = w.m1_2; tk.m2 = w.m2_2; tk.counter = w.counters + 3;
    KbPrepParams kp{};
    kp.W2 = t.conv[1].w; kp.st2 = w.bn[1]; kp.m1 = w.m1_2; kp.m2 = w.m2_2; kp.Kmat = w.Kmat; kp.cvec = w.cvec;
    int g_b4 = 0;       // rows of pmax written by the tcgen05 pass-A kernel
#ifndef PGPD_EMU
    if (tcp) {
        const int tpc = idiv_up(a.N, tc::KA_NT), ntiles = a.B * tpc;
        tc::KaParams p{(const __half*)w.wimg_s, w.inv_s, w.uvec, w.bn[1].scale, w.bn[1].shift, t.bn[1].gamma, t.bn[1].beta,
                       w.Y2, w.da2s, w.slot, a.B, a.N, tpc, ntiles, w.DZ2, w.ka_part, w.pmax, w.fpart};
        const int grid = tc::launch_ka(p, tc::dev_info().sms, s);
        g_b4 = grid * tc::KA_EPI_ROWS;
        tk.gpart = w.fpart; tk.n_g = grid; tk.gcols = 2 * C2 * C2; tk.sym = 1;      // per-CTA hi.hi / hi.lo accumulators
        tk.bnpart = w.ka_part; tk.n_bn = g_b4;
        tk.pmax = w.pmax; tk.n_pm = g_b4; tk.act_scale = tc::ACT_SCALE;
        kp.pmx = w.pmx; kp.esc = w.esc; kp.einv = w.einv;
        kp.img1 = w.wimg_kb; kp.img2 = (unsigned char*)w.wimg_kb + tc::KB_A1_BYTES; kp.ginv = w.inv_s; kp.act_scale = tc::ACT_SCALE;
    } else
#endif
    {
        // BatchNorm partial rows go to ka_part so that fpart is free for the Gram partials
        ProbL2BwdA p{w.Y2, w.bn[1], w.Q, w.uvec, w.da2s, w.slot, w.DZ2, w.ka_part, M};
        launch_gemm<ProbL2BwdA::Cfg>(p, dim3(w.nb_l2), s);
        ProbGram pg{w.Y2, w.bn[1].scale, w.bn[1].shift, w.fpart, M};
        launch_gemm<ProbGram::Cfg>(pg, dim3(w.nb_gram), s);
        tk.gpart = w.fpart; tk.n_g = w.nb_gram; tk.gcols = C2 * C2; tk.sym = 0;
        tk.bnpart = w.ka_part; tk.n_bn = w.nb_l2;
        tk.act_scale = 1.f;
    }
    launch(k_tail_ka, dim3(tk.gcols / 256 + 8 + (tcp ? 4 : 0)), dim3(1024), 0, s, tk);

    // ---- dW3 ---------------------------------------------------------------------------------------------------------------
    launch(k_dw3, dim3(C3 + C1), dim3(512), 0, s, (const float*)w.coef, (const int*)w.idx, (const float*)w.Y2, w.bn[1], a.B, a.N,
           (const float*)w.dvec, (const float*)w.evec, t.conv[2].w, (const float*)w.gram, (const double*)w.S1, g.conv[2].dw, g.conv[2].db, kp);

    // ---- layers 2 and 1: one fused pass over (dz2, a1) (l2bwd.cuh / tc_kb.cuh) -------------------------------------------------
    int nparts = 0, nrows = 0, rpc = 0;
#ifndef PGPD_EMU
    if (tcp) {
        __half* img1 = (__half*)w.wimg_kb;
        __half* img2 = img1 + tc::KB_A1_BYTES / 2;
        const int tpc = idiv_up(a.N, tc::KB_NT), ntiles = a.B * tpc;
        tc::KbParams p{img1, img2, w.inv_s, w.cvec, t.bn[0].gamma, t.bn[0].beta, w.esc, w.einv, w.DZ2, a.x, a.trans,
                       t.conv[0].w, w.bn[0].scale, w.bn[0].shift, a.B, a.N, tpc, ntiles, w.kb_Cpart, w.kb_G1part, w.kb_bn, w.kb_H};
        nparts = tc::launch_kb(p, tc::dev_info().sms, s);
        nrows = nparts * tc::KB_EPI_GROUPS; rpc = tpc * tc::KB_EPI_GROUPS;
    } else
#endif
    {
        const int tpc = idiv_up(a.N, KB_REF_NT), ntiles = a.B * tpc;
        const int grid = ntiles < KB_REF_MAX_BLOCKS ? ntiles : KB_REF_MAX_BLOCKS;
        KbRefParams p{w.DZ2, w.A1, a.x, t.conv[1].w, w.bn[1].scale, w.Kmat, w.cvec, t.bn[0].gamma, t.bn[0].beta,
                      a.B, a.N, tpc, ntiles, w.kb_Cpart, w.kb_G1part, w.kb_bn, w.kb_H};
        launch(k_kb_ref, dim3(grid), dim3(256), 0, s, p);
        nparts = grid; nrows = ntiles; rpc = tpc;
    }
    {
        TailKbParams p{w.kb_Cpart, w.kb_G1part, nparts, w.kbC, w.kbG1, w.kb_bn, nrows, count, w.rtmp,
                       g.bn[0].dgamma, g.bn[0].dbeta, w.m1_1, w.m2_1, w.counters + 4};
        launch(k_tail_kb, dim3(TKB_BLOCKS), dim3(1024), 0, s, p);
    }
    {
        Dw2Params d{w.kbC, w.kbG1, w.S1a, t.conv[1].w, w.bn[1], w.m1_2, w.m2_2, g.conv[1].dw, g.conv[1].db};
        launch(k_kb_l1, dim3(a.B + C2), dim3(768), 0, s, a.B, (const float*)w.kb_H, rpc, (const double*)w.xmom, a.trans, t.conv[0].w, w.bn[0],
               (const float*)w.m1_1, (const float*)w.m2_1, w.fpart, a.trans ? dtrans_out : (float*)nullptr, w.counters + 5,
               g.conv[0].dw, g.conv[0].db, d);
    }
}

}  // namespace pgpd
