// dual.cuh -- the reference's dual-cloud network: SimpleSTN3d (PointNetGPD/model/pointnet.py:48-85,
// tower 3->64->128->256 + head 256->128->64->9), DualPointNetfeat (:88-120, two T-Nets on the two 3-channel halves of a 6-channel
// cloud, trunk 6->64->128->1024) and DualPointNetCls (:157-174).  No reference script constructs these classes, so this is the
// straightforward fp32 CUDA-core formulation, written once for any layer width: every Conv1d(k=1)/Linear + BatchNorm + ReLU block --
// per point (rows = B*N) or per cloud (rows = B) -- is the SAME three steps
//     u = act(prev) W^T   (gemm_simt.cuh tile GEMM; BatchNorm+ReLU of the previous block applied while loading)
//     exact two-pass batch statistics of u (fixed-order partial sums in double, finalised by the last block)
//     consumer applies scale/shift(+ReLU) on the fly
// and its backward never materialises du: the gradient of the bias-free pre-activation, s(dz - <dz> - yhat <dz yhat>), is formed
// inside the loaders of the dW and d(prev) GEMMs from the stored u, the dense dz (or, behind the max-pool, the B x C routed
// gradients and arg-max indices).  Stored per point: u1, u2, u3 (+ dz1, dz2 in the backward); the big fused tower of tower.cuh /
// tc_*.cuh stays the product path of PointNetCls.
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"

namespace pgpd {
namespace dual {

constexpr int W1 = 64, W2 = 128;                 // widths of the first two tower layers (pointnet.py:51-52,93-94)
constexpr int STN_C3 = 256, STN_H1 = 128, STN_H2 = 64;   // SimpleSTN3d: conv3 / fc1 / fc2 (pointnet.py:53,55-56)
constexpr int RED_MAX_BLOCKS = 512;              // partial rows of a column reduction
constexpr size_t PART_ELEMS = (size_t)4 << 20;   // 16 MB of split-K partials (dW GEMMs)

// ---- what a consumer sees of a stored pre-activation: a(p,k) = [relu](scale[k] * U[p][k] + shift[k]); scale == null: a = U -------
struct Act {
    const float* U; const float* scale; const float* shift; int C; int relu;
    __device__ __forceinline__ float at(size_t p, int k) const {
        float v = U[p * C + k];
        if (scale) { v = fmaf(scale[k], v, shift[k]); if (relu) v = relu_nan(v); }
        return v;
    }
};

// ---- gradient w.r.t. the bias-free pre-activation u of a block, formed on the fly ----------------------------------------------
//   kind 0: plain dense D[p][c] (a Linear without BatchNorm: the fc3 layers)
//   kind 1: dense dz D[p][c]      + BatchNorm correction  du = s (dz - m1 - yhat m2)
//   kind 2: max-pool routed: dz(p,c) = G[b][c] if idx[b][c] == n else 0 (p = b*N + n)  + BatchNorm correction
struct DuSrc {
    int kind; const float* D; const int* idx; int N; const float* U;
    const float* mean; const float* rstd; const float* scale; const float* m1; const float* m2; int C;
    __device__ __forceinline__ float at(size_t p, int c) const {
        if (kind == 0) return D[p * C + c];
        float dz;
        if (kind == 1) dz = D[p * C + c];
        else { const size_t b = p / (size_t)N; const int n = (int)(p - b * N); dz = idx[b * C + c] == n ? D[b * C + c] : 0.f; }
        const float yh = (U[p * C + c] - mean[c]) * rstd[c];
        return scale[c] * (dz - m1[c] - yh * m2[c]);
    }
};

// ---- GEMM problems (gemm_simt.cuh) ----------------------------------------------------------------------------------------------
// forward: out[m][n] = sum_k act(m,k) W[n][k]  (+ bias[n] + identity of the 3x3 T-Net output when `bias` is given: the fc3 layers)
struct ProbFwd {
    static constexpr bool A_KFAST = true, B_NFAST = false;
    static constexpr int SCRATCH = 0;
    using Cfg = CfgSmall;
    Act a; const float* W; const float* bias; int iden; float* out; int M, Cout;
    struct Blk { int m0, n0, k0, k1; };
    __device__ void setup(Blk& b) const { b.m0 = (int)blockIdx.x * Cfg::BM; b.n0 = (int)blockIdx.y * Cfg::BN; b.k0 = 0; b.k1 = a.C; }
    __device__ void prologue(const Blk&, float*) const {}
    __device__ float loadA(const Blk&, const float*, int m, int k) const { return m < M ? a.at((size_t)m, k) : 0.f; }
    __device__ float loadB(const Blk&, const float*, int k, int n) const { return n < Cout ? W[(size_t)n * a.C + k] : 0.f; }
    __device__ void epilogue(const Blk& b, const float*, float (&acc)[Cfg::TM][Cfg::TN], int ty, int tx, void*) const {
#pragma unroll
        for (int i = 0; i < Cfg::TM; ++i) {
            const int m = b.m0 + Cfg::row_of(ty, i);
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < Cfg::TN; ++j) {
                const int n = b.n0 + Cfg::col_of(tx, j);
                if (n >= Cout) continue;
                float v = acc[i][j];
                if (bias) v += bias[n] + ((iden && (n == 0 || n == 4 || n == 8)) ? 1.f : 0.f);
                out[(size_t)m * Cout + n] = v;
            }
        }
    }
};

// weight gradient: part[z][co][ci] = sum_{p in slice z} du(p,co) act(p,ci)
struct ProbDw {
    static constexpr bool A_KFAST = false, B_NFAST = true;
    static constexpr int SCRATCH = 0;
    using Cfg = CfgSmall;
    DuSrc du; Act a; float* part; int M, ksl;
    struct Blk { int m0, n0, k0, k1; };
    __device__ void setup(Blk& b) const {
        b.m0 = (int)blockIdx.y * Cfg::BM; b.n0 = (int)blockIdx.x * Cfg::BN;
        const long long k0 = (long long)blockIdx.z * ksl, k1 = k0 + ksl;
        b.k0 = (int)k0; b.k1 = (int)(k1 < M ? k1 : M);
    }
    __device__ void prologue(const Blk&, float*) const {}
    __device__ float loadA(const Blk&, const float*, int m, int k) const { return m < du.C ? du.at((size_t)k, m) : 0.f; }
    __device__ float loadB(const Blk&, const float*, int k, int n) const { return n < a.C ? a.at((size_t)k, n) : 0.f; }
    __device__ void epilogue(const Blk& b, const float*, float (&acc)[Cfg::TM][Cfg::TN], int ty, int tx, void*) const {
        float* o = part + (size_t)blockIdx.z * du.C * a.C;
#pragma unroll
        for (int i = 0; i < Cfg::TM; ++i) {
            const int m = b.m0 + Cfg::row_of(ty, i);
            if (m >= du.C) continue;
#pragma unroll
            for (int j = 0; j < Cfg::TN; ++j) {
                const int n = b.n0 + Cfg::col_of(tx, j);
                if (n < a.C) o[(size_t)m * a.C + n] = acc[i][j];
            }
        }
    }
};

// input gradient: dprev[p][ci] = (sum_co du(p,co) W[co][ci]) masked by [z_prev(p,ci) > 0] when the previous block has a ReLU
struct ProbDx {
    static constexpr bool A_KFAST = true, B_NFAST = true;
    static constexpr int SCRATCH = 0;
    using Cfg = CfgSmall;
    DuSrc du; const float* W; Act prev; float* out; int M;      // prev.C = Cin; prev.scale == null: no mask (the tower input)
    struct Blk { int m0, n0, k0, k1; };
    __device__ void setup(Blk& b) const { b.m0 = (int)blockIdx.x * Cfg::BM; b.n0 = (int)blockIdx.y * Cfg::BN; b.k0 = 0; b.k1 = du.C; }
    __device__ void prologue(const Blk&, float*) const {}
    __device__ float loadA(const Blk&, const float*, int m, int k) const { return m < M ? du.at((size_t)m, k) : 0.f; }
    __device__ float loadB(const Blk&, const float*, int k, int n) const { return n < prev.C ? W[(size_t)k * prev.C + n] : 0.f; }
    __device__ void epilogue(const Blk& b, const float*, float (&acc)[Cfg::TM][Cfg::TN], int ty, int tx, void*) const {
#pragma unroll
        for (int i = 0; i < Cfg::TM; ++i) {
            const int m = b.m0 + Cfg::row_of(ty, i);
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < Cfg::TN; ++j) {
                const int n = b.n0 + Cfg::col_of(tx, j);
                if (n >= prev.C) continue;
                float v = acc[i][j];
                if (prev.scale && prev.relu) {
                    const float z = fmaf(prev.scale[n], prev.U[(size_t)m * prev.C + n], prev.shift[n]);
                    v = z > 0.f ? v : 0.f;
                }
                out[(size_t)m * prev.C + n] = v;
            }
        }
    }
};

// ---- column reductions over the rows of an [M][C] matrix, deterministic -----------------------------------------------------------
// Every block sums a contiguous range of rows (lanes in a fixed order), writes one partial row in double; the last block to arrive
// adds the partial rows in block order and finalises.
//   RED_MEAN : mean[c] = sum X / M
//   RED_VAR  : var[c] = sum (X - mean)^2 / M  -> BatchNorm finalisation (folded affine + running statistics)
//   RED_BNBWD: s1 = sum DZ, s2 = sum DZ * yhat(U)  -> dbeta, dgamma, m1 = s1 / M, m2 = s2 / M
//   RED_SUM  : o1[c] = sum X                   (bias gradient of a Linear without BatchNorm)
enum { RED_MEAN = 0, RED_VAR = 1, RED_BNBWD = 2, RED_SUM = 3 };
struct RedParams {
    int mode; const float* X; const float* U; int M, C; BnState st; pgpd_bn bn; const float* bias;
    double* part; unsigned* counter; float* o1; float* o2; float* m1; float* m2;
};

__global__ void __launch_bounds__(256) k_colred(RedParams p) {
    __shared__ double sh1[256], sh2[256];
    const int tid = (int)threadIdx.x, nblk = (int)gridDim.x;
    const int cw = p.C < 256 ? p.C : 256, lanes = 256 / cw;
    const int lane = tid / cw, c0 = tid - lane * cw;
    const bool live = lane < lanes;
    const int chunk = idiv_up(p.M, nblk);
    const int r0 = (int)blockIdx.x * chunk, r1 = r0 + chunk < p.M ? r0 + chunk : p.M;
    const int two = p.mode == RED_BNBWD ? 2 : 1;
    for (int cb = 0; cb < p.C; cb += cw) {
        const int c = cb + c0;
        double a1 = 0.0, a2 = 0.0;
        if (live && c < p.C) {
            const float mu = (p.mode == RED_VAR || p.mode == RED_BNBWD) ? p.st.mean[c] : 0.f;
            const float rs = p.mode == RED_BNBWD ? p.st.rstd[c] : 0.f;
            for (int r = r0 + lane; r < r1; r += lanes) {
                const float x = p.X[(size_t)r * p.C + c];
                if (p.mode == RED_VAR) { const float d = x - mu; a1 += (double)d * (double)d; }
                else if (p.mode == RED_BNBWD) { a1 += (double)x; a2 += (double)x * (double)((p.U[(size_t)r * p.C + c] - mu) * rs); }
                else a1 += (double)x;
            }
        }
        __syncthreads();
        sh1[tid] = a1; sh2[tid] = a2;
        __syncthreads();
        if (lane == 0 && c < p.C) {
            double t1 = 0.0, t2 = 0.0;
            for (int l = 0; l < lanes; ++l) { t1 += sh1[l * cw + c0]; t2 += sh2[l * cw + c0]; }
            p.part[((size_t)blockIdx.x * two + 0) * p.C + c] = t1;
            if (two == 2) p.part[((size_t)blockIdx.x * two + 1) * p.C + c] = t2;
        }
    }
    if (!last_block_done(p.counter, (unsigned)nblk)) return;
    const double count = (double)p.M;
    for (int c = tid; c < p.C; c += 256) {
        double t1 = 0.0, t2 = 0.0;
        for (int b = 0; b < nblk; ++b) {
            t1 += p.part[((size_t)b * two + 0) * p.C + c];
            if (two == 2) t2 += p.part[((size_t)b * two + 1) * p.C + c];
        }
        if (p.mode == RED_MEAN) p.st.mean[c] = (float)(t1 / count);
        else if (p.mode == RED_VAR) bn_finalize_train(c, (double)p.st.mean[c], t1 / count, count, p.bias, p.bn, p.st);
        else if (p.mode == RED_BNBWD) {
            p.o1[c] = (float)t2;            // dgamma
            p.o2[c] = (float)t1;            // dbeta
            p.m1[c] = (float)(t1 / count);
            p.m2[c] = (float)(t2 / count);
        } else p.o1[c] = (float)t1;
    }
}

// eval mode: fold the running statistics and the conv / fc bias into scale / shift
__global__ void k_bn_fold_eval(int C, const float* __restrict__ bias, pgpd_bn bn, BnState st) {
    const int c = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (c >= C) return;
    const float sc = bn.gamma[c] / sqrtf(bn.running_var[c] + BN_EPS);
    st.scale[c] = sc;
    st.shift[c] = bn.beta[c] + sc * ((bias ? bias[c] : 0.f) - bn.running_mean[c]);
}

// ---- tower input: xp[p][h*3 + i] = sum_j x[b][c0 + 3h + j][n] T_h[b][j][i]   (torch.bmm(x^T, trans), pointnet.py:107-109);
// T_h == null: the coordinates themselves ------------------------------------------------------------------------------------------
__global__ void k_xprep(const float* __restrict__ x, int Cx, int c0, int halves, const float* __restrict__ T0, const float* __restrict__ T1,
                        int B, int N, float* __restrict__ xp) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)B * N) return;
    const size_t b = p / (size_t)N; const int n = (int)(p - b * N);
    for (int h = 0; h < halves; ++h) {
        const float* T = h == 0 ? T0 : T1;
        float v[3];
        for (int j = 0; j < 3; ++j) v[j] = x[(b * Cx + c0 + 3 * h + j) * N + n];
        for (int i = 0; i < 3; ++i) {
            float o = v[i];
            if (T) { const float* t = T + b * 9; o = fmaf(v[2], t[6 + i], fmaf(v[1], t[3 + i], v[0] * t[i])); }
            xp[p * (3 * halves) + 3 * h + i] = o;
        }
    }
}

// d trans_h[b][j][i] = sum_n x[b][3h + j][n] dxp[b*N + n][3h + i]  (+ the caller's gradient w.r.t. the returned trans1 + trans2)
__global__ void __launch_bounds__(256) k_dtrans(const float* __restrict__ x, const float* __restrict__ dxp, int N, const float* __restrict__ dtrans_user,
                                                float* __restrict__ dO1, float* __restrict__ dO2) {
    __shared__ float sh[256];
    const int b = (int)blockIdx.x, tid = (int)threadIdx.x;
    float acc[18];
    for (int q = 0; q < 18; ++q) acc[q] = 0.f;
    for (int n = tid; n < N; n += 256) {
        const float* d = dxp + ((size_t)b * N + n) * 6;
        for (int h = 0; h < 2; ++h)
            for (int j = 0; j < 3; ++j) {
                const float xv = x[((size_t)b * 6 + 3 * h + j) * N + n];
                for (int i = 0; i < 3; ++i) acc[h * 9 + j * 3 + i] = fmaf(xv, d[3 * h + i], acc[h * 9 + j * 3 + i]);
            }
    }
    for (int q = 0; q < 18; ++q) {
        __syncthreads();
        sh[tid] = acc[q];
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int l = 0; l < 256; ++l) t += sh[l];
            const int ji = q % 9;
            const float u = dtrans_user ? dtrans_user[(size_t)b * 9 + ji] : 0.f;
            (q < 9 ? dO1 : dO2)[(size_t)b * 9 + ji] = t + u;
        }
    }
}

// ---- global max-pool over the N points of a cloud (MaxPool1d(num_points), pointnet.py:73,114), first arg-max, NaN wins -----------
__global__ void __launch_bounds__(256) k_pool_fwd(Act a, int N, float* __restrict__ pooled, int* __restrict__ idx) {
    __shared__ unsigned long long sh[256];
    const int b = (int)blockIdx.x, tid = (int)threadIdx.x, cl = tid & 63, lane = tid >> 6;
    const int c = (int)blockIdx.y * 64 + cl;
    unsigned long long best = 0ull;
    if (c < a.C)
        for (int n = lane; n < N; n += 4) {
            const float v = a.at((size_t)b * N + n, c);
            const unsigned long long key = ((unsigned long long)ord_encode(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)n);
            best = key > best ? key : best;
        }
    sh[tid] = best;
    __syncthreads();
    if (lane == 0 && c < a.C) {
        for (int l = 1; l < 4; ++l) { const unsigned long long o = sh[l * 64 + cl]; best = o > best ? o : best; }
        pooled[(size_t)b * a.C + c] = ord_decode((unsigned)(best >> 32));
        if (idx) idx[(size_t)b * a.C + c] = (int)(0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFull));
    }
}

// backward of the pool + BatchNorm sums of the pooled layer: G[b][c] = d pooled masked by the ReLU, dgamma / dbeta, m1, m2.
// One thread per channel, clouds in order (deterministic).
__global__ void k_pool_bwd(const float* __restrict__ dP, const int* __restrict__ idx, Act a, BnState st, int B, int N,
                           float* __restrict__ G, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ m1, float* __restrict__ m2) {
    const int c = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (c >= a.C) return;
    const float sc = st.scale[c], sf = st.shift[c], mu = st.mean[c], rs = st.rstd[c];
    double s1 = 0.0, s2 = 0.0;
    for (int b = 0; b < B; ++b) {
        const size_t i = (size_t)b * a.C + c;
        const float u = a.U[((size_t)b * N + idx[i]) * a.C + c];
        float g = dP[i];
        if (a.relu && !(fmaf(sc, u, sf) > 0.f)) g = 0.f;
        G[i] = g;
        s1 += (double)g;
        s2 += (double)g * (double)((u - mu) * rs);
    }
    const double count = (double)B * (double)N;
    dgamma[c] = (float)s2; dbeta[c] = (float)s1;
    m1[c] = (float)(s1 / count); m2[c] = (float)(s2 / count);
}

// out[i] = sum_z part[z][i]  (split-K partials of a weight gradient; fixed order).  zero != null: n_zero zeros (the bias gradient of a
// conv / fc feeding a train-mode BatchNorm is identically zero).
__global__ void k_sum_slices(const float* __restrict__ part, int nsl, size_t n, float* __restrict__ out, float* __restrict__ zero, int n_zero) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float v = part[i];
        for (int z = 1; z < nsl; ++z) v += part[(size_t)z * n + i];
        out[i] = v;
    }
    if (zero && i < (size_t)n_zero) zero[i] = 0.f;
}

// log_softmax over the k logits of a row (pointnet.py:174)
__global__ void k_logsoftmax(const float* __restrict__ logits, int B, int K, float* __restrict__ keep, float* __restrict__ user) {
    const int b = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (b >= B) return;
    const float* l = logits + (size_t)b * K;
    float mx = l[0];
    for (int j = 1; j < K; ++j) mx = fmaxf(mx, l[j]);
    float s = 0.f;
    for (int j = 0; j < K; ++j) s += expf(l[j] - mx);
    const float lse = mx + logf(s);
    for (int j = 0; j < K; ++j) {
        const float v = (l[j] != l[j] || s != s) ? NAN : l[j] - lse;
        if (keep) keep[(size_t)b * K + j] = v;
        user[(size_t)b * K + j] = v;
    }
}

__global__ void k_logsoftmax_bwd(const float* __restrict__ logp, const float* __restrict__ dlogp, int B, int K, float* __restrict__ dO) {
    const int b = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (b >= B) return;
    float s = 0.f;
    for (int j = 0; j < K; ++j) s += dlogp[(size_t)b * K + j];
    for (int j = 0; j < K; ++j) dO[(size_t)b * K + j] = dlogp[(size_t)b * K + j] - expf(logp[(size_t)b * K + j]) * s;
}

// ================================================================================================================================
// host side
// ================================================================================================================================
struct Block {          // one Conv1d(k=1)/Linear (+ BatchNorm [+ ReLU]) block
    int Cin, Cout;
    float* U;           // [rows][Cout] bias-free pre-activation (fc3: the output itself)
    BnState st;
    float* DZ;          // [rows][Cout] dense dz (backward; not for a pooled block)
    float* m1; float* m2;   // [Cout]
};

struct Shared {         // scratch shared by every block of a call
    double* red;        // [RED_MAX_BLOCKS][2][1024]
    float* part;        // [PART_ELEMS]
    unsigned* counter;  // one ticket (self-resetting; zeroed once per call)
};

struct TowerW {         // 3 per-point blocks + pool
    int Cin, C3; bool relu_last;
    Block L[3];
    float* xp;          // [M][Cin]
    float* dxp;         // [M][Cin] (trunk backward)
    float* pooled;      // [B][C3]
    int* idx;           // [B][C3]
    float* G;           // [B][C3]
    float* dP;          // [B][C3] gradient w.r.t. pooled (from the head)
};

struct HeadW {          // 2 per-cloud blocks + Linear
    int Cin, H1, H2, out;
    Block L[3];
    float* dO;          // [B][out]
};

inline void plan_block(Carver& c, Block& b, int Cin, int Cout, size_t rows, bool bn, bool dense_dz) {
    b.Cin = Cin; b.Cout = Cout;
    b.U = c.take<float>(rows * Cout);
    if (bn) { b.st.carve(c, Cout); b.m1 = c.take<float>(Cout); b.m2 = c.take<float>(Cout); }
    else { b.st = BnState{}; b.m1 = b.m2 = nullptr; }
    b.DZ = dense_dz ? c.take<float>(rows * Cout) : nullptr;
}

inline void plan_tower(Carver& c, TowerW& t, int Cin, int C3, bool relu_last, int B, int N, bool backward, bool need_dx) {
    const size_t M = (size_t)B * N;
    t.Cin = Cin; t.C3 = C3; t.relu_last = relu_last;
    t.xp = c.take<float>(M * Cin);
    plan_block(c, t.L[0], Cin, W1, M, true, backward);
    plan_block(c, t.L[1], W1, W2, M, true, backward);
    plan_block(c, t.L[2], W2, C3, M, true, false);
    t.pooled = c.take<float>((size_t)B * C3);
    t.idx = c.take<int>((size_t)B * C3);
    t.G = t.dP = t.dxp = nullptr;
    if (backward) {
        t.G = c.take<float>((size_t)B * C3);
        t.dP = c.take<float>((size_t)B * C3);
        if (need_dx) t.dxp = c.take<float>(M * Cin);
    }
}

inline void plan_head(Carver& c, HeadW& h, int Cin, int H1_, int H2_, int out, int B, bool backward) {
    h.Cin = Cin; h.H1 = H1_; h.H2 = H2_; h.out = out;
    plan_block(c, h.L[0], Cin, H1_, (size_t)B, true, backward);
    plan_block(c, h.L[1], H1_, H2_, (size_t)B, true, backward);
    plan_block(c, h.L[2], H2_, out, (size_t)B, false, false);
    h.dO = backward ? c.take<float>((size_t)B * out) : nullptr;
}

inline void plan_shared(Carver& c, Shared& s) {
    s.red = c.take<double>((size_t)RED_MAX_BLOCKS * 2 * 1024);
    s.part = c.take<float>(PART_ELEMS);
    s.counter = c.take<unsigned>(4);
}

inline int red_blocks(int M) { const int n = idiv_up(M, 64); return n < RED_MAX_BLOCKS ? n : RED_MAX_BLOCKS; }

inline Act act_of(const Block& b, bool relu) { return Act{b.U, b.st.scale, b.st.shift, b.Cout, relu ? 1 : 0}; }
inline Act act_plain(const float* X, int C) { return Act{X, nullptr, nullptr, C, 0}; }

// u = act(in) W^T, then batch statistics (train) or folded running statistics (eval)
inline void block_forward(const Act& in, int rows, const pgpd_lin& lin, const pgpd_bn& bn, Block& b, bool train, Shared& sh, cudaStream_t s) {
    ProbFwd p{in, lin.w, nullptr, 0, b.U, rows, b.Cout};
    launch_gemm<CfgSmall>(p, dim3(idiv_up(rows, CfgSmall::BM), idiv_up(b.Cout, CfgSmall::BN)), s);
    if (train) {
        RedParams r{};
        r.X = b.U; r.M = rows; r.C = b.Cout; r.st = b.st; r.bn = bn; r.bias = lin.b; r.part = sh.red; r.counter = sh.counter;
        const int nb = red_blocks(rows);
        r.mode = RED_MEAN; launch(k_colred, dim3(nb), dim3(256), 0, s, r);
        r.mode = RED_VAR;  launch(k_colred, dim3(nb), dim3(256), 0, s, r);
    } else {
        launch(k_bn_fold_eval, grid1d(b.Cout, 128), dim3(128), 0, s, b.Cout, lin.b, bn, b.st);
    }
}

// Given du of block b (formed on the fly by `du`): dW (split over the rows), db (zero / column sum), and the masked gradient of the
// previous block's pre-ReLU output (dprev, may be null).
inline void block_backward(const DuSrc& du, const Act& in, int rows, const pgpd_lin& lin, const pgpd_lin_grad& g, bool has_bn,
                           const Act& prev_mask, float* dprev, Shared& sh, cudaStream_t s) {
    const int Cout = du.C, Cin = in.C;
    const size_t nW = (size_t)Cout * Cin;
    int nsl = idiv_up(rows, 128);
    const int cap = (int)std::min<size_t>(PART_ELEMS / nW, 64);
    if (nsl > cap) nsl = cap;
    if (nsl < 1) nsl = 1;
    int ksl = idiv_up(rows, nsl);
    ksl = idiv_up(ksl, CfgSmall::BK) * CfgSmall::BK;
    nsl = idiv_up(rows, ksl);
    ProbDw pw{du, in, sh.part, rows, ksl};
    launch_gemm<CfgSmall>(pw, dim3(idiv_up(Cin, CfgSmall::BN), idiv_up(Cout, CfgSmall::BM), nsl), s);
    launch(k_sum_slices, grid1d(nW, 256), dim3(256), 0, s, (const float*)sh.part, nsl, nW, g.dw, has_bn ? g.db : (float*)nullptr, Cout);
    if (!has_bn) {
        RedParams r{};
        r.mode = RED_SUM; r.X = du.D; r.M = rows; r.C = Cout; r.part = sh.red; r.counter = sh.counter; r.o1 = g.db;
        launch(k_colred, dim3(red_blocks(rows)), dim3(256), 0, s, r);
    }
    if (dprev) {
        ProbDx px{du, lin.w, prev_mask, dprev, rows};
        launch_gemm<CfgSmall>(px, dim3(idiv_up(rows, CfgSmall::BM), idiv_up(Cin, CfgSmall::BN)), s);
    }
}

// BatchNorm-backward sums of a block whose dense dz is in b.DZ
inline void bn_sums(Block& b, int rows, const pgpd_bn_grad& g, Shared& sh, cudaStream_t s) {
    RedParams r{};
    r.mode = RED_BNBWD; r.X = b.DZ; r.U = b.U; r.M = rows; r.C = b.Cout; r.st = b.st; r.part = sh.red; r.counter = sh.counter;
    r.o1 = g.dgamma; r.o2 = g.dbeta; r.m1 = b.m1; r.m2 = b.m2;
    launch(k_colred, dim3(red_blocks(rows)), dim3(256), 0, s, r);
}

inline DuSrc du_dense(const Block& b) { return DuSrc{1, b.DZ, nullptr, 1, b.U, b.st.mean, b.st.rstd, b.st.scale, b.m1, b.m2, b.Cout}; }

// ---- tower: x (channels c0 .. c0 + 3*halves of a [B][Cx][N] tensor, optionally transformed) -> pooled [B][C3] ---------------------
inline void tower_forward(const pgpd_tower& t, TowerW& w, const float* x, int Cx, int c0, const float* T0, const float* T1,
                          int B, int N, bool train, Shared& sh, cudaStream_t s) {
    const int M = B * N;
    launch(k_xprep, grid1d((size_t)M, 256), dim3(256), 0, s, x, Cx, c0, w.Cin / 3, T0, T1, B, N, w.xp);
    block_forward(act_plain(w.xp, w.Cin), M, t.conv[0], t.bn[0], w.L[0], train, sh, s);
    block_forward(act_of(w.L[0], true), M, t.conv[1], t.bn[1], w.L[1], train, sh, s);
    block_forward(act_of(w.L[1], true), M, t.conv[2], t.bn[2], w.L[2], train, sh, s);
    launch(k_pool_fwd, dim3(B, idiv_up(w.C3, 64)), dim3(256), 0, s, act_of(w.L[2], w.relu_last), N, w.pooled, w.idx);
}

// w.dP (gradient w.r.t. pooled) -> parameter gradients (+ w.dxp when planned)
inline void tower_backward(const pgpd_tower& t, const pgpd_tower_grad& g, TowerW& w, int B, int N, Shared& sh, cudaStream_t s) {
    const int M = B * N;
    Block& L0 = w.L[0]; Block& L1 = w.L[1]; Block& L2 = w.L[2];
    launch(k_pool_bwd, grid1d(w.C3, 128), dim3(128), 0, s, (const float*)w.dP, (const int*)w.idx, act_of(L2, w.relu_last), L2.st, B, N,
           w.G, g.bn[2].dgamma, g.bn[2].dbeta, L2.m1, L2.m2);
    DuSrc du3{2, w.G, w.idx, N, L2.U, L2.st.mean, L2.st.rstd, L2.st.scale, L2.m1, L2.m2, w.C3};
    block_backward(du3, act_of(L1, true), M, t.conv[2], g.conv[2], true, act_of(L1, true), L1.DZ, sh, s);
    bn_sums(L1, M, g.bn[1], sh, s);
    block_backward(du_dense(L1), act_of(L0, true), M, t.conv[1], g.conv[1], true, act_of(L0, true), L0.DZ, sh, s);
    bn_sums(L0, M, g.bn[0], sh, s);
    block_backward(du_dense(L0), act_plain(w.xp, w.Cin), M, t.conv[0], g.conv[0], true, act_plain(nullptr, w.Cin), w.dxp, sh, s);
}

// ---- head: X [B][Cin] -> fc1+bn+relu -> fc2+bn+relu -> fc3 (+ bias, + identity for a T-Net); result in w.L[2].U ------------------
inline void head_forward(const pgpd_head& h, HeadW& w, const float* X, int B, bool train, bool is_stn, Shared& sh, cudaStream_t s) {
    block_forward(act_plain(X, w.Cin), B, h.fc[0], h.bn[0], w.L[0], train, sh, s);
    block_forward(act_of(w.L[0], true), B, h.fc[1], h.bn[1], w.L[1], train, sh, s);
    ProbFwd p{act_of(w.L[1], true), h.fc[2].w, h.fc[2].b, is_stn ? 1 : 0, w.L[2].U, B, w.out};
    launch_gemm<CfgSmall>(p, dim3(idiv_up(B, CfgSmall::BM), idiv_up(w.out, CfgSmall::BN)), s);
}

// w.dO (gradient w.r.t. the fc3 output) -> parameter gradients and dX [B][Cin]
inline void head_backward(const pgpd_head& h, const pgpd_head_grad& g, HeadW& w, const float* X, int B, float* dX, Shared& sh, cudaStream_t s) {
    Block& L0 = w.L[0]; Block& L1 = w.L[1];
    DuSrc du3{0, w.dO, nullptr, 1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, w.out};
    block_backward(du3, act_of(L1, true), B, h.fc[2], g.fc[2], false, act_of(L1, true), L1.DZ, sh, s);
    bn_sums(L1, B, g.bn[1], sh, s);
    block_backward(du_dense(L1), act_of(L0, true), B, h.fc[1], g.fc[1], true, act_of(L0, true), L0.DZ, sh, s);
    bn_sums(L0, B, g.bn[0], sh, s);
    block_backward(du_dense(L0), act_plain(X, w.Cin), B, h.fc[0], g.fc[0], true, act_plain(nullptr, w.Cin), dX, sh, s);
}

// ---- the three modules ------------------------------------------------------------------------------------------------------------
struct DualWs {
    TowerW stn_t[2], trunk;
    HeadW stn_h[2], cls;
    Shared sh;
    float* logp;        // [B][k]
    size_t bytes;
};

inline void plan_dual(void* base, int what, int B, int N, int k, bool backward, DualWs& w) {
    Carver c(base);
    plan_shared(c, w.sh);
    const int nstn = what == PGPD_DUAL_STN ? 1 : 2;
    for (int i = 0; i < nstn; ++i) {
        plan_tower(c, w.stn_t[i], 3, STN_C3, true, B, N, backward, false);
        plan_head(c, w.stn_h[i], STN_C3, STN_H1, STN_H2, 9, B, backward);
    }
    w.logp = nullptr;
    if (what != PGPD_DUAL_STN) {
        plan_tower(c, w.trunk, 6, C3, false, B, N, backward, true);
        if (what == PGPD_DUAL_CLS) {
            plan_head(c, w.cls, C3, H1, H2, k, B, backward);
            w.logp = c.take<float>((size_t)B * k);
        }
    }
    w.bytes = (c.off + 255) & ~(size_t)255;
}

__global__ void k_add3(const float* a, const float* b, const float* c, float* out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + (b ? b[i] : 0.f) + (c ? c[i] : 0.f);
}

inline void dual_forward(int what, const pgpd_dual& m, const float* x, int B, int N, int k, bool train, float* out, float* trans,
                         DualWs& w, cudaStream_t s) {
    cudaMemsetAsync(w.sh.counter, 0, 4 * sizeof(unsigned), s);
    const int Cx = what == PGPD_DUAL_STN ? 3 : 6;
    const pgpd_tower* st[2] = {&m.stn1_tower, &m.stn2_tower};
    const pgpd_head* shd[2] = {&m.stn1_head, &m.stn2_head};
    const int nstn = what == PGPD_DUAL_STN ? 1 : 2;
    for (int i = 0; i < nstn; ++i) {
        tower_forward(*st[i], w.stn_t[i], x, Cx, 3 * i, nullptr, nullptr, B, N, train, w.sh, s);
        head_forward(*shd[i], w.stn_h[i], w.stn_t[i].pooled, B, train, true, w.sh, s);
    }
    // returned transform: trans1 (+ trans2) (pointnet.py:85,117)
    launch(k_add3, grid1d((size_t)B * 9, 128), dim3(128), 0, s, (const float*)w.stn_h[0].L[2].U,
           nstn == 2 ? (const float*)w.stn_h[1].L[2].U : (const float*)nullptr, (const float*)nullptr, trans, (size_t)B * 9);
    if (what == PGPD_DUAL_STN) return;
    tower_forward(m.trunk, w.trunk, x, 6, 0, w.stn_h[0].L[2].U, w.stn_h[1].L[2].U, B, N, train, w.sh, s);
    if (what == PGPD_DUAL_FEAT) {
        cudaMemcpyAsync(out, w.trunk.pooled, (size_t)B * C3 * sizeof(float), cudaMemcpyDeviceToDevice, s);
        return;
    }
    head_forward(m.cls_head, w.cls, w.trunk.pooled, B, train, false, w.sh, s);
    launch(k_logsoftmax, grid1d(B, 128), dim3(128), 0, s, (const float*)w.cls.L[2].U, B, k, w.logp, out);
}

inline void dual_backward(int what, const pgpd_dual& m, const pgpd_dual_grad& g, const float* x, int B, int N, int k,
                          const float* dout, const float* dtrans, DualWs& w, cudaStream_t s) {
    cudaMemsetAsync(w.sh.counter, 0, 4 * sizeof(unsigned), s);
    const pgpd_tower* st[2] = {&m.stn1_tower, &m.stn2_tower};
    const pgpd_head* shd[2] = {&m.stn1_head, &m.stn2_head};
    const pgpd_tower_grad* gt[2] = {&g.stn1_tower, &g.stn2_tower};
    const pgpd_head_grad* gh[2] = {&g.stn1_head, &g.stn2_head};
    const int nstn = what == PGPD_DUAL_STN ? 1 : 2;
    if (what != PGPD_DUAL_STN) {
        if (what == PGPD_DUAL_CLS) {
            launch(k_logsoftmax_bwd, grid1d(B, 128), dim3(128), 0, s, (const float*)w.logp, dout, B, k, w.cls.dO);
            head_backward(m.cls_head, g.cls_head, w.cls, w.trunk.pooled, B, w.trunk.dP, w.sh, s);
        } else {
            cudaMemcpyAsync(w.trunk.dP, dout, (size_t)B * C3 * sizeof(float), cudaMemcpyDeviceToDevice, s);
        }
        tower_backward(m.trunk, g.trunk, w.trunk, B, N, w.sh, s);
        launch(k_dtrans, dim3(B), dim3(256), 0, s, x, (const float*)w.trunk.dxp, N, dtrans, w.stn_h[0].dO, w.stn_h[1].dO);
    } else {
        cudaMemcpyAsync(w.stn_h[0].dO, dtrans, (size_t)B * 9 * sizeof(float), cudaMemcpyDeviceToDevice, s);
    }
    for (int i = 0; i < nstn; ++i) {
        head_backward(*shd[i], *gh[i], w.stn_h[i], w.stn_t[i].pooled, B, w.stn_t[i].dP, w.sh, s);
        tower_backward(*st[i], *gt[i], w.stn_t[i], B, N, w.sh, s);
    }
}

}  // namespace dual
}  // namespace pgpd
