// head.cuh -- the small dense heads 1024->512->256->out with BatchNorm over the batch:
//   STN3d regression head     pointnet.py:35-44  (out = 9, + identity)
//   PointNetCls classifier    pointnet.py:191-194 (out = k, log_softmax)
// 2.6 MFLOP per grasp: launch count, not flops, is what matters here.  fc1 / fc2 (forward, dW, dX) are split-K GEMMs
// (tcgen05: tc_gemm.cuh; CUDA-core: gemm_simt.cuh) that write their K-slice partials; the kernel that CONSUMES a GEMM's
// output (BatchNorm forward / backward, gradient finish) sums the slices in a fixed order on the way in, so no separate
// "split-K finish" launch exists.  fc3 (9 / k outputs) is a CUDA-core kernel fused with its bias, the identity of the
// T-Net and log_softmax.
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"
#ifndef PGPD_EMU
#include "tc_gemm.cuh"
#endif

namespace pgpd {

constexpr size_t HEAD_PART_ELEMS = (size_t)4 << 20;   // 16 MB of fp32 split-K partials per buffer (two buffers)
constexpr int HEAD_AMAX_BLOCKS = 64;                  // partial maxima per weight matrix (forward operand scales)
constexpr int HEAD_AMAX_ELEMS = 2 * HEAD_AMAX_BLOCKS + H1 / 8 + H2 / 8;   // [fc1.weight | fc2.weight] x 64, dU1 x 64, dU2 x 32 (one per BatchNorm-backward block)

struct HeadWs {
    float* U1;        // [B][512]  bias-free pre-activation of fc1
    float* U2;        // [B][256]
    float* Hm1;       // [B][512]  relu(bn(U1)), materialised
    float* Hm2;       // [B][256]
    float* out;       // [B][out]  fc3 output (+bias, + identity for the STN head) = logits / t9
    BnState bn[2];
    float* partA;     // split-K partials of the first GEMM of a pair
    float* partB;     // ... of the second
    // backward scratch
    float* DZ1;       // [B][512]  dU1
    float* DZ2;       // [B][256]  dz2, then (in place) dU2
    float* dO;        // [B][out]  gradient w.r.t. fc3 output
    unsigned* amax;   // [HEAD_AMAX_ELEMS] bit patterns of partial max |x| (operand scales of the tcgen05 GEMMs)
};

inline void plan_head(Carver& c, HeadWs& w, int B, int out, bool backward) {
    w.U1 = c.take<float>((size_t)B * H1);
    w.U2 = c.take<float>((size_t)B * H2);
    w.Hm1 = c.take<float>((size_t)B * H1);
    w.Hm2 = c.take<float>((size_t)B * H2);
    w.out = c.take<float>((size_t)B * out);
    w.bn[0].carve(c, H1); w.bn[1].carve(c, H2);
    w.partA = c.take<float>(HEAD_PART_ELEMS);
    w.partB = c.take<float>(HEAD_PART_ELEMS);
    w.amax = c.take<unsigned>(HEAD_AMAX_ELEMS);
    if (backward) {
        w.DZ1 = c.take<float>((size_t)B * H1);
        w.DZ2 = c.take<float>((size_t)B * H2);
        w.dO = c.take<float>((size_t)B * out);
    }
}

// ---- plain fp32 GEMM on row-major matrices, split-K partials -------------------------------------------------------
//   part[z][m][n] = sum_{k in slice z} A(m,k) B(k,n);  A(m,k) = A[m*sam + k*sak], B(k,n) = Bm[k*sbk + n*sbn]
template <bool AK, bool BNF>
struct ProbPlain {
    static constexpr bool A_KFAST = AK, B_NFAST = BNF;
    static constexpr int SCRATCH = 0;
    using Cfg = CfgSmall;
    const float* A; const float* Bm; float* part;
    int Mr, Nc, K; size_t sam, sak, sbk, sbn;
    int ksl;            // K-slice length; blockIdx.z selects the slice
    struct Blk { int m0, n0, k0, k1; };
    __device__ void setup(Blk& b) const {
        b.m0 = (int)blockIdx.y * Cfg::BM; b.n0 = (int)blockIdx.x * Cfg::BN;
        b.k0 = (int)blockIdx.z * ksl; b.k1 = b.k0 + ksl < K ? b.k0 + ksl : K;
    }
    __device__ void prologue(const Blk&, float*) const {}
    __device__ float loadA(const Blk&, const float*, int m, int k) const { return m < Mr ? A[(size_t)m * sam + (size_t)k * sak] : 0.f; }
    __device__ float loadB(const Blk&, const float*, int k, int n) const { return n < Nc ? Bm[(size_t)k * sbk + (size_t)n * sbn] : 0.f; }
    __device__ void epilogue(const Blk& b, const float*, float (&acc)[Cfg::TM][Cfg::TN], int ty, int tx, void*) const {
        float* out = part + (size_t)blockIdx.z * Mr * Nc;
#pragma unroll
        for (int i = 0; i < Cfg::TM; ++i) {
            const int m = b.m0 + Cfg::row_of(ty, i);
            if (m >= Mr) continue;
#pragma unroll
            for (int j = 0; j < Cfg::TN; ++j) {
                const int n = b.n0 + Cfg::col_of(tx, j);
                if (n < Nc) out[(size_t)m * Nc + n] = acc[i][j];
            }
        }
    }
};

// number of K slices.  fixed_nsl > 0: exactly that many (forward GEMMs: the summation order must not depend on the batch
// size, so that eval outputs are bit-identical however the clouds are batched); 0: by occupancy.
inline int pick_slices(int tiles, int Mr, int Nc, int K, int min_slice, int fixed_nsl, int target) {
    int nsl = 1;
    if (fixed_nsl > 0) nsl = fixed_nsl;
    else
        while (tiles * nsl < target && K / (nsl * 2) >= min_slice && (size_t)(nsl * 2) * Mr * Nc <= HEAD_PART_ELEMS) nsl *= 2;
    while (nsl > 1 && (size_t)nsl * Mr * Nc > HEAD_PART_ELEMS) nsl /= 2;
    return nsl;
}

// returns the number of slices written to `part`
template <bool AK, bool BNF>
inline int run_plain(ProbPlain<AK, BNF> p, cudaStream_t s, int fixed_nsl) {
    const int tm = idiv_up(p.Mr, CfgSmall::BM), tn = idiv_up(p.Nc, CfgSmall::BN);
    int nsl = pick_slices(tm * tn, p.Mr, p.Nc, p.K, 64, fixed_nsl, 296);
    p.ksl = idiv_up(p.K, nsl);
    nsl = idiv_up(p.K, p.ksl);
    launch_gemm<CfgSmall>(p, dim3(tn, tm, nsl), s);
    return nsl;
}

#ifndef PGPD_EMU
// C[M][N] = sum_k A(m,k) B(n,k) on the tensor cores; returns the number of K slices written to `part`
inline int run_gemm_tc(tc::GemmOp A, tc::GemmOp Bo, int M, int N, int K, float* part, int nsl_fixed, cudaStream_t s) {
    const int tiles = idiv_up(M, tc::GM_T) * idiv_up(N, tc::GM_T);
    int nsl = pick_slices(tiles, M, N, K, tc::GM_KC, nsl_fixed, 120);
    const int kslice = idiv_up(idiv_up(K, nsl), tc::GM_KC) * tc::GM_KC;
    nsl = idiv_up(K, kslice);                       // no empty slices
    tc::GemmParams p{A, Bo, M, N, K, kslice, part, nullptr};
    tc::launch_gemm_tc(p, nsl, s);
    return nsl;
}
#endif

// ---- BatchNorm over the batch, forward -----------------------------------------------------------------------------------
// U[b][c] = sum_z part[z][b][c] (fixed order; written out, the backward needs it); train: batch statistics of U[:,c]
// (two-pass, double) -> finalisation + running statistics; eval: folded running statistics; H = relu(scale*U + shift).
// limit: an activation beyond it (the fp16 operand range of the tensor-core GEMM that consumes H) is replaced by NaN
// instead of being clamped silently.
// block = 1024 = BNH_CH channels x BNH_LANES row lanes (lane sums in row order, lanes added in order: deterministic); a thread
// keeps its rows in registers across the three passes when the batch allows (B <= BNH_LANES * BNH_RMAX), else re-reads U.
constexpr int BNH_CH = 8, BNH_LANES = 128, BNH_RMAX = 8;

__global__ void __launch_bounds__(1024) k_bn_head_fwd(const float* __restrict__ part, int nsl, int B, int C, int train,
                                                      const float* bias, pgpd_bn bn, BnState st, float limit,
                                                      float* __restrict__ U, float* __restrict__ Hout) {
    __shared__ double sh[BNH_LANES][BNH_CH + 1];
    __shared__ double smean[BNH_CH];
    __shared__ float s_sc[BNH_CH], s_sf[BNH_CH];
    const int tid = (int)threadIdx.x, cx = tid & (BNH_CH - 1), ry = tid >> 3;
    const int c = (int)blockIdx.x * BNH_CH + cx;
    const size_t slice = (size_t)B * C;
    const bool inreg = B <= BNH_LANES * BNH_RMAX;
    float ur[BNH_RMAX];
    double s = 0.0;
    if (c < C) {
        if (inreg) {
#pragma unroll
            for (int i = 0; i < BNH_RMAX; ++i) {
                const int b = ry + BNH_LANES * i;
                ur[i] = 0.f;
                if (b < B) {
                    const size_t e = (size_t)b * C + c;
                    float u = part[e];
#pragma unroll 8
                    for (int z = 1; z < nsl; ++z) u += part[(size_t)z * slice + e];      // loads are independent of the running sum
                    U[e] = u;
                    ur[i] = u;
                    s += (double)u;
                }
            }
        } else {
            for (int b = ry; b < B; b += BNH_LANES) {
                const size_t e = (size_t)b * C + c;
                float u = part[e];
#pragma unroll 8
                for (int z = 1; z < nsl; ++z) u += part[(size_t)z * slice + e];
                U[e] = u;
                s += (double)u;
            }
        }
    }
    if (train) {
        sh[ry][cx] = s;
        __syncthreads();
        if (ry == 0) {
            double t = 0.0;
#pragma unroll 8
            for (int q = 0; q < BNH_LANES; ++q) t += sh[q][cx];
            smean[cx] = t / B;
        }
        __syncthreads();
        const double mean = smean[cx];
        double v = 0.0;
        if (c < C) {
            if (inreg) {
#pragma unroll
                for (int i = 0; i < BNH_RMAX; ++i)
                    if (ry + BNH_LANES * i < B) { const double d = (double)ur[i] - mean; v += d * d; }
            } else {
                for (int b = ry; b < B; b += BNH_LANES) { const double d = (double)U[(size_t)b * C + c] - mean; v += d * d; }
            }
        }
        sh[ry][cx] = v;
        __syncthreads();
        if (ry == 0 && c < C) {
            double t = 0.0;
#pragma unroll 8
            for (int q = 0; q < BNH_LANES; ++q) t += sh[q][cx];
            bn_finalize_train(c, mean, t / B, (double)B, bias, bn, st);
            s_sc[cx] = st.scale[c]; s_sf[cx] = st.shift[c];
        }
    } else if (ry == 0 && c < C) {
        const float rstd = 1.0f / sqrtf(bn.running_var[c] + BN_EPS);
        const float sc = bn.gamma[c] * rstd;
        const float bb = bias ? bias[c] : 0.f;
        s_sc[cx] = sc; s_sf[cx] = bn.beta[c] + sc * (bb - bn.running_mean[c]);
    }
    __syncthreads();
    if (c < C) {
        const float sc = s_sc[cx], sf = s_sf[cx];
        const float qnan = __uint_as_float(0x7FC00000u);
        if (inreg) {
#pragma unroll
            for (int i = 0; i < BNH_RMAX; ++i) {
                const int b = ry + BNH_LANES * i;
                if (b < B) {
                    float h = relu_nan(sc * ur[i] + sf);
                    if (!(h <= limit)) h = qnan;
                    Hout[(size_t)b * C + c] = h;
                }
            }
        } else {
            for (int b = ry; b < B; b += BNH_LANES) {
                float h = relu_nan(sc * U[(size_t)b * C + c] + sf);
                if (!(h <= limit)) h = qnan;
                Hout[(size_t)b * C + c] = h;
            }
        }
    }
}

// ---- fc3 + bias (+ identity of the T-Net, pointnet.py:39-43) (+ log_softmax, pointnet.py:194) ------------------------------
// one warp per cloud, fixed shuffle tree (the result of a cloud does not depend on the batch it is in).
//   out [B][J]: fc3 output incl. bias / identity (kept: logits / t9);  user1: copy of out (STN: trans) or null;
//   logp / user2: log_softmax(out) (classifier) or null.
__global__ void __launch_bounds__(256) k_fc3_out(const float* __restrict__ Hm2, const float* __restrict__ W, const float* __restrict__ bias,
                                                 int B, int J, int add_identity, float* __restrict__ out, float* __restrict__ user1,
                                                 float* __restrict__ logp, float* __restrict__ user2) {
    const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
    const int b = (int)blockIdx.x * 8 + warp;
    if (b >= B) return;
    float h[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) h[u] = Hm2[(size_t)b * H2 + lane + 32 * u];
    float mx = -INFINITY;
    for (int j = 0; j < J; ++j) {
        float s = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) s = fmaf(h[u], W[(size_t)j * H2 + lane + 32 * u], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        s += bias[j];
        if (add_identity && (j % 4 == 0)) s += 1.f;          // entries 0, 4, 8 of the flattened 3x3
        if (lane == 0) { out[(size_t)b * J + j] = s; if (user1) user1[(size_t)b * J + j] = s; }
        mx = s > mx || s != s ? s : mx;                       // NaN wins (log_softmax of a row with NaN is NaN)
    }
    if (!logp) return;
    __syncwarp();
    // log_softmax over J values written by lane 0 of this warp: every lane re-reads them (J is small: k classes)
    float sum = 0.f;
    for (int j = 0; j < J; ++j) sum += expf(out[(size_t)b * J + j] - mx);
    const float lse = mx + logf(sum);
    for (int j = lane; j < J; j += 32) {
        const float v = out[(size_t)b * J + j] - lse;
        logp[(size_t)b * J + j] = v;
        if (user2) user2[(size_t)b * J + j] = v;
    }
}

// dlogits = dlogp - softmax * sum_j dlogp
__global__ void k_log_softmax_bwd(const float* __restrict__ logp, const float* __restrict__ dlogp, int B, int K,
                                  float* __restrict__ dlogits) {
    int b = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (b >= B) return;
    float s = 0.f;
    for (int j = 0; j < K; ++j) s += dlogp[(size_t)b * K + j];
    for (int j = 0; j < K; ++j) dlogits[(size_t)b * K + j] = dlogp[(size_t)b * K + j] - expf(logp[(size_t)b * K + j]) * s;
}

// ---- fc3 backward: dW3[j][i] = sum_b dO[b][j] H2[b][i];  db3[j] = sum_b dO[b][j];  dz2[b][i] = (sum_j dO[b][j] W3[j][i]) [H2 > 0]
// block = 1024 = 256 columns i x 4 cloud lanes.  blocks [0, J): one output row j of dW3 (each lane sums its clouds in order, the 4
// lanes are added in order) + db3[j];  blocks [J, J + ceil(B/32)): dz2 of 32 clouds (8 per lane).
__global__ void __launch_bounds__(1024) k_fc3_bwd(const float* __restrict__ dO, const float* __restrict__ Hm2, const float* __restrict__ W3,
                                                  int B, int J, float* __restrict__ dW3, float* __restrict__ db3, float* __restrict__ DZ2) {
    __shared__ float s_a[4][H2];
    __shared__ double s_b[4];
    const int i = (int)threadIdx.x & 255, ln = (int)threadIdx.x >> 8;
    if ((int)blockIdx.x < J) {
        const int j = (int)blockIdx.x;
        float a = 0.f;
        double bsum = 0.0;
#pragma unroll 8
        for (int b = ln; b < B; b += 4) {
            const float d = dO[(size_t)b * J + j];
            a = fmaf(d, Hm2[(size_t)b * H2 + i], a);
            bsum += (double)d;
        }
        s_a[ln][i] = a;
        if (i == 0) s_b[ln] = bsum;
        __syncthreads();
        if (ln == 0) {
            dW3[(size_t)j * H2 + i] = ((s_a[0][i] + s_a[1][i]) + s_a[2][i]) + s_a[3][i];
            if (i == 0) db3[j] = (float)(((s_b[0] + s_b[1]) + s_b[2]) + s_b[3]);
        }
        return;
    }
    const int b0 = ((int)blockIdx.x - J) * 32 + ln * 8;
    for (int bb = 0; bb < 8 && b0 + bb < B; ++bb) {
        const int b = b0 + bb;
        float s = 0.f;
        for (int j = 0; j < J; ++j) s = fmaf(dO[(size_t)b * J + j], W3[(size_t)j * H2 + i], s);
        DZ2[(size_t)b * H2 + i] = Hm2[(size_t)b * H2 + i] > 0.f ? s : 0.f;
    }
}

// ---- BatchNorm-over-batch backward -------------------------------------------------------------------------------------------
// dz[b][c] = sum_z part[z][b][c] masked by H[b][c] > 0 (part != null: the dZ GEMM's split-K partials) or DZ as given;
// sums -> dgamma, dbeta; dU = s*(dz - m1 - yhat*m2) written to DZ; db (the Linear bias feeding this BatchNorm) = 0;
// amax_part[block] = max |dU| of the block (operand scale of the tcgen05 GEMMs that consume dU).
// block = 1024 = BNH_CH channels x BNH_LANES row lanes; rows kept in registers between the two passes when B allows.
__device__ __forceinline__ void bn_head_bwd_block(int blk, const float* __restrict__ part, int nsl, const float* __restrict__ Hmask,
                                                  float* __restrict__ DZ, const float* __restrict__ U, int B, int C, BnState st,
                                                  float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ db,
                                                  unsigned* __restrict__ amax_part) {
    __shared__ double sh1[BNH_LANES][BNH_CH + 1], sh2[BNH_LANES][BNH_CH + 1];
    __shared__ float sm1[BNH_CH], sm2[BNH_CH];
    __shared__ float smx[32];
    const int tid = (int)threadIdx.x, cx = tid & (BNH_CH - 1), ry = tid >> 3;
    const int c = blk * BNH_CH + cx;
    const float mu = c < C ? st.mean[c] : 0.f, r = c < C ? st.rstd[c] : 0.f, sc = c < C ? st.scale[c] : 0.f;
    const size_t slice = (size_t)B * C;
    const bool inreg = B <= BNH_LANES * BNH_RMAX;
    float dzr[BNH_RMAX], yhr[BNH_RMAX];
    double s1 = 0.0, s2 = 0.0;
    if (c < C) {
        if (inreg) {
#pragma unroll
            for (int i = 0; i < BNH_RMAX; ++i) {
                const int b = ry + BNH_LANES * i;
                dzr[i] = 0.f; yhr[i] = 0.f;
                if (b < B) {
                    const size_t e = (size_t)b * C + c;
                    float dzf;
                    if (part) {
                        dzf = part[e];
#pragma unroll 8
                        for (int z = 1; z < nsl; ++z) dzf += part[(size_t)z * slice + e];
                        if (!(Hmask[e] > 0.f)) dzf = 0.f;
                    } else dzf = DZ[e];
                    const float yh = (U[e] - mu) * r;
                    dzr[i] = dzf; yhr[i] = yh;
                    s1 += (double)dzf; s2 += (double)dzf * (double)yh;
                }
            }
        } else {
            for (int b = ry; b < B; b += BNH_LANES) {
                const size_t e = (size_t)b * C + c;
                float dzf;
                if (part) {
                    dzf = part[e];
#pragma unroll 8
                    for (int z = 1; z < nsl; ++z) dzf += part[(size_t)z * slice + e];
                    if (!(Hmask[e] > 0.f)) dzf = 0.f;
                    DZ[e] = dzf;
                } else dzf = DZ[e];
                const double yhat = (double)((U[e] - mu) * r);
                s1 += (double)dzf; s2 += (double)dzf * yhat;
            }
        }
    }
    sh1[ry][cx] = s1; sh2[ry][cx] = s2;
    __syncthreads();
    if (ry == 0) {
        double t1 = 0.0, t2 = 0.0;
#pragma unroll 8
        for (int q = 0; q < BNH_LANES; ++q) { t1 += sh1[q][cx]; t2 += sh2[q][cx]; }
        if (c < C) { dgamma[c] = (float)t2; dbeta[c] = (float)t1; if (db) db[c] = 0.f; }
        sm1[cx] = (float)(t1 / B); sm2[cx] = (float)(t2 / B);
    }
    __syncthreads();
    float mx = 0.f;
    if (c < C) {
        const float m1 = sm1[cx], m2 = sm2[cx];
        if (inreg) {
#pragma unroll
            for (int i = 0; i < BNH_RMAX; ++i) {
                const int b = ry + BNH_LANES * i;
                if (b < B) {
                    const float du = sc * (dzr[i] - m1 - yhr[i] * m2);
                    DZ[(size_t)b * C + c] = du;
                    mx = fmaxf(mx, fabsf(du));
                }
            }
        } else {
            for (int b = ry; b < B; b += BNH_LANES) {
                const size_t e = (size_t)b * C + c;
                const float yhat = (U[e] - mu) * r;
                const float du = sc * (DZ[e] - m1 - yhat * m2);
                DZ[e] = du;
                mx = fmaxf(mx, fabsf(du));
            }
        }
    }
    if (amax_part) {     // max |dU| of this block (max is order-independent: deterministic)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((tid & 31) == 0) smx[tid >> 5] = mx;
        __syncthreads();
        if (tid == 0) {
            float m = smx[0];
            for (int q = 1; q < 32; ++q) m = fmaxf(m, smx[q]);
            amax_part[blk] = __float_as_uint(m);
        }
    }
}

__global__ void __launch_bounds__(1024) k_bn_head_bwd(float* __restrict__ DZ, const float* __restrict__ U, int B, int C, BnState st,
                                                      float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ db,
                                                      unsigned* __restrict__ amax_part) {
    bn_head_bwd_block((int)blockIdx.x, nullptr, 0, nullptr, DZ, U, B, C, st, dgamma, dbeta, db, amax_part);
}

// blocks [0, nbn): BatchNorm backward of the layer below, reading the dZ GEMM's partials (sum + ReLU mask);
// remaining blocks: dW[i] = sum_z partW[z][i] (the dW GEMM's partials)
__global__ void __launch_bounds__(1024) k_head_mid(int nbn, const float* __restrict__ partZ, int nslZ, const float* __restrict__ Hmask,
                                                   float* __restrict__ DZ, const float* __restrict__ U, int B, int C, BnState st,
                                                   float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ db,
                                                   unsigned* __restrict__ amax_part,
                                                   const float* __restrict__ partW, int nslW, size_t nW, float* __restrict__ dW) {
    if ((int)blockIdx.x < nbn) {
        bn_head_bwd_block((int)blockIdx.x, partZ, nslZ, Hmask, DZ, U, B, C, st, dgamma, dbeta, db, amax_part);
        return;
    }
    const size_t i = (size_t)((int)blockIdx.x - nbn) * 1024 + threadIdx.x;
    if (i >= nW) return;
    float v = partW[i];
#pragma unroll 8
    for (int z = 1; z < nslW; ++z) v += partW[(size_t)z * nW + i];
    dW[i] = v;
}

// out1[i] = sum_z part1[z][i] (i < n1);  out2[i] = sum_z part2[z][i] (i < n2)
__global__ void __launch_bounds__(1024) k_finish2(const float* __restrict__ part1, int nsl1, size_t n1, float* __restrict__ out1,
                                                  const float* __restrict__ part2, int nsl2, size_t n2, float* __restrict__ out2) {
    size_t i = (size_t)blockIdx.x * 1024 + threadIdx.x;
    const size_t n1r = ((n1 + 1023) / 1024) * 1024;
    if (i < n1r) {
        if (i < n1) {
            float v = part1[i];
#pragma unroll 8
            for (int z = 1; z < nsl1; ++z) v += part1[(size_t)z * n1 + i];
            out1[i] = v;
        }
        return;
    }
    i -= n1r;
    if (i < n2) {
        float v = part2[i];
#pragma unroll 8
        for (int z = 1; z < nsl2; ++z) v += part2[(size_t)z * n2 + i];
        out2[i] = v;
    }
}

struct HeadArgs {
    const pgpd_head* h;
    const float* X;      // [B][1024]
    int B, out;
    bool train, is_stn;
    cudaStream_t stream;
    bool use_tc;         // tcgen05 GEMMs for fc1 / fc2 (never in the emulator build)
};

// X -> w.out (logits, or t9 + identity); user_out: STN: trans [B][9]; classifier: log-probs [B][k].  logp_keep: [B][k] kept for the
// backward (classifier) or null.
inline void head_forward(const HeadArgs& a, HeadWs& w, float* user_out, float* logp_keep) {
    const pgpd_head& h = *a.h;
    cudaStream_t s = a.stream;
    const int B = a.B;
    bool tcg = false;
#ifndef PGPD_EMU
    tcg = a.use_tc;
#endif
    const float limit = tcg ? TC_ACT_LIMIT : INFINITY;
    int nsl = 0;
    // fc1: U1[b][j] = sum_i X[b][i] W1[j][i]   (8 K-slices, whatever the batch)
#ifndef PGPD_EMU
    if (tcg) {
        launch(tc::k_absmax2, dim3(HEAD_AMAX_BLOCKS, 2), dim3(256), 0, s, h.fc[0].w, (size_t)H1 * C3, h.fc[1].w, (size_t)H2 * H1, w.amax);
        nsl = run_gemm_tc(tc::GemmOp{a.X, C3, 0, nullptr, 0, tc::ACT_SCALE}, tc::GemmOp{h.fc[0].w, C3, 0, w.amax, HEAD_AMAX_BLOCKS, 1.f},
                          B, H1, C3, w.partA, 8, s);
    } else
#endif
    nsl = run_plain(ProbPlain<true, false>{a.X, h.fc[0].w, w.partA, B, H1, C3, (size_t)C3, 1, 1, (size_t)C3, 0}, s, 8);
    launch(k_bn_head_fwd, grid1d(H1, BNH_CH), dim3(1024), 0, s, (const float*)w.partA, nsl, B, H1, a.train ? 1 : 0, h.fc[0].b, h.bn[0], w.bn[0],
           limit, w.U1, w.Hm1);
    // fc2
#ifndef PGPD_EMU
    if (tcg)
        nsl = run_gemm_tc(tc::GemmOp{w.Hm1, H1, 0, nullptr, 0, tc::ACT_SCALE},
                          tc::GemmOp{h.fc[1].w, H1, 0, w.amax + HEAD_AMAX_BLOCKS, HEAD_AMAX_BLOCKS, 1.f}, B, H2, H1, w.partB, 4, s);
    else
#endif
    nsl = run_plain(ProbPlain<true, false>{w.Hm1, h.fc[1].w, w.partB, B, H2, H1, (size_t)H1, 1, 1, (size_t)H1, 0}, s, 4);
    launch(k_bn_head_fwd, grid1d(H2, BNH_CH), dim3(1024), 0, s, (const float*)w.partB, nsl, B, H2, a.train ? 1 : 0, h.fc[1].b, h.bn[1], w.bn[1],
           INFINITY, w.U2, w.Hm2);
    // fc3 (+ bias, + identity for the T-Net, + log_softmax for the classifier)
    launch(k_fc3_out, grid1d(B, 8), dim3(256), 0, s, (const float*)w.Hm2, h.fc[2].w, h.fc[2].b, B, a.out, a.is_stn ? 1 : 0, w.out,
           a.is_stn ? user_out : (float*)nullptr, a.is_stn ? (float*)nullptr : logp_keep, a.is_stn ? (float*)nullptr : user_out);
}

// w.dO (gradient w.r.t. w.out) -> parameter gradients and dX [B][1024]
inline void head_backward(const HeadArgs& a, HeadWs& w, const pgpd_head_grad& g, float* dX) {
    const pgpd_head& h = *a.h;
    cudaStream_t s = a.stream;
    const int B = a.B, J3 = a.out;
    bool tcg = false;
#ifndef PGPD_EMU
    tcg = a.use_tc;
#endif
    unsigned* amax_du1 = w.amax + 2 * HEAD_AMAX_BLOCKS;        // [H1 / 8] partial max |dU1|
    unsigned* amax_du2 = amax_du1 + H1 / BNH_CH;               // [H2 / 8]
    // ---- fc3: dW3, db3, dz2
    launch(k_fc3_bwd, dim3(J3 + idiv_up(B, 32)), dim3(1024), 0, s, (const float*)w.dO, (const float*)w.Hm2, h.fc[2].w, B, J3,
           g.fc[2].dw, g.fc[2].db, w.DZ2);
    launch(k_bn_head_bwd, grid1d(H2, BNH_CH), dim3(1024), 0, s, w.DZ2, (const float*)w.U2, B, H2, w.bn[1], g.bn[1].dgamma, g.bn[1].dbeta,
           g.fc[1].db, tcg ? amax_du2 : (unsigned*)nullptr);
    // ---- fc2 (w.DZ2 now holds dU2):  dW2 = dU2^T Hm1,  dz1 = (dU2 W2) masked by Hm1 > 0
    int nslW = 0, nslZ = 0;
#ifndef PGPD_EMU
    if (tcg) {
        nslW = run_gemm_tc(tc::GemmOp{w.DZ2, H2, 1, amax_du2, H2 / BNH_CH, 1.f}, tc::GemmOp{w.Hm1, H1, 1, nullptr, 0, tc::ACT_SCALE}, H2, H1, B,
                           w.partA, 0, s);
        nslZ = run_gemm_tc(tc::GemmOp{w.DZ2, H2, 0, amax_du2, H2 / BNH_CH, 1.f},
                           tc::GemmOp{h.fc[1].w, H1, 1, w.amax + HEAD_AMAX_BLOCKS, HEAD_AMAX_BLOCKS, 1.f}, B, H1, H2, w.partB, 0, s);
    } else
#endif
    {
        nslW = run_plain(ProbPlain<false, true>{w.DZ2, w.Hm1, w.partA, H2, H1, B, 1, (size_t)H2, (size_t)H1, 1, 0}, s, 0);
        nslZ = run_plain(ProbPlain<true, true>{w.DZ2, h.fc[1].w, w.partB, B, H1, H2, (size_t)H2, 1, (size_t)H1, 1, 0}, s, 0);
    }
    {
        const int nbn = idiv_up(H1, BNH_CH);
        const size_t nW = (size_t)H2 * H1;
        launch(k_head_mid, dim3(nbn + (unsigned)((nW + 1023) / 1024)), dim3(1024), 0, s, nbn, (const float*)w.partB, nslZ, (const float*)w.Hm1,
               w.DZ1, (const float*)w.U1, B, H1, w.bn[0], g.bn[0].dgamma, g.bn[0].dbeta, g.fc[0].db, tcg ? amax_du1 : (unsigned*)nullptr,
               (const float*)w.partA, nslW, nW, g.fc[1].dw);
    }
    // ---- fc1 (w.DZ1 now holds dU1):  dW1 = dU1^T X,  dX = dU1 W1
    int nsl1 = 0, nslX = 0;
#ifndef PGPD_EMU
    if (tcg) {
        nsl1 = run_gemm_tc(tc::GemmOp{w.DZ1, H1, 1, amax_du1, H1 / BNH_CH, 1.f}, tc::GemmOp{a.X, C3, 1, nullptr, 0, tc::ACT_SCALE}, H1, C3, B,
                           w.partA, 0, s);
        nslX = run_gemm_tc(tc::GemmOp{w.DZ1, H1, 0, amax_du1, H1 / BNH_CH, 1.f}, tc::GemmOp{h.fc[0].w, C3, 1, w.amax, HEAD_AMAX_BLOCKS, 1.f}, B, C3, H1,
                           w.partB, 0, s);
    } else
#endif
    {
        nsl1 = run_plain(ProbPlain<false, true>{w.DZ1, a.X, w.partA, H1, C3, B, 1, (size_t)H1, (size_t)C3, 1, 0}, s, 0);
        nslX = run_plain(ProbPlain<true, true>{w.DZ1, h.fc[0].w, w.partB, B, C3, H1, (size_t)H1, 1, (size_t)C3, 1, 0}, s, 0);
    }
    {
        const size_t n1 = (size_t)H1 * C3, n2 = (size_t)B * C3;
        launch(k_finish2, dim3((unsigned)((n1 + 1023) / 1024 + (n2 + 1023) / 1024)), dim3(1024), 0, s, (const float*)w.partA, nsl1, n1,
               g.fc[0].dw, (const float*)w.partB, nslX, n2, dX);
    }
}

}  // namespace pgpd
