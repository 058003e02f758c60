// head.cuh -- the small dense heads 1024->512->256->out with BatchNorm over the batch:
//   STN3d regression head     pointnet.py:35-44  (out = 9, + identity)
//   PointNetCls classifier    pointnet.py:191-194 (out = k, log_softmax)
// CUDA-core fp32 kernels (2.6 MFLOP per grasp: launch count, not flops, is what matters here).
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"
#ifndef PGPD_EMU
#include "tc_gemm.cuh"
#endif

namespace pgpd {

struct HeadWs {
    float* U1;        // [B][512]  bias-free pre-activation of fc1
    float* U2;        // [B][256]
    float* Hm1;       // [B][512]  relu(bn(U1)), materialised
    float* Hm2;       // [B][256]
    float* out;       // [B][out]  fc3 output (+bias, + identity for the STN head) = logits / t9
    BnState bn[2];
    float* part;      // split-K partials: HEAD_PART_ELEMS floats
    // backward scratch
    float* DZ1;       // [B][512]  dz1, then (in place) dU1
    float* DZ2;       // [B][256]
    float* dO;        // [B][out]  gradient w.r.t. fc3 output
    unsigned* amax;   // [4] bit patterns of max |x|: fc1.weight, fc2.weight, dU1, dU2 (operand scales of the tcgen05 GEMMs)
};

constexpr size_t HEAD_PART_ELEMS = (size_t)4 << 20;   // 16 MB of fp32 partials

inline void plan_head(Carver& c, HeadWs& w, int B, int out, bool backward) {
    w.U1 = c.take<float>((size_t)B * H1);
    w.U2 = c.take<float>((size_t)B * H2);
    w.Hm1 = c.take<float>((size_t)B * H1);
    w.Hm2 = c.take<float>((size_t)B * H2);
    w.out = c.take<float>((size_t)B * out);
    w.bn[0].carve(c, H1); w.bn[1].carve(c, H2);
    w.part = c.take<float>(HEAD_PART_ELEMS);
    w.amax = c.take<unsigned>(4);
    if (backward) {
        w.DZ1 = c.take<float>((size_t)B * H1);
        w.DZ2 = c.take<float>((size_t)B * H2);
        w.dO = c.take<float>((size_t)B * out);
    }
}

// ---- plain fp32 GEMM on row-major matrices with optional split-K ---------------------------------------
//   C[m][n] = sum_k A(m,k) B(k,n);  A(m,k) = A[m*sam + k*sak], B(k,n) = Bm[k*sbk + n*sbn]
// epilogue (only when the whole K range is handled by one block, ksl == 0):
//   EPI_BIAS: + bias[n] (+1 on the diagonal of a flattened 3x3 when add_identity);  EPI_MASK: zero where mask[m][n] <= 0
enum { EPI_NONE = 0, EPI_BIAS = 1, EPI_MASK = 2 };
template <bool AK, bool BNF>
struct ProbPlain {
    static constexpr bool A_KFAST = AK, B_NFAST = BNF;
    static constexpr int SCRATCH = 0;
    using Cfg = CfgSmall;
    const float* A; const float* Bm; float* C; float* part;
    int Mr, Nc, K; size_t sam, sak, sbk, sbn;
    int ksl;            // 0: no split; else K-slice length, blockIdx.z selects the slice, output goes to part[z][Mr][Nc]
    int epi; const float* bias; int add_identity; const float* mask;
    struct Blk { int m0, n0, k0, k1; };
    __device__ void setup(Blk& b) const {
        b.m0 = (int)blockIdx.y * Cfg::BM; b.n0 = (int)blockIdx.x * Cfg::BN; b.k0 = 0; b.k1 = K;
        if (ksl > 0) { b.k0 = (int)blockIdx.z * ksl; b.k1 = b.k0 + ksl < K ? b.k0 + ksl : K; }
    }
    __device__ void prologue(const Blk&, float*) const {}
    __device__ float loadA(const Blk&, const float*, int m, int k) const { return m < Mr ? A[(size_t)m * sam + (size_t)k * sak] : 0.f; }
    __device__ float loadB(const Blk&, const float*, int k, int n) const { return n < Nc ? Bm[(size_t)k * sbk + (size_t)n * sbn] : 0.f; }
    __device__ void epilogue(const Blk& b, const float*, float (&acc)[Cfg::TM][Cfg::TN], int ty, int tx, void*) const {
        float* out = ksl > 0 ? part + (size_t)blockIdx.z * Mr * Nc : C;
#pragma unroll
        for (int i = 0; i < Cfg::TM; ++i) {
            const int m = b.m0 + Cfg::row_of(ty, i);
            if (m >= Mr) continue;
#pragma unroll
            for (int j = 0; j < Cfg::TN; ++j) {
                const int n = b.n0 + Cfg::col_of(tx, j);
                if (n >= Nc) continue;
                float v = acc[i][j];
                if (ksl == 0) v = finish(v, m, n);
                out[(size_t)m * Nc + n] = v;
            }
        }
    }
    __device__ float finish(float v, int m, int n) const {
        if (epi == EPI_BIAS) { v += bias[n]; if (add_identity && (n % 4 == 0)) v += 1.f; }   // entries 0,4,8 (pointnet.py:39-43)
        else if (epi == EPI_MASK) { if (!(mask[(size_t)m * Nc + n] > 0.f)) v = 0.f; }
        return v;
    }
};

// sums the split-K partials in a fixed order and applies the epilogue
template <bool AK, bool BNF>
__global__ void k_splitk_finish(ProbPlain<AK, BNF> p, int nsl) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)p.Mr * p.Nc;
    if (i >= total) return;
    float v = 0.f;
    for (int z = 0; z < nsl; ++z) v += p.part[(size_t)z * total + i];
    p.C[i] = p.finish(v, (int)(i / p.Nc), (int)(i % p.Nc));
}

// fixed_nsl > 0: use exactly that many K slices (forward GEMMs: the summation order must not depend on the batch
// size, so that eval outputs are bit-identical however the clouds are batched); 0: pick by occupancy.
template <bool AK, bool BNF>
inline void run_plain(ProbPlain<AK, BNF> p, cudaStream_t s, int fixed_nsl = 0) {
    const int tm = idiv_up(p.Mr, CfgSmall::BM), tn = idiv_up(p.Nc, CfgSmall::BN);
    int nsl = 1;
    if (fixed_nsl > 0) {
        nsl = fixed_nsl;
        while (nsl > 1 && (size_t)nsl * p.Mr * p.Nc > HEAD_PART_ELEMS) nsl /= 2;    // huge batches: fewer slices (still batch-size independent below 4M outputs)
    } else {
        // enough blocks to occupy the chip (>= ~2 per SM), K slices of at least 64
        while (tm * tn * nsl < 296 && p.K / (nsl * 2) >= 64 && (size_t)(nsl * 2) * p.Mr * p.Nc <= HEAD_PART_ELEMS) nsl *= 2;
    }
    p.ksl = nsl > 1 ? idiv_up(p.K, nsl) : 0;
    launch_gemm<CfgSmall>(p, dim3(tn, tm, nsl), s);
    if (nsl > 1) launch(k_splitk_finish<AK, BNF>, grid1d((size_t)p.Mr * p.Nc, 256), dim3(256), 0, s, p, nsl);
}

// batch statistics of U[:,c] (two-pass, double) -> BatchNorm finalisation -> H = relu(scale*U + shift).
// block = 32 channels x 32 row lanes (fixed-order shared-memory reduction: deterministic)
__global__ void k_bn_batch_stats_apply(const float* __restrict__ U, int B, int C, const float* bias, pgpd_bn bn, BnState st,
                                       float* __restrict__ Hout) {
    __shared__ double sh[32][33];
    __shared__ double smean[32];
    const int tid = (int)threadIdx.x, cx = tid & 31, ry = tid >> 5;
    const int c = (int)blockIdx.x * 32 + cx;
    double s = 0.0;
    if (c < C) {
#pragma unroll 4
        for (int b = ry; b < B; b += 32) s += (double)U[(size_t)b * C + c];
    }
    sh[ry][cx] = s;
    __syncthreads();
    if (ry == 0) {
        double t = 0.0;
        for (int q = 0; q < 32; ++q) t += sh[q][cx];
        smean[cx] = t / B;
    }
    __syncthreads();
    const double mean = smean[cx];
    double v = 0.0;
    if (c < C)
        for (int b = ry; b < B; b += 32) { double d = (double)U[(size_t)b * C + c] - mean; v += d * d; }
    sh[ry][cx] = v;
    __syncthreads();
    if (ry == 0 && c < C) {
        double t = 0.0;
        for (int q = 0; q < 32; ++q) t += sh[q][cx];
        bn_finalize_train(c, mean, t / B, (double)B, bias, bn, st);
    }
    __syncthreads();
    if (c < C) {
        const float sc = st.scale[c], sf = st.shift[c];
        for (int b = ry; b < B; b += 32) Hout[(size_t)b * C + c] = fmaxf(sc * U[(size_t)b * C + c] + sf, 0.f);
    }
}

// H = relu(scale*U + shift) (eval mode, after k_bn_eval_affine)
__global__ void k_bn_apply(const float* __restrict__ U, size_t total, int C, BnState st, float* __restrict__ Hout) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    Hout[i] = fmaxf(st.scale[c] * U[i] + st.shift[c], 0.f);
}

// log_softmax over the last dim (pointnet.py:194); thread = row
__global__ void k_log_softmax(const float* __restrict__ logits, int B, int K, float* __restrict__ logp) {
    int b = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (b >= B) return;
    const float* l = logits + (size_t)b * K;
    float m = l[0];
    for (int j = 1; j < K; ++j) m = fmaxf(m, l[j]);
    float s = 0.f;
    for (int j = 0; j < K; ++j) s += expf(l[j] - m);
    float lse = m + logf(s);
    for (int j = 0; j < K; ++j) logp[(size_t)b * K + j] = l[j] - lse;
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

// out[j] = sum_b G[b][j]   (bias gradient of fc3); block = column, 256 row lanes summed in a fixed order
__global__ void k_colsum(const float* __restrict__ G, int B, int J, float* __restrict__ out) {
    __shared__ double sh[256];
    const int j = (int)blockIdx.x, tid = (int)threadIdx.x;
    double s = 0.0;
    for (int b = tid; b < B; b += 256) s += (double)G[(size_t)b * J + j];
    sh[tid] = s;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (tid < st) sh[tid] += sh[tid + st];
        __syncthreads();
    }
    if (tid == 0) out[j] = (float)sh[0];
}

// BatchNorm-over-batch backward: sums -> dgamma, dbeta; then dU = s*(dz - m1 - yhat*m2) written IN PLACE over dz.
// block = 32 channels x 32 row lanes.
__global__ void k_bn_batch_bwd_apply(float* __restrict__ DZ, const float* __restrict__ U, int B, int C, BnState st,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta, unsigned* __restrict__ amax) {
    __shared__ double sh1[32][33], sh2[32][33];
    __shared__ float sm1[32], sm2[32];
    const int tid = (int)threadIdx.x, cx = tid & 31, ry = tid >> 5;
    const int c = (int)blockIdx.x * 32 + cx;
    const float mu = c < C ? st.mean[c] : 0.f, r = c < C ? st.rstd[c] : 0.f, sc = c < C ? st.scale[c] : 0.f;
    double s1 = 0.0, s2 = 0.0;
    if (c < C) {
        for (int b = ry; b < B; b += 32) {
            double dz = (double)DZ[(size_t)b * C + c];
            double yhat = (double)((U[(size_t)b * C + c] - mu) * r);
            s1 += dz; s2 += dz * yhat;
        }
    }
    sh1[ry][cx] = s1; sh2[ry][cx] = s2;
    __syncthreads();
    if (ry == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int q = 0; q < 32; ++q) { t1 += sh1[q][cx]; t2 += sh2[q][cx]; }
        if (c < C) { dgamma[c] = (float)t2; dbeta[c] = (float)t1; }
        sm1[cx] = (float)(t1 / B); sm2[cx] = (float)(t2 / B);
    }
    __syncthreads();
    float mx = 0.f;
    if (c < C) {
        const float m1 = sm1[cx], m2 = sm2[cx];
        for (int b = ry; b < B; b += 32) {
            const size_t i = (size_t)b * C + c;
            const float yhat = (U[i] - mu) * r;
            const float du = sc * (DZ[i] - m1 - yhat * m2);
            DZ[i] = du;
            mx = fmaxf(mx, fabsf(du));
        }
    }
    if (amax) {     // max |dU|: operand scale of the tcgen05 GEMMs that consume dU (max is order-independent: deterministic)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (cx == 0) atomicMax(amax, __float_as_uint(mx));
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

#ifndef PGPD_EMU
// C[M][N] = sum_k A(m,k) B(n,k) on the tensor cores; split-K partials go through `part` and are summed in a fixed order.
// nsl_fixed > 0: exactly that many K slices (forward: the summation order must not depend on the batch size).
inline void run_gemm_tc(tc::GemmOp A, tc::GemmOp Bo, int M, int N, int K, float* C, float* part, const float* mask,
                        int nsl_fixed, cudaStream_t s) {
    const int tiles = idiv_up(M, tc::GM_T) * idiv_up(N, tc::GM_T);
    int nsl = nsl_fixed > 0 ? nsl_fixed : 1;
    if (nsl_fixed <= 0)
        while (tiles * nsl < 120 && K / (nsl * 2) >= tc::GM_KC && (size_t)(nsl * 2) * M * N <= HEAD_PART_ELEMS) nsl *= 2;
    while (nsl > 1 && (size_t)nsl * M * N > HEAD_PART_ELEMS) nsl /= 2;
    int kslice = idiv_up(idiv_up(K, nsl), tc::GM_KC) * tc::GM_KC;
    nsl = idiv_up(K, kslice);                       // no empty slices
    tc::GemmParams p{A, Bo, M, N, K, kslice, nsl > 1 ? part : C, mask};
    tc::launch_gemm_tc(p, nsl, s);
    if (nsl > 1) {
        ProbPlain<true, true> f{nullptr, nullptr, C, part, M, N, K, 0, 0, 0, 0, kslice, mask ? EPI_MASK : EPI_NONE, nullptr, 0, mask};
        launch(k_splitk_finish<true, true>, grid1d((size_t)M * N, 256), dim3(256), 0, s, f, nsl);
    }
}
#endif

// X -> w.out  (logits, or t9 + identity)
inline void head_forward(const HeadArgs& a, HeadWs& w) {
    const pgpd_head& h = *a.h;
    cudaStream_t s = a.stream;
    const int B = a.B;
    // fc1: U1[b][j] = sum_i X[b][i] W1[j][i]
#ifndef PGPD_EMU
    if (a.use_tc) {
        cudaMemsetAsync(w.amax, 0, 4 * sizeof(unsigned), s);
        launch(tc::k_absmax2, dim3(64, 2), dim3(256), 0, s, h.fc[0].w, (size_t)H1 * C3, h.fc[1].w, (size_t)H2 * H1, w.amax);
        run_gemm_tc(tc::GemmOp{a.X, C3, 0, nullptr, tc::ACT_SCALE}, tc::GemmOp{h.fc[0].w, C3, 0, w.amax + 0, 1.f}, B, H1, C3,
                    w.U1, w.part, nullptr, 8, s);
    } else
#endif
    run_plain(ProbPlain<true, false>{a.X, h.fc[0].w, w.U1, w.part, B, H1, C3, (size_t)C3, 1, 1, (size_t)C3, 0, EPI_NONE, nullptr, 0, nullptr}, s, 8);
    if (a.train) launch(k_bn_batch_stats_apply, grid1d(H1, 32), dim3(1024), 0, s, (const float*)w.U1, B, H1, h.fc[0].b, h.bn[0], w.bn[0], w.Hm1);
    else {
        launch(k_bn_eval_affine, grid1d(H1, 128), dim3(128), 0, s, H1, h.fc[0].b, h.bn[0], w.bn[0]);
        launch(k_bn_apply, grid1d((size_t)B * H1, 256), dim3(256), 0, s, (const float*)w.U1, (size_t)B * H1, H1, w.bn[0], w.Hm1);
    }
    // fc2
#ifndef PGPD_EMU
    if (a.use_tc)
        run_gemm_tc(tc::GemmOp{w.Hm1, H1, 0, nullptr, tc::ACT_SCALE}, tc::GemmOp{h.fc[1].w, H1, 0, w.amax + 1, 1.f}, B, H2, H1,
                    w.U2, w.part, nullptr, 4, s);
    else
#endif
    run_plain(ProbPlain<true, false>{w.Hm1, h.fc[1].w, w.U2, w.part, B, H2, H1, (size_t)H1, 1, 1, (size_t)H1, 0, EPI_NONE, nullptr, 0, nullptr}, s, 4);
    if (a.train) launch(k_bn_batch_stats_apply, grid1d(H2, 32), dim3(1024), 0, s, (const float*)w.U2, B, H2, h.fc[1].b, h.bn[1], w.bn[1], w.Hm2);
    else {
        launch(k_bn_eval_affine, grid1d(H2, 128), dim3(128), 0, s, H2, h.fc[1].b, h.bn[1], w.bn[1]);
        launch(k_bn_apply, grid1d((size_t)B * H2, 256), dim3(256), 0, s, (const float*)w.U2, (size_t)B * H2, H2, w.bn[1], w.Hm2);
    }
    // fc3 (+ bias, + identity for the T-Net)
    run_plain(ProbPlain<true, false>{w.Hm2, h.fc[2].w, w.out, w.part, B, a.out, H2, (size_t)H2, 1, 1, (size_t)H2, 0, EPI_BIAS, h.fc[2].b,
                                     a.is_stn ? 1 : 0, nullptr}, s, 4);
}

// w.dO (gradient w.r.t. w.out) -> parameter gradients and dX [B][1024]
inline void head_backward(const HeadArgs& a, HeadWs& w, const pgpd_head_grad& g, float* dX) {
    const pgpd_head& h = *a.h;
    cudaStream_t s = a.stream;
    const int B = a.B, J3 = a.out;
    // ---- fc3:  dW3[j][i] = sum_b dO[b][j] H2[b][i] ;  db3 = colsum(dO) ;  dz2 = (dO W3) masked by H2 > 0
    run_plain(ProbPlain<false, true>{w.dO, w.Hm2, g.fc[2].dw, w.part, J3, H2, B, 1, (size_t)J3, (size_t)H2, 1, 0, EPI_NONE, nullptr, 0, nullptr}, s);
    launch(k_colsum, dim3(J3), dim3(256), 0, s, (const float*)w.dO, B, J3, g.fc[2].db);
    run_plain(ProbPlain<true, true>{w.dO, h.fc[2].w, w.DZ2, w.part, B, H2, J3, (size_t)J3, 1, (size_t)H2, 1, 0, EPI_MASK, nullptr, 0, w.Hm2}, s);
    bool tcg = false;
#ifndef PGPD_EMU
    tcg = a.use_tc;
#endif
    launch(k_bn_batch_bwd_apply, grid1d(H2, 32), dim3(1024), 0, s, w.DZ2, (const float*)w.U2, B, H2, w.bn[1], g.bn[1].dgamma, g.bn[1].dbeta,
           tcg ? w.amax + 3 : (unsigned*)nullptr);
    // ---- fc2 (w.DZ2 now holds dU2):  dW2 = dU2^T Hm1,  dz1 = (dU2 W2) masked by Hm1 > 0
    launch(k_fill, grid1d(H2, 128), dim3(128), 0, s, g.fc[1].db, (size_t)H2, 0.f);
#ifndef PGPD_EMU
    if (tcg) {
        run_gemm_tc(tc::GemmOp{w.DZ2, H2, 1, w.amax + 3, 1.f}, tc::GemmOp{w.Hm1, H1, 1, nullptr, tc::ACT_SCALE}, H2, H1, B,
                    g.fc[1].dw, w.part, nullptr, 0, s);
        run_gemm_tc(tc::GemmOp{w.DZ2, H2, 0, w.amax + 3, 1.f}, tc::GemmOp{h.fc[1].w, H1, 1, w.amax + 1, 1.f}, B, H1, H2,
                    w.DZ1, w.part, w.Hm1, 0, s);
    } else
#endif
    {
        run_plain(ProbPlain<false, true>{w.DZ2, w.Hm1, g.fc[1].dw, w.part, H2, H1, B, 1, (size_t)H2, (size_t)H1, 1, 0, EPI_NONE, nullptr, 0, nullptr}, s);
        run_plain(ProbPlain<true, true>{w.DZ2, h.fc[1].w, w.DZ1, w.part, B, H1, H2, (size_t)H2, 1, (size_t)H1, 1, 0, EPI_MASK, nullptr, 0, w.Hm1}, s);
    }
    launch(k_bn_batch_bwd_apply, grid1d(H1, 32), dim3(1024), 0, s, w.DZ1, (const float*)w.U1, B, H1, w.bn[0], g.bn[0].dgamma, g.bn[0].dbeta,
           tcg ? w.amax + 2 : (unsigned*)nullptr);
    // ---- fc1 (w.DZ1 now holds dU1):  dW1 = dU1^T X,  dX = dU1 W1
    launch(k_fill, grid1d(H1, 128), dim3(128), 0, s, g.fc[0].db, (size_t)H1, 0.f);
#ifndef PGPD_EMU
    if (tcg) {
        run_gemm_tc(tc::GemmOp{w.DZ1, H1, 1, w.amax + 2, 1.f}, tc::GemmOp{a.X, C3, 1, nullptr, tc::ACT_SCALE}, H1, C3, B,
                    g.fc[0].dw, w.part, nullptr, 0, s);
        run_gemm_tc(tc::GemmOp{w.DZ1, H1, 0, w.amax + 2, 1.f}, tc::GemmOp{h.fc[0].w, C3, 1, w.amax + 0, 1.f}, B, C3, H1,
                    dX, w.part, nullptr, 0, s);
    } else
#endif
    {
        run_plain(ProbPlain<false, true>{w.DZ1, a.X, g.fc[0].dw, w.part, H1, C3, B, 1, (size_t)H1, (size_t)C3, 1, 0, EPI_NONE, nullptr, 0, nullptr}, s);
        run_plain(ProbPlain<true, true>{w.DZ1, h.fc[0].w, dX, w.part, B, C3, H1, (size_t)H1, 1, (size_t)C3, 1, 0, EPI_NONE, nullptr, 0, nullptr}, s);
    }
}

}  // namespace pgpd
