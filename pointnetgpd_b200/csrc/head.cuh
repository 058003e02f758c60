// head.cuh -- the small dense heads 1024->512->256->out with BatchNorm over the batch:
//   STN3d regression head     pointnet.py:35-44  (out = 9, + identity)
//   PointNetCls classifier    pointnet.py:191-194 (out = k, log_softmax)
// CUDA-core fp32 kernels (2.6 MFLOP per grasp: launch count, not flops, is what matters here).
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"

namespace pgpd {

struct HeadWs {
    float* U1;        // [B][512]  bias-free pre-activation of fc1
    float* U2;        // [B][256]
    float* out;       // [B][out]  fc3 output (+bias, + identity for the STN head) = logits / t9
    BnState bn[2];
    // backward scratch
    float* DZ1;       // [B][512]
    float* DZ2;       // [B][256]
    float* m1_1; float* m2_1;   // [512]
    float* m1_2; float* m2_2;   // [256]
    float* dO;        // [B][out]  gradient w.r.t. fc3 output
};

inline void plan_head(Carver& c, HeadWs& w, int B, int out, bool backward) {
    w.U1 = c.take<float>((size_t)B * H1);
    w.U2 = c.take<float>((size_t)B * H2);
    w.out = c.take<float>((size_t)B * out);
    w.bn[0].carve(c, H1); w.bn[1].carve(c, H2);
    if (backward) {
        w.DZ1 = c.take<float>((size_t)B * H1);
        w.DZ2 = c.take<float>((size_t)B * H2);
        w.m1_1 = c.take<float>(H1); w.m2_1 = c.take<float>(H1);
        w.m1_2 = c.take<float>(H2); w.m2_2 = c.take<float>(H2);
        w.dO = c.take<float>((size_t)B * out);
    }
}

// input activation of a Linear layer: either a raw feature matrix or relu(scale*U + shift)
struct ActIn {
    const float* U; const float* scale; const float* shift; int ld;
    __device__ float at(int b, int i) const {
        float v = U[(size_t)b * ld + i];
        return scale ? fmaxf(scale[i] * v + shift[i], 0.f) : v;
    }
};

// gradient w.r.t. a Linear output: either given directly, or through a train-mode BatchNorm:
//   dU = s*(dz - m1 - yhat*m2)
struct GradOut {
    const float* DZ; const float* U; BnState st; const float* m1; const float* m2; int ld; bool bn;
    __device__ float at(int b, int j) const {
        float dz = DZ[(size_t)b * ld + j];
        if (!bn) return dz;
        float yhat = (U[(size_t)b * ld + j] - st.mean[j]) * st.rstd[j];
        return st.scale[j] * (dz - m1[j] - yhat * m2[j]);
    }
};

// U[b][j] = sum_i act(b,i) W[j][i]  (+ bias[j]) (+ 1 on the diagonal entries of a flattened 3x3)
struct ProbLinFwd {
    static constexpr bool A_KFAST = true, B_NFAST = false;
    static constexpr int SCRATCH = 0;
    using Cfg = CfgHead;
    ActIn in; const float* W; const float* bias; float* U; int B, J, I; int add_identity;
    struct Blk { int m0, n0, k0, k1; };
    __device__ void setup(Blk& b) const { b.m0 = (int)blockIdx.y * Cfg::BM; b.n0 = (int)blockIdx.x * Cfg::BN; b.k0 = 0; b.k1 = I; }
    __device__ void prologue(const Blk&, float*) const {}
    __device__ float loadA(const Blk&, const float*, int m, int k) const { return m < B ? in.at(m, k) : 0.f; }
    __device__ float loadB(const Blk&, const float*, int k, int n) const { return n < J ? W[(size_t)n * I + k] : 0.f; }
    __device__ void epilogue(const Blk& b, const float*, float (&acc)[Cfg::TM][Cfg::TN], int ty, int tx, void*) const {
#pragma unroll
        for (int i = 0; i < Cfg::TM; ++i) {
            int m = b.m0 + Cfg::row_of(ty, i);
            if (m >= B) continue;
#pragma unroll
            for (int j = 0; j < Cfg::TN; ++j) {
                int n = b.n0 + Cfg::col_of(tx, j);
                if (n >= J) continue;
                float v = acc[i][j];
                if (bias) v += bias[n];
                if (add_identity && (n % 4 == 0)) v += 1.f;   // entries 0,4,8 of the flattened 3x3 (pointnet.py:39-43)
                U[(size_t)m * J + n] = v;
            }
        }
    }
};

// batch statistics of U[:,c] (two-pass, double) -> BatchNorm finalisation.
// block = 32 channels x 8 row lanes (fixed-order shared-memory reduction: deterministic)
__global__ void k_bn_batch_stats(const float* __restrict__ U, int B, int C, const float* bias, pgpd_bn bn, BnState st) {
    __shared__ double sh[8][33];
    __shared__ double smean[32];
    const int tid = (int)threadIdx.x, cx = tid & 31, ry = tid >> 5;
    const int c = (int)blockIdx.x * 32 + cx;
    double s = 0.0;
    if (c < C)
        for (int b = ry; b < B; b += 8) s += (double)U[(size_t)b * C + c];
    sh[ry][cx] = s;
    __syncthreads();
    if (ry == 0) {
        double t = 0.0;
        for (int q = 0; q < 8; ++q) t += sh[q][cx];
        smean[cx] = t / B;
    }
    __syncthreads();
    const double mean = smean[cx];
    double v = 0.0;
    if (c < C)
        for (int b = ry; b < B; b += 8) { double d = (double)U[(size_t)b * C + c] - mean; v += d * d; }
    sh[ry][cx] = v;
    __syncthreads();
    if (ry == 0 && c < C) {
        double t = 0.0;
        for (int q = 0; q < 8; ++q) t += sh[q][cx];
        bn_finalize_train(c, mean, t / B, (double)B, bias, bn, st);
    }
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

// out[j] = sum_b G[b][j]   (bias gradient of fc3); thread = column
__global__ void k_colsum(const float* __restrict__ G, int B, int J, float* __restrict__ out) {
    int j = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (j >= J) return;
    double s = 0.0;
    for (int b = 0; b < B; ++b) s += (double)G[(size_t)b * J + j];
    out[j] = (float)s;
}

// dW[j][i] = sum_b dY(b,j) * X(b,i)
struct ProbLinBwdW {
    static constexpr bool A_KFAST = false, B_NFAST = true;
    static constexpr int SCRATCH = 0;
    using Cfg = CfgHead;
    GradOut dy; ActIn in; float* dW; int B, J, I;
    struct Blk { int m0, n0, k0, k1; };
    __device__ void setup(Blk& b) const { b.m0 = (int)blockIdx.y * Cfg::BM; b.n0 = (int)blockIdx.x * Cfg::BN; b.k0 = 0; b.k1 = B; }
    __device__ void prologue(const Blk&, float*) const {}
    __device__ float loadA(const Blk&, const float*, int m, int k) const { return m < J ? dy.at(k, m) : 0.f; }
    __device__ float loadB(const Blk&, const float*, int k, int n) const { return n < I ? in.at(k, n) : 0.f; }
    __device__ void epilogue(const Blk& b, const float*, float (&acc)[Cfg::TM][Cfg::TN], int ty, int tx, void*) const {
#pragma unroll
        for (int i = 0; i < Cfg::TM; ++i) {
            int m = b.m0 + Cfg::row_of(ty, i);
            if (m >= J) continue;
#pragma unroll
            for (int j = 0; j < Cfg::TN; ++j) {
                int n = b.n0 + Cfg::col_of(tx, j);
                if (n < I) dW[(size_t)m * I + n] = acc[i][j];
            }
        }
    }
};

// dX[b][i] = sum_j dY(b,j) W[j][i]; if the input was relu(bn(Uin)) the ReLU mask is applied and the
// result stored as dz of the previous layer, else it is the gradient of the head input.
struct ProbLinBwdX {
    static constexpr bool A_KFAST = true, B_NFAST = true;
    static constexpr int SCRATCH = 0;
    using Cfg = CfgHead;
    GradOut dy; const float* W; ActIn in; float* dX; int B, J, I;
    struct Blk { int m0, n0, k0, k1; };
    __device__ void setup(Blk& b) const { b.m0 = (int)blockIdx.y * Cfg::BM; b.n0 = (int)blockIdx.x * Cfg::BN; b.k0 = 0; b.k1 = J; }
    __device__ void prologue(const Blk&, float*) const {}
    __device__ float loadA(const Blk&, const float*, int m, int k) const { return m < B ? dy.at(m, k) : 0.f; }
    __device__ float loadB(const Blk&, const float*, int k, int n) const { return n < I ? W[(size_t)k * I + n] : 0.f; }
    __device__ void epilogue(const Blk& b, const float*, float (&acc)[Cfg::TM][Cfg::TN], int ty, int tx, void*) const {
#pragma unroll
        for (int i = 0; i < Cfg::TM; ++i) {
            int m = b.m0 + Cfg::row_of(ty, i);
            if (m >= B) continue;
#pragma unroll
            for (int j = 0; j < Cfg::TN; ++j) {
                int n = b.n0 + Cfg::col_of(tx, j);
                if (n >= I) continue;
                float v = acc[i][j];
                if (in.scale) {
                    float z = in.scale[n] * in.U[(size_t)m * in.ld + n] + in.shift[n];
                    if (!(z > 0.f)) v = 0.f;
                }
                dX[(size_t)m * I + n] = v;
            }
        }
    }
};

// BatchNorm-over-batch backward sums.  block = 32 channels x 8 row lanes.
__global__ void k_bn_batch_bwd(const float* __restrict__ DZ, const float* __restrict__ U, int B, int C, BnState st,
                               float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ m1, float* __restrict__ m2) {
    __shared__ double sh1[8][33], sh2[8][33];
    const int tid = (int)threadIdx.x, cx = tid & 31, ry = tid >> 5;
    const int c = (int)blockIdx.x * 32 + cx;
    double s1 = 0.0, s2 = 0.0;
    if (c < C) {
        const float mu = st.mean[c], r = st.rstd[c];
        for (int b = ry; b < B; b += 8) {
            double dz = (double)DZ[(size_t)b * C + c];
            double yhat = (double)((U[(size_t)b * C + c] - mu) * r);
            s1 += dz; s2 += dz * yhat;
        }
    }
    sh1[ry][cx] = s1; sh2[ry][cx] = s2;
    __syncthreads();
    if (ry == 0 && c < C) {
        double t1 = 0.0, t2 = 0.0;
        for (int q = 0; q < 8; ++q) { t1 += sh1[q][cx]; t2 += sh2[q][cx]; }
        dgamma[c] = (float)t2; dbeta[c] = (float)t1;
        m1[c] = (float)(t1 / B); m2[c] = (float)(t2 / B);
    }
}

struct HeadArgs {
    const pgpd_head* h;
    const float* X;      // [B][1024]
    int B, out;
    bool train, is_stn;
    cudaStream_t stream;
};

inline dim3 lin_grid(int rows, int cols) { return dim3(idiv_up(cols, CfgHead::BN), idiv_up(rows, CfgHead::BM)); }

// X -> w.out  (logits, or t9 + identity)
inline void head_forward(const HeadArgs& a, HeadWs& w) {
    const pgpd_head& h = *a.h;
    cudaStream_t s = a.stream;
    const int B = a.B;
    {
        ProbLinFwd p{ActIn{a.X, nullptr, nullptr, C3}, h.fc[0].w, nullptr, w.U1, B, H1, C3, 0};
        launch_gemm<CfgHead>(p, lin_grid(B, H1), s);
    }
    if (a.train) launch(k_bn_batch_stats, grid1d(H1, 32), dim3(256), 0, s, (const float*)w.U1, B, H1, h.fc[0].b, h.bn[0], w.bn[0]);
    else launch(k_bn_eval_affine, grid1d(H1, 128), dim3(128), 0, s, H1, h.fc[0].b, h.bn[0], w.bn[0]);
    {
        ProbLinFwd p{ActIn{w.U1, w.bn[0].scale, w.bn[0].shift, H1}, h.fc[1].w, nullptr, w.U2, B, H2, H1, 0};
        launch_gemm<CfgHead>(p, lin_grid(B, H2), s);
    }
    if (a.train) launch(k_bn_batch_stats, grid1d(H2, 32), dim3(256), 0, s, (const float*)w.U2, B, H2, h.fc[1].b, h.bn[1], w.bn[1]);
    else launch(k_bn_eval_affine, grid1d(H2, 128), dim3(128), 0, s, H2, h.fc[1].b, h.bn[1], w.bn[1]);
    {
        ProbLinFwd p{ActIn{w.U2, w.bn[1].scale, w.bn[1].shift, H2}, h.fc[2].w, h.fc[2].b, w.out, B, a.out, H2, a.is_stn ? 1 : 0};
        launch_gemm<CfgHead>(p, lin_grid(B, a.out), s);
    }
}

// w.dO (gradient w.r.t. w.out) -> parameter gradients and dX [B][1024]
inline void head_backward(const HeadArgs& a, HeadWs& w, const pgpd_head_grad& g, float* dX) {
    const pgpd_head& h = *a.h;
    cudaStream_t s = a.stream;
    const int B = a.B, J3 = a.out;
    BnState none{};
    GradOut d3{w.dO, nullptr, none, nullptr, nullptr, J3, false};
    ActIn in3{w.U2, w.bn[1].scale, w.bn[1].shift, H2};
    ActIn in2{w.U1, w.bn[0].scale, w.bn[0].shift, H1};
    ActIn in1{a.X, nullptr, nullptr, C3};
    // fc3
    { ProbLinBwdW p{d3, in3, g.fc[2].dw, B, J3, H2}; launch_gemm<CfgHead>(p, lin_grid(J3, H2), s); }
    launch(k_colsum, grid1d(J3, 32), dim3(32), 0, s, (const float*)w.dO, B, J3, g.fc[2].db);
    { ProbLinBwdX p{d3, h.fc[2].w, in3, w.DZ2, B, J3, H2}; launch_gemm<CfgHead>(p, lin_grid(B, H2), s); }
    launch(k_bn_batch_bwd, grid1d(H2, 32), dim3(256), 0, s, (const float*)w.DZ2, (const float*)w.U2, B, H2, w.bn[1],
           g.bn[1].dgamma, g.bn[1].dbeta, w.m1_2, w.m2_2);
    // fc2
    GradOut d2{w.DZ2, w.U2, w.bn[1], w.m1_2, w.m2_2, H2, true};
    { ProbLinBwdW p{d2, in2, g.fc[1].dw, B, H2, H1}; launch_gemm<CfgHead>(p, lin_grid(H2, H1), s); }
    launch(k_fill, grid1d(H2, 128), dim3(128), 0, s, g.fc[1].db, (size_t)H2, 0.f);
    { ProbLinBwdX p{d2, h.fc[1].w, in2, w.DZ1, B, H2, H1}; launch_gemm<CfgHead>(p, lin_grid(B, H1), s); }
    launch(k_bn_batch_bwd, grid1d(H1, 32), dim3(256), 0, s, (const float*)w.DZ1, (const float*)w.U1, B, H1, w.bn[0],
           g.bn[0].dgamma, g.bn[0].dbeta, w.m1_1, w.m2_1);
    // fc1
    GradOut d1{w.DZ1, w.U1, w.bn[0], w.m1_1, w.m2_1, H1, true};
    { ProbLinBwdW p{d1, in1, g.fc[0].dw, B, H1, C3}; launch_gemm<CfgHead>(p, lin_grid(H1, C3), s); }
    launch(k_fill, grid1d(H1, 128), dim3(128), 0, s, g.fc[0].db, (size_t)H1, 0.f);
    { ProbLinBwdX p{d1, h.fc[0].w, in1, dX, B, H1, C3}; launch_gemm<CfgHead>(p, lin_grid(B, C3), s); }
}

}  // namespace pgpd
