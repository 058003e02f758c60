// l2bwd.cuh -- fused backward of layers 2 and 1 of a tower ("K_B"), CUDA-core parts.
//
// Reference semantics: autograd of conv2+bn2+relu and conv1+bn1+relu (+ the input transform) in
// STN3d.forward / PointNetfeat.forward, PointNetGPD/model/pointnet.py:29-30,140-146.
//
// Inputs are dz2 = d(loss)/d(bn2 output) masked by the ReLU (written by the layer-2 backward pass 1) and a1.
// With s = gamma2*rstd2, yhat2 = (y2 - mu2) rstd2, y2 = W2 a1 and the BatchNorm2-backward means m1, m2:
//
//   dy2  = s (dz2 - m1 - yhat2 m2)
//   da1  = W2^T dy2 = W2^T (s . dz2)  -  K a1  +  cvec          K = W2^T diag(s r m2) W2   (64 x 64)
//                                                                cvec = W2^T (s (r m2 mu2 - m1))
//   dW2  = sum_P dy2 a1^T = diag(s) [ C - m1 S1a^T - diag(r m2) (W2 Gram1 - mu2 S1a^T) ]
//          C = sum_P dz2 a1^T (128 x 64),  Gram1 = sum_P a1 a1^T,  S1a = sum_P a1
//
// so neither y2 nor dy2 is ever read or formed per point: one pass over (dz2, a1) yields da1 and the two
// accumulations C, Gram1.  Layer 1 is folded into the same pass: dz1 = da1 . [a1 > 0] is consumed on the spot,
//   BatchNorm1 backward sums   sum dz1, sum dz1 yhat1         (yhat1 = (a1 - beta1)/gamma1 where a1 > 0)
//   per-cloud                  H_b[k][j] = sum_{n in b} dz1[k][n] x_j[n]
// and with the per-cloud raw moments X1_b = sum x, X2_b = sum x x^T (kept from the forward)
//   G_b[k][j] = sum_n dy1[k][n] x_j[n] = s1_k ( H_b - m1'_k X1_b[j] - m2'_k r1_k ( (V_b X2_b)[k][j] - mu1_k X1_b[j] ) ),
//   V_b = W1 T_b^T,   dW1_b = G_b T_b,   dT_b[j][i] = sum_k W1[k][i] G_b[k][j]
// so dz1 is never written either.  Per point the pass reads 768 B (dz2 + a1) + 12 B (x) and writes nothing.
//
// This file: a plain fp32 CUDA-core version of the pass (PGPD_F_SIMT path and the reference every run of the tcgen05
// version in tc_kb.cuh is tested against) and the per-cloud layer-1 finalisation; the 64x64 / 64-vector precompute and
// the dW2 finalisation are parts of the fused tail kernels (tails.cuh: k_tail_ka, k_tail_kb).
#pragma once
#include "common.cuh"

namespace pgpd {

constexpr int KB_REF_NT = 32;          // points per tile of the CUDA-core pass
constexpr int KB_REF_MAX_BLOCKS = 592;

struct KbRefParams {
    const float* DZ2; const float* A1; const float* x; const float* W2; const float* scale2;
    const float* Kmat; const float* cvec; const float* gamma1; const float* beta1;
    int B, N, tiles_per_cloud, ntiles;
    float* Cpart;     // [gridDim.x][128*64]
    float* G1part;    // [gridDim.x][64*64]
    float* bnpart;    // [ntiles][2][64]
    float* Hpart;     // [ntiles][64*3]
};

// block = 256 threads, persistent over a contiguous range of 32-point tiles (tiles never straddle clouds)
__global__ void __launch_bounds__(256) k_kb_ref(KbRefParams p) {
    __shared__ float sdz[KB_REF_NT][C2];
    __shared__ float sa1[KB_REF_NT][C1];
    __shared__ float sx[3][KB_REF_NT];
    __shared__ float red[4][C1][5];
    const int tid = (int)threadIdx.x;
    const int G = (int)gridDim.x, blk = (int)blockIdx.x;
    const int t_begin = (int)(((long long)p.ntiles * blk) / G), t_end = (int)(((long long)p.ntiles * (blk + 1)) / G);
    float accC[32], accG[16];
#pragma unroll
    for (int i = 0; i < 32; ++i) accC[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) accG[i] = 0.f;
    const int k = tid & 63, q = tid >> 6;
    const float cv = p.cvec[k], be = p.beta1[k];
    const float gm = p.gamma1[k], ginv = gm != 0.f ? 1.0f / gm : 0.f;
    for (int t = t_begin; t < t_end; ++t) {
        const int b = t / p.tiles_per_cloud, tt = t % p.tiles_per_cloud;
        const int n0 = tt * KB_REF_NT;
        const int nv = (p.N - n0 < KB_REF_NT) ? p.N - n0 : KB_REF_NT;
        const size_t P0 = (size_t)b * p.N + n0;
        for (int i = tid; i < KB_REF_NT * C2; i += 256) {
            const int pp = i >> 7, c = i & 127;
            sdz[pp][c] = pp < nv ? p.DZ2[(P0 + pp) * C2 + c] : 0.f;
        }
        for (int i = tid; i < KB_REF_NT * C1; i += 256) {
            const int pp = i >> 6, kk = i & 63;
            sa1[pp][kk] = pp < nv ? p.A1[(P0 + pp) * C1 + kk] : 0.f;
        }
        if (tid < 3 * KB_REF_NT) {
            const int j = tid / KB_REF_NT, pp = tid % KB_REF_NT;
            sx[j][pp] = pp < nv ? p.x[(size_t)b * 3 * p.N + (size_t)j * p.N + n0 + pp] : 0.f;
        }
        __syncthreads();
        // ---- d a1 -> dz1 -> BatchNorm1 backward sums and H
        float s1 = 0.f, s2 = 0.f, h0 = 0.f, h1 = 0.f, h2 = 0.f;
        for (int pp = q; pp < nv; pp += 4) {
            float acc = cv;
#pragma unroll 8
            for (int c = 0; c < C2; ++c) acc = fmaf(p.W2[c * C1 + k] * p.scale2[c], sdz[pp][c], acc);
#pragma unroll 8
            for (int kk = 0; kk < C1; ++kk) acc = fmaf(-p.Kmat[k * C1 + kk], sa1[pp][kk], acc);
            const float a = sa1[pp][k];
            const float dz1 = a > 0.f ? acc : 0.f;
            const float yh = (a - be) * ginv;          // only used where dz1 != 0
            s1 += dz1;
            s2 = fmaf(dz1, yh, s2);
            h0 = fmaf(dz1, sx[0][pp], h0); h1 = fmaf(dz1, sx[1][pp], h1); h2 = fmaf(dz1, sx[2][pp], h2);
        }
        red[q][k][0] = s1; red[q][k][1] = s2; red[q][k][2] = h0; red[q][k][3] = h1; red[q][k][4] = h2;
        __syncthreads();
        if (tid < C1) {
            float r[5];
#pragma unroll
            for (int e = 0; e < 5; ++e) r[e] = ((red[0][tid][e] + red[1][tid][e]) + red[2][tid][e]) + red[3][tid][e];
            p.bnpart[((size_t)t * 2 + 0) * C1 + tid] = r[0];
            p.bnpart[((size_t)t * 2 + 1) * C1 + tid] = r[1];
            float* h = p.Hpart + (size_t)t * (C1 * 3) + tid * 3;
            h[0] = r[2]; h[1] = r[3]; h[2] = r[4];
        }
        // ---- C += dz2^T a1  (thread: channel c, 32 of the 64 a1 columns)
        {
            const int c = tid & 127, kg = tid >> 7;
            for (int pp = 0; pp < nv; ++pp) {
                const float d = sdz[pp][c];
#pragma unroll
                for (int i = 0; i < 32; ++i) accC[i] = fmaf(d, sa1[pp][kg * 32 + i], accC[i]);
            }
        }
        // ---- Gram1 += a1 a1^T  (thread: row k, 16 of the 64 columns)
        {
            for (int pp = 0; pp < nv; ++pp) {
                const float a = sa1[pp][k];
#pragma unroll
                for (int i = 0; i < 16; ++i) accG[i] = fmaf(a, sa1[pp][q * 16 + i], accG[i]);
            }
        }
        __syncthreads();
    }
    {
        const int c = tid & 127, kg = tid >> 7;
        float* o = p.Cpart + (size_t)blk * (C2 * C1) + (size_t)c * C1 + kg * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = accC[i];
        float* g = p.G1part + (size_t)blk * (C1 * C1) + (size_t)k * C1 + q * 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) g[i] = accG[i];
    }
}

// per cloud: H_b (sum of its `rpc` partial rows, fixed order) -> G_b -> dW1 partial of the cloud and d trans.
// grid = B, block = 768 = 192 (k, j) x 4 lanes.  xmom[b] = { X1 (3), X2 (3x3 row-major) } raw-coordinate moments (double).
// The last block sums the per-cloud partials of dW1 (4 lanes, clouds in order within a lane, lanes added in order:
// deterministic) and zeroes the conv1 bias gradient.
// Blocks [B, gridDim.x) are a different job that shares the launch: rows of dW2 (tails.cuh: dw2_row).
__global__ void __launch_bounds__(768) k_kb_l1(int B, const float* __restrict__ Hpart, int rpc, const double* __restrict__ xmom,
                        const float* __restrict__ trans, const float* __restrict__ W1, BnState st1, const float* __restrict__ m1,
                        const float* __restrict__ m2, float* __restrict__ dW1part, float* __restrict__ dtrans, unsigned* counter,
                        float* __restrict__ dW1, float* __restrict__ db1, Dw2Params d2) {
    if ((int)blockIdx.x >= B) { dw2_row(d2, (int)blockIdx.x - B); return; }
    __shared__ float G[C1 * 3];
    __shared__ float hs[4][C1 * 3];
    __shared__ double ds[24][C1 * 3];
    const int b = (int)blockIdx.x, tid = (int)threadIdx.x, e = tid % 192, ln = tid / 192;
    float T[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (trans)
        for (int q = 0; q < 9; ++q) T[q] = trans[(size_t)b * 9 + q];
    {
        float h = 0.f;
        const float* hp = Hpart + (size_t)b * rpc * (C1 * 3) + e;
#pragma unroll 8
        for (int r = ln; r < rpc; r += 4) h += hp[(size_t)r * (C1 * 3)];
        hs[ln][e] = h;
    }
    __syncthreads();
    if (tid < C1 * 3) {
        const int k = tid / 3, j = tid % 3;
        const float h = ((hs[0][tid] + hs[1][tid]) + hs[2][tid]) + hs[3][tid];
        const double* mo = xmom + (size_t)b * 12;
        const double w0 = W1[k * 3 + 0], w1 = W1[k * 3 + 1], w2 = W1[k * 3 + 2];
        // V[k][j'] = sum_i W1[k][i] T[j'][i]
        const double v0 = w0 * T[0] + w1 * T[1] + w2 * T[2];
        const double v1 = w0 * T[3] + w1 * T[4] + w2 * T[5];
        const double v2 = w0 * T[6] + w1 * T[7] + w2 * T[8];
        const double vx2 = v0 * mo[3 + 0 * 3 + j] + v1 * mo[3 + 1 * 3 + j] + v2 * mo[3 + 2 * 3 + j];
        const double g = (double)st1.scale[k] * ((double)h - (double)m1[k] * mo[j]
                                                 - (double)m2[k] * (double)st1.rstd[k] * (vx2 - (double)st1.mean[k] * mo[j]));
        G[tid] = (float)g;
    }
    __syncthreads();
    if (tid < C1 * 3) {
        const int kk = tid / 3, i = tid % 3;
        // dW1_b[kk][i] = sum_j T[j][i] G[kk][j]
        dW1part[(size_t)b * (C1 * 3) + tid] = T[i] * G[kk * 3 + 0] + T[3 + i] * G[kk * 3 + 1] + T[6 + i] * G[kk * 3 + 2];
    }
    if (dtrans && tid < 9) {
        const int j = tid / 3, i = tid % 3;
        float s = 0.f;
        for (int kk = 0; kk < C1; ++kk) s = fmaf(W1[kk * 3 + i], G[kk * 3 + j], s);
        dtrans[(size_t)b * 9 + tid] = s;
    }
    if (!last_block_done(counter, (unsigned)B)) return;
    // dW1 = sum over the clouds: warp w takes clouds w, w+24, ... (each lane 6 of the 192 columns, rows 4-fold unrolled: 24 loads
    // in flight per lane), then the 24 warp rows are added in order
    {
        const int warp = tid >> 5, lane = tid & 31;
        double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
        for (int bb = warp; bb < B; bb += 24) {
            const float* row = dW1part + (size_t)bb * (C1 * 3);
#pragma unroll
            for (int u = 0; u < 6; ++u) acc[u] += (double)row[lane + 32 * u];
        }
#pragma unroll
        for (int u = 0; u < 6; ++u) ds[warp][lane + 32 * u] = acc[u];
    }
    __syncthreads();
    if (tid < C1 * 3) {
        double t = 0.0;
#pragma unroll 8
        for (int w2 = 0; w2 < 24; ++w2) t += ds[w2][tid];
        dW1[tid] = (float)t;
    }
    if (tid < C1 && db1) db1[tid] = 0.f;          // bias feeding a train-mode BatchNorm: gradient is identically zero
}

}  // namespace pgpd
