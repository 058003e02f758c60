// gpd.cuh -- GPDClassifier, the paper's baseline CNN (PointNetGPD/model/gpd.py:5-31), forward and backward:
//
//   x [B][C][60][60] -> conv1 (C -> 20, 5x5, valid) -> maxpool 2x2 -> conv2 (20 -> 50, 5x5) -> maxpool 2x2 -> flatten 7200
//     -> fc1 (500) -> ReLU -> [Dropout2d when if_dropout: not supported, see pgpd.h] -> fc2 (2) -> log_softmax
//
// (no non-linearity between the convolutions and the pools, exactly as the reference writes it).  SURVEY.md section 8f row 4:
// a different network from the PointNet hot path, outside north_star's metric; provided for the completeness of the `model`
// package (main_1v_gpd.py / main_fullv_gpd.py construct it).  fp32 CUDA-core kernels; fc1 (7200 -> 500, 85 % of the
// parameters) reuses the split-K GEMMs of head.cuh (tcgen05 when available).  Convolution + pool are fused (the pooled value
// and the position of the maximum inside the 2x2 window are kept; the un-pooled activation is never stored), and the
// backward routes gradients through the kept positions: gather formulations only, fixed summation orders, no atomics.
#pragma once
#include "common.cuh"
#include "head.cuh"

namespace pgpd {

constexpr int GPD_IN = 60, GPD_K = 5;
constexpr int GPD_C1 = 20, GPD_P1 = 28;        // conv1: 56 x 56 -> pooled 28 x 28
constexpr int GPD_C2 = 50, GPD_P2 = 12;        // conv2: 24 x 24 -> pooled 12 x 12
constexpr int GPD_FLAT = GPD_C2 * GPD_P2 * GPD_P2;   // 7200
constexpr int GPD_H = 500, GPD_OUT = 2;

struct GpdWs {
    float* P1; unsigned char* I1;              // [B][20][28][28] pooled conv1 output, position of the max (0..3 = dy*2+dx)
    float* P2; unsigned char* I2;              // [B][50][12][12]
    float* U; float* Hh;                       // [B][500] fc1 pre-activation (incl. bias), relu(U)
    float* logits;                             // [B][2]
    float* logp;                               // [B][2] kept for the backward
    float* partA; float* partB;                // split-K partials
    unsigned* amax;                            // [64 + 64] partial max |x| for the tcgen05 GEMM operand scales
    // backward
    float* dO; float* dH; float* dP2; float* dP1;
    size_t bytes;
};

inline void plan_gpd(void* base, int B, bool backward, GpdWs& w) {
    Carver c(base);
    w.P1 = c.take<float>((size_t)B * GPD_C1 * GPD_P1 * GPD_P1);
    w.I1 = c.take<unsigned char>((size_t)B * GPD_C1 * GPD_P1 * GPD_P1);
    w.P2 = c.take<float>((size_t)B * GPD_FLAT);
    w.I2 = c.take<unsigned char>((size_t)B * GPD_FLAT);
    w.U = c.take<float>((size_t)B * GPD_H);
    w.Hh = c.take<float>((size_t)B * GPD_H);
    w.logits = c.take<float>((size_t)B * GPD_OUT);
    w.logp = c.take<float>((size_t)B * GPD_OUT);
    w.partA = c.take<float>(std::max(HEAD_PART_ELEMS, (size_t)GPD_H * GPD_FLAT));     // forward fc1 partials / dW1 (one slice)
    w.partB = c.take<float>(std::max(HEAD_PART_ELEMS, (size_t)B * GPD_FLAT));         // dP2 (one slice)
    w.amax = c.take<unsigned>(256);
    if (backward) {
        w.dO = c.take<float>((size_t)B * GPD_OUT);
        w.dH = c.take<float>((size_t)B * GPD_H);
        w.dP2 = c.take<float>((size_t)B * GPD_FLAT);
        w.dP1 = c.take<float>((size_t)B * GPD_C1 * GPD_P1 * GPD_P1);
    }
    w.bytes = (c.off + 255) & ~(size_t)255;
}

// ---- convolution 5x5 (valid) + bias + 2x2 max-pool, fused.  thread = one pooled output (b, oc, py, px) ---------------------------
// in [B][CIN][HIN][HIN] -> out [B][COUT][HP][HP] with HP = (HIN - 4) / 2; pos = dy*2+dx of the FIRST maximum in row-major window
// order (torch's MaxPool2d keeps the first maximum).
__global__ void __launch_bounds__(256) k_gpd_conv_pool(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
                                                       int B, int CIN, int HIN, int COUT, int HP, float* __restrict__ out,
                                                       unsigned char* __restrict__ pos) {
    const size_t total = (size_t)B * COUT * HP * HP;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int px = (int)(i % HP), py = (int)((i / HP) % HP), oc = (int)((i / ((size_t)HP * HP)) % COUT), b = (int)(i / ((size_t)HP * HP * COUT));
    float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
    const float* wp = W + (size_t)oc * CIN * 25;
    for (int ic = 0; ic < CIN; ++ic) {
        const float* ip = in + (((size_t)b * CIN + ic) * HIN + 2 * py) * HIN + 2 * px;
        // 6 x 6 input patch of this channel: conv outputs (0,0), (0,1), (1,0), (1,1) of the pooling window
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            float v[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) v[q] = ip[(size_t)r * HIN + q];
#pragma unroll
            for (int ky = 0; ky < 5; ++ky) {
                const int dy = r - ky;                      // conv row offset inside the window that uses input row r with kernel row ky
                if (dy != 0 && dy != 1) continue;
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) {
                    const float wv = wp[ic * 25 + ky * 5 + kx];
                    if (dy == 0) { a00 = fmaf(wv, v[kx], a00); a01 = fmaf(wv, v[kx + 1], a01); }
                    else { a10 = fmaf(wv, v[kx], a10); a11 = fmaf(wv, v[kx + 1], a11); }
                }
            }
        }
    }
    const float bb = bias[oc];
    a00 += bb; a01 += bb; a10 += bb; a11 += bb;
    float m = a00; int p = 0;
    if (a01 > m || a01 != a01) { if (!(m != m)) { m = a01; p = 1; } }
    if (a10 > m || a10 != a10) { if (!(m != m)) { m = a10; p = 2; } }
    if (a11 > m || a11 != a11) { if (!(m != m)) { m = a11; p = 3; } }
    out[i] = m;
    pos[i] = (unsigned char)p;
}

// U = sum of split-K partials + bias; H = relu(U)
__global__ void k_gpd_fc1_finish(const float* __restrict__ part, int nsl, size_t n, const float* __restrict__ bias, float* __restrict__ U,
                                 float* __restrict__ H) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = part[i];
#pragma unroll 8
    for (int z = 1; z < nsl; ++z) v += part[(size_t)z * n + i];
    v += bias[i % GPD_H];
    U[i] = v;
    H[i] = relu_nan(v);
}

// logits = H W2^T + b2, log_softmax.  one warp per sample
__global__ void __launch_bounds__(256) k_gpd_fc2_out(const float* __restrict__ H, const float* __restrict__ W2, const float* __restrict__ b2, int B,
                                                     float* __restrict__ logits, float* __restrict__ logp, float* __restrict__ user) {
    const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
    const int b = (int)blockIdx.x * 8 + warp;
    if (b >= B) return;
    float s0 = 0.f, s1 = 0.f;
    for (int k = lane; k < GPD_H; k += 32) {
        const float h = H[(size_t)b * GPD_H + k];
        s0 = fmaf(h, W2[k], s0);
        s1 = fmaf(h, W2[GPD_H + k], s1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
    if (lane == 0) {
        s0 += b2[0]; s1 += b2[1];
        const float mx = (s0 > s1 || s0 != s0) ? s0 : s1;
        const float lse = mx + logf(expf(s0 - mx) + expf(s1 - mx));
        logits[(size_t)b * 2] = s0; logits[(size_t)b * 2 + 1] = s1;
        logp[(size_t)b * 2] = s0 - lse; logp[(size_t)b * 2 + 1] = s1 - lse;
        if (user) { user[(size_t)b * 2] = s0 - lse; user[(size_t)b * 2 + 1] = s1 - lse; }
    }
}

// ---- backward ------------------------------------------------------------------------------------------------------------------
// blocks [0, 2): dW2[j][k] = sum_b dO[b][j] H[b][k], db2[j] = sum_b dO[b][j] (thread = k, clouds in order);
// blocks [2, ..): dU[b][k] = (sum_j dO[b][j] W2[j][k]) [U > 0], 2 samples per block of 1024 (thread = k of a sample)
__global__ void __launch_bounds__(1024) k_gpd_fc2_bwd(const float* __restrict__ logp, const float* __restrict__ dlogp, const float* __restrict__ H,
                                                      const float* __restrict__ U, const float* __restrict__ W2, int B,
                                                      float* __restrict__ dW2, float* __restrict__ db2, float* __restrict__ dU) {
    const int tid = (int)threadIdx.x;
    if ((int)blockIdx.x < 2) {
        const int j = (int)blockIdx.x;
        if (tid >= GPD_H && tid != 1023) return;
        float acc = 0.f; double bs = 0.0;
        for (int b = 0; b < B; ++b) {
            const float sdl = dlogp[(size_t)b * 2] + dlogp[(size_t)b * 2 + 1];
            const float d = dlogp[(size_t)b * 2 + j] - expf(logp[(size_t)b * 2 + j]) * sdl;   // d logits = d logp - softmax * sum d logp
            if (tid < GPD_H) acc = fmaf(d, H[(size_t)b * GPD_H + tid], acc); else bs += (double)d;
        }
        if (tid < GPD_H) dW2[(size_t)j * GPD_H + tid] = acc; else db2[j] = (float)bs;
        return;
    }
    const int b = ((int)blockIdx.x - 2) * 2 + (tid >> 9), k = tid & 511;
    if (b >= B || k >= GPD_H) return;
    const float sdl = dlogp[(size_t)b * 2] + dlogp[(size_t)b * 2 + 1];
    const float d0 = dlogp[(size_t)b * 2] - expf(logp[(size_t)b * 2]) * sdl, d1 = dlogp[(size_t)b * 2 + 1] - expf(logp[(size_t)b * 2 + 1]) * sdl;
    const float v = d0 * W2[k] + d1 * W2[GPD_H + k];
    dU[(size_t)b * GPD_H + k] = U[(size_t)b * GPD_H + k] > 0.f ? v : 0.f;
}

// out1 = sum of partials (dW fc1 [500][7200]); out2 = sum of partials (dP2 [B][7200]); db1[k] = sum_b dU[b][k] (last 2 blocks)
__global__ void __launch_bounds__(1024) k_gpd_fc1_bwd_finish(const float* __restrict__ part1, int nsl1, size_t n1, float* __restrict__ out1,
                                                             const float* __restrict__ part2, int nsl2, size_t n2, float* __restrict__ out2,
                                                             const float* __restrict__ dU, int B, float* __restrict__ db1) {
    const size_t nb1 = (n1 + 1023) / 1024, nb2 = (n2 + 1023) / 1024;
    size_t blk = blockIdx.x;
    if (blk < nb1) {
        const size_t i = blk * 1024 + threadIdx.x;
        if (i < n1) {
            float v = part1[i];
#pragma unroll 8
            for (int z = 1; z < nsl1; ++z) v += part1[(size_t)z * n1 + i];
            out1[i] = v;
        }
        return;
    }
    blk -= nb1;
    if (blk < nb2) {
        const size_t i = blk * 1024 + threadIdx.x;
        if (i < n2) {
            float v = part2[i];
#pragma unroll 8
            for (int z = 1; z < nsl2; ++z) v += part2[(size_t)z * n2 + i];
            out2[i] = v;
        }
        return;
    }
    const int k = (int)threadIdx.x;
    if (k < GPD_H) {
        double s = 0.0;
        for (int b = 0; b < B; ++b) s += (double)dU[(size_t)b * GPD_H + k];
        db1[k] = (float)s;
    }
}

// weight / bias gradient of a conv + pool stage.  Only the conv output at the kept position of each pooling window receives gradient:
//   dW[oc][ic][ky][kx] = sum_{b, cell} dP[b][oc][cell] * in[b][ic][2 py + dy + ky][2 px + dx + kx],   db[oc] = sum dP[b][oc][cell]
// grid = (COUT, CIN), block = 256: thread t sums its (b, cell) pairs (stride 256) into 25 registers; lanes are added in a fixed order.
__global__ void __launch_bounds__(256) k_gpd_conv_wgrad(const float* __restrict__ dP, const unsigned char* __restrict__ pos, const float* __restrict__ in,
                                                        int B, int CIN, int HIN, int COUT, int HP, float* __restrict__ dW, float* __restrict__ db) {
    __shared__ float sh[256][26];
    const int oc = (int)blockIdx.x, ic = (int)blockIdx.y, tid = (int)threadIdx.x;
    float acc[25];
#pragma unroll
    for (int q = 0; q < 25; ++q) acc[q] = 0.f;
    float bsum = 0.f;
    const int cells = HP * HP;
    const size_t total = (size_t)B * cells;
    for (size_t e = tid; e < total; e += 256) {
        const int b = (int)(e / cells), cell = (int)(e % cells), py = cell / HP, px = cell % HP;
        const size_t o = ((size_t)b * COUT + oc) * cells + cell;
        const float g = dP[o];
        const int p = pos[o];
        bsum += g;
        const float* ip = in + (((size_t)b * CIN + ic) * HIN + 2 * py + (p >> 1)) * HIN + 2 * px + (p & 1);
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) acc[ky * 5 + kx] = fmaf(g, ip[(size_t)ky * HIN + kx], acc[ky * 5 + kx]);
    }
#pragma unroll
    for (int q = 0; q < 25; ++q) sh[tid][q] = acc[q];
    sh[tid][25] = bsum;
    __syncthreads();
    if (tid < 26) {
        double s = 0.0;
        for (int l = 0; l < 256; ++l) s += (double)sh[l][tid];
        if (tid < 25) dW[((size_t)oc * CIN + ic) * 25 + tid] = (float)s;
        else if (ic == 0) db[oc] = (float)s;
    }
}

// input gradient of the conv2 + pool2 stage (gather form): dIn[b][ic][y][x] = sum_oc sum over the conv positions (cy, cx) with
// y - cy = ky, x - cx = kx in [0, 5) that ARE the kept position of their pooling window:  dP[b][oc][cy/2][cx/2] * W[oc][ic][ky][kx].
// thread = one input element.
__global__ void __launch_bounds__(256) k_gpd_conv_dgrad(const float* __restrict__ dP, const unsigned char* __restrict__ pos, const float* __restrict__ W,
                                                        int B, int CIN, int HIN, int COUT, int HP, float* __restrict__ dIn) {
    const size_t total = (size_t)B * CIN * HIN * HIN;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % HIN), y = (int)((i / HIN) % HIN), ic = (int)((i / ((size_t)HIN * HIN)) % CIN), b = (int)(i / ((size_t)HIN * HIN * CIN));
    // conv positions (cy, cx) with y - 4 <= cy <= y lie in the pooling windows py in [ceil((y-5)/2), y/2]: at most 3 x 3 windows,
    // each with ONE kept position
    const int py0 = (y - 4) > 0 ? ((y - 4) >> 1) : 0, py1 = (y >> 1) < HP ? (y >> 1) : HP - 1;
    const int px0 = (x - 4) > 0 ? ((x - 4) >> 1) : 0, px1 = (x >> 1) < HP ? (x >> 1) : HP - 1;
    float acc = 0.f;
    for (int oc = 0; oc < COUT; ++oc) {
        const float* wp = W + ((size_t)oc * CIN + ic) * 25;
        const size_t ob = ((size_t)b * COUT + oc) * HP * HP;
        for (int py = py0; py <= py1; ++py)
            for (int px = px0; px <= px1; ++px) {
                const size_t o = ob + (size_t)py * HP + px;
                const int p = pos[o];
                const int ky = y - (2 * py + (p >> 1)), kx = x - (2 * px + (p & 1));
                if (ky >= 0 && ky < 5 && kx >= 0 && kx < 5) acc = fmaf(dP[o], wp[ky * 5 + kx], acc);
            }
    }
    dIn[i] = acc;
}

struct GpdArgs { const pgpd_gpd* m; const float* x; int B, C; cudaStream_t stream; bool use_tc; };

inline void gpd_forward(const GpdArgs& a, GpdWs& w, float* logp_user) {
    cudaStream_t s = a.stream;
    const int B = a.B;
    const pgpd_gpd& m = *a.m;
    launch(k_gpd_conv_pool, grid1d((size_t)B * GPD_C1 * GPD_P1 * GPD_P1, 256), dim3(256), 0, s, a.x, m.conv1.w, m.conv1.b, B, a.C, GPD_IN, GPD_C1,
           GPD_P1, w.P1, w.I1);
    launch(k_gpd_conv_pool, grid1d((size_t)B * GPD_FLAT, 256), dim3(256), 0, s, (const float*)w.P1, m.conv2.w, m.conv2.b, B, GPD_C1, GPD_P1, GPD_C2,
           GPD_P2, w.P2, w.I2);
    int nsl = 0;
#ifndef PGPD_EMU
    if (a.use_tc) {
        // operand scales: measured max |x| of both operands (the pooled conv features are not O(1) by construction)
        launch(tc::k_absmax2, dim3(64, 2), dim3(256), 0, s, (const float*)w.P2, (size_t)B * GPD_FLAT, m.fc1.w, (size_t)GPD_H * GPD_FLAT, w.amax);
        nsl = run_gemm_tc(tc::GemmOp{w.P2, GPD_FLAT, 0, w.amax, 64, 1.f}, tc::GemmOp{m.fc1.w, GPD_FLAT, 0, w.amax + 64, 64, 1.f}, B, GPD_H, GPD_FLAT,
                          w.partA, 8, s);
    } else
#endif
    nsl = run_plain(ProbPlain<true, false>{w.P2, m.fc1.w, w.partA, B, GPD_H, GPD_FLAT, (size_t)GPD_FLAT, 1, 1, (size_t)GPD_FLAT, 0}, s, 8);
    launch(k_gpd_fc1_finish, grid1d((size_t)B * GPD_H, 256), dim3(256), 0, s, (const float*)w.partA, nsl, (size_t)B * GPD_H, m.fc1.b, w.U, w.Hh);
    launch(k_gpd_fc2_out, grid1d(B, 8), dim3(256), 0, s, (const float*)w.Hh, m.fc2.w, m.fc2.b, B, w.logits, w.logp, logp_user);
}

inline void gpd_backward(const GpdArgs& a, GpdWs& w, const pgpd_gpd_grad& g, const float* dlogp) {
    cudaStream_t s = a.stream;
    const int B = a.B;
    const pgpd_gpd& m = *a.m;
    launch(k_gpd_fc2_bwd, dim3(2 + idiv_up(B, 2)), dim3(1024), 0, s, (const float*)w.logp, dlogp, (const float*)w.Hh, (const float*)w.U, m.fc2.w, B,
           g.fc2.dw, g.fc2.db, w.dH);
    // fc1: dW1 = dU^T P2 (K = B), dP2 = dU W1 (K = 500)
    int nslW = 0, nslX = 0;
#ifndef PGPD_EMU
    if (a.use_tc) {
        launch(tc::k_absmax2, dim3(64, 2), dim3(256), 0, s, (const float*)w.dH, (size_t)B * GPD_H, (const float*)nullptr, (size_t)0, w.amax + 128);
        nslW = run_gemm_tc(tc::GemmOp{w.dH, GPD_H, 1, w.amax + 128, 64, 1.f}, tc::GemmOp{w.P2, GPD_FLAT, 1, w.amax, 64, 1.f}, GPD_H, GPD_FLAT, B,
                           w.partA, 0, s);
        nslX = run_gemm_tc(tc::GemmOp{w.dH, GPD_H, 0, w.amax + 128, 64, 1.f}, tc::GemmOp{m.fc1.w, GPD_FLAT, 1, w.amax + 64, 64, 1.f}, B, GPD_FLAT, GPD_H,
                           w.partB, 0, s);
    } else
#endif
    {
        nslW = run_plain(ProbPlain<false, true>{w.dH, w.P2, w.partA, GPD_H, GPD_FLAT, B, 1, (size_t)GPD_H, (size_t)GPD_FLAT, 1, 0}, s, 0);
        nslX = run_plain(ProbPlain<true, true>{w.dH, m.fc1.w, w.partB, B, GPD_FLAT, GPD_H, (size_t)GPD_H, 1, (size_t)GPD_FLAT, 1, 0}, s, 0);
    }
    {
        const size_t n1 = (size_t)GPD_H * GPD_FLAT, n2 = (size_t)B * GPD_FLAT;
        launch(k_gpd_fc1_bwd_finish, dim3((unsigned)((n1 + 1023) / 1024 + (n2 + 1023) / 1024 + 1)), dim3(1024), 0, s, (const float*)w.partA, nslW, n1,
               g.fc1.dw, (const float*)w.partB, nslX, n2, w.dP2, (const float*)w.dH, B, g.fc1.db);
    }
    // conv2 + pool2
    launch(k_gpd_conv_wgrad, dim3(GPD_C2, GPD_C1), dim3(256), 0, s, (const float*)w.dP2, (const unsigned char*)w.I2, (const float*)w.P1, B, GPD_C1, GPD_P1,
           GPD_C2, GPD_P2, g.conv2.dw, g.conv2.db);
    launch(k_gpd_conv_dgrad, grid1d((size_t)B * GPD_C1 * GPD_P1 * GPD_P1, 256), dim3(256), 0, s, (const float*)w.dP2, (const unsigned char*)w.I2,
           m.conv2.w, B, GPD_C1, GPD_P1, GPD_C2, GPD_P2, w.dP1);
    // conv1 + pool1 (the input gets no gradient)
    launch(k_gpd_conv_wgrad, dim3(GPD_C1, a.C), dim3(256), 0, s, (const float*)w.dP1, (const unsigned char*)w.I1, a.x, B, a.C, GPD_IN, GPD_C1, GPD_P1,
           g.conv1.dw, g.conv1.db);
}

}  // namespace pgpd
