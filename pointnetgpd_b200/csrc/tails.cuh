// tails.cuh -- the small kernels around the big streaming / tensor-core kernels of a tower, FUSED.
//
// Round 1 ran every reduction, BatchNorm finalisation, mean propagation and weight pre-pack as its own launch
// (~150 launches of 2-10 us per training step, a third of the step).  Here each chain between two big kernels is ONE
// launch: blocks write partial results, the last block to finish (common.cuh: last_block_done) reduces them in a fixed
// order and finalises, or the work is split by output channel so that no cross-block step exists at all.
// Everything stays deterministic (fixed summation orders, no floating-point atomics).
//
// Reference semantics of what is finalised here: nn.BatchNorm1d in training mode (batch statistics, running-stat
// update; pointnet.py:21-25,130-132), MaxPool1d over the points (pointnet.py:32,148) -- see tower.cuh.
#pragma once
#include "common.cuh"
#ifndef PGPD_EMU
#include "tc_ptx.cuh"
#endif

namespace pgpd {

constexpr int PRE_THREADS = 256;
constexpr int PRE_SAMPLE = 256;      // points of the pilot estimate of mean(a1): one per thread of the finalising block
static_assert(PRE_SAMPLE == PRE_THREADS, "k_tower_pre: one sample point per thread");
constexpr int A1_CHUNK = 64;        // points per staging chunk of k_a1
constexpr int A1_CPB = 8;           // chunks per block: one partial row of the a1 sums per 512 points

// ================================================================================================
// F1: k_tower_pre -- everything a tower forward needs before its first per-point kernel
//   role MOMENTS (train): one block per cloud: first and second moments of the TRANSFORMED points x' = T^T x in double
//                 (mom[b] = {sum x'_i (3), sum x'_i x'_i2 (3x3)}), raw-coordinate moments kept for the backward; the last
//                 of these blocks sums the clouds in a fixed order and finalises BatchNorm1 analytically:
//                 u1 = W1 x'  =>  mean = W1 m,  var_c = w_c^T Cov w_c.
//   role AFFINE  (eval): one block: folded scale/shift of the three BatchNorms from the running statistics.
//   role W2IMG / W3IMG (tensor-core path): hi/lo fp16 operand images of conv2.weight / conv3.weight (swizzled, rows
//                 scaled by powers of two; layout as consumed by tc_kf.cuh / tc_l3.cuh), sign(gamma3) folded into W3.
//   role SIGN    (CUDA-core path): sgn[c] = sign(gamma3[c]).
// ================================================================================================
struct PreParams {
    const float* x; const float* trans; int B, N;
    double* moments; double* rawmom;          // [B][12]; rawmom may be null
    pgpd_lin conv[3]; pgpd_bn bn[3]; BnState st[3];
    int train; double count;
    unsigned* counter;                        // ticket of the MOMENTS role
    unsigned* bad;                            // [B+1] flags; [B] = "a weight is not finite"
    int n_mom, n_w2, n_w3;                    // blocks per role (n_w2 = n_w3 = 0 on the CUDA-core path: SIGN role instead)
    void* wimg2; float* inv2;                 // W2 image (32 KB) + per-row inverse scale [128]
    void* wimg3; float* sgn;                  // W3 image (512 KB) + inv / sign [1024]
    int act_shift;
    float* centre2;                           // tensor-core train path: [128] pilot estimate of mean(u2) = W2 mean(a1) over a sample of
                                              // points (the layer-2 kernel centres its squares on it and also sums u2 exactly); or null
};

#ifndef PGPD_EMU
namespace tc {
// one row of a K-major hi/lo operand image: row r of [kb][part][128 rows][64 halves], 128-byte swizzle.
// w: this thread's element k of the row (already sign-adjusted), mx: the row's max |w|.  Returns the exponent e used.
__device__ __forceinline__ int prepack_elem(float w, float mx, int r, int k, __half* img, size_t img_base_halves,
                                            size_t block_halves = 8192) {     // 8192: 128-row blocks; 4096: 64-row blocks
    int ex = 0;
    if (mx > 0.f) frexpf(mx, &ex);               // mx in [2^(ex-1), 2^ex)
    const int e = (mx > 0.f && mx < INFINITY) ? 14 - ex : 0;     // mx * 2^e in [2^13, 2^14)
    const float ws = ldexpf(w, e);
    const __half hi = __float2half_rn(ws);
    const __half lo = __float2half_rn(ws - __half2float(hi));
    const int kb = k >> 6, j = k & 63, chunk = j >> 3, within = j & 7;
    const size_t base = img_base_halves + (size_t)(kb * 2) * block_halves;
    const size_t off = (size_t)r * 64 + (size_t)((chunk ^ (r & 7)) << 3) + within;
    img[base + off] = hi;
    img[base + block_halves + off] = lo;
    return e;
}
}  // namespace tc
#endif

__global__ void __launch_bounds__(PRE_THREADS) k_tower_pre(PreParams p) {
    __shared__ double sh[PRE_THREADS];
    __shared__ double raw[12];
    __shared__ float redf[PRE_THREADS];
    const int tid = (int)threadIdx.x;
    int blk = (int)blockIdx.x;
    if (blk < p.n_mom) {
        if (!p.train) {
            // ---- AFFINE (eval): y = gamma*(u + b - rm)/sqrt(rv+eps) + beta for the three layers
            for (int L = 0; L < 3; ++L) {
                const int C = L == 0 ? C1 : (L == 1 ? C2 : C3);
                for (int c = tid; c < C; c += PRE_THREADS) {
                    const pgpd_bn& bn = p.bn[L];
                    const float rstd = 1.0f / sqrtf(bn.running_var[c] + BN_EPS);
                    const float sc = bn.gamma[c] * rstd;
                    const float b = p.conv[L].b ? p.conv[L].b[c] : 0.f;
                    p.st[L].mean[c] = bn.running_mean[c] - b;
                    p.st[L].rstd[c] = rstd;
                    p.st[L].scale[c] = sc;
                    p.st[L].shift[c] = bn.beta[c] + sc * (b - bn.running_mean[c]);
                }
            }
            return;
        }
        // ---- MOMENTS (train): cloud b = blk
        const int b = blk;
        const float* xb = p.x + (size_t)b * 3 * p.N;
        double acc[9];
        for (int q = 0; q < 9; ++q) acc[q] = 0.0;
        for (int n = tid; n < p.N; n += PRE_THREADS) {
            const double p0 = xb[n], p1 = xb[p.N + n], p2 = xb[2 * p.N + n];
            acc[0] += p0; acc[1] += p1; acc[2] += p2;
            acc[3] += p0 * p0; acc[4] += p0 * p1; acc[5] += p0 * p2;
            acc[6] += p1 * p1; acc[7] += p1 * p2; acc[8] += p2 * p2;
        }
        for (int q = 0; q < 9; ++q) {
            sh[tid] = acc[q];
            __syncthreads();
            for (int s = PRE_THREADS / 2; s > 0; s >>= 1) {
                if (tid < s) sh[tid] += sh[tid + s];
                __syncthreads();
            }
            if (tid == 0) raw[q] = sh[0];
            __syncthreads();
        }
        if (tid < 12) {
            double T[3][3];
            for (int j = 0; j < 3; ++j)
                for (int i = 0; i < 3; ++i) T[j][i] = p.trans ? (double)p.trans[(size_t)b * 9 + j * 3 + i] : (i == j ? 1.0 : 0.0);
            const double s1[3] = {raw[0], raw[1], raw[2]};
            const double X[3][3] = {{raw[3], raw[4], raw[5]}, {raw[4], raw[6], raw[7]}, {raw[5], raw[7], raw[8]}};
            double v = 0.0;
            if (tid < 3) {
                for (int j = 0; j < 3; ++j) v += T[j][tid] * s1[j];
            } else {
                const int i = (tid - 3) / 3, i2 = (tid - 3) % 3;
                for (int j = 0; j < 3; ++j)
                    for (int j2 = 0; j2 < 3; ++j2) v += T[j][i] * X[j][j2] * T[j2][i2];
            }
            p.moments[(size_t)b * 12 + tid] = v;
            if (p.rawmom) p.rawmom[(size_t)b * 12 + tid] = tid < 3 ? s1[tid] : X[(tid - 3) / 3][(tid - 3) % 3];
        }
        if (!last_block_done(p.counter, (unsigned)p.n_mom)) return;
        // ---- last cloud block: sum the clouds (12 entries x 21 lanes, lanes added in order) and finalise BatchNorm1
        {
            const int e = tid % 12, ln = tid / 12;          // 252 active threads
            double s = 0.0;
            if (ln < 21)
                for (int bb = ln; bb < p.B; bb += 21) s += p.moments[(size_t)bb * 12 + e];
            sh[tid] = s;
            __syncthreads();
            if (tid < 12) {
                double t = 0.0;
                for (int l2 = 0; l2 < 21; ++l2) t += sh[l2 * 12 + tid];
                raw[tid] = t / p.count;
            }
            __syncthreads();
            if (tid < C1) {
                const int c = tid;
                const double w[3] = {p.conv[0].w[c * 3 + 0], p.conv[0].w[c * 3 + 1], p.conv[0].w[c * 3 + 2]};
                const double mean_u = w[0] * raw[0] + w[1] * raw[1] + w[2] * raw[2];
                double var = 0;
                for (int i = 0; i < 3; ++i)
                    for (int i2 = 0; i2 < 3; ++i2) var += w[i] * (raw[3 + i * 3 + i2] - raw[i] * raw[i2]) * w[i2];
                bn_finalize_train(c, mean_u, var, p.count, p.conv[0].b, p.bn[0], p.st[0]);
            }
            if (p.centre2) {
                // pilot mean of a1 over PRE_SAMPLE points spread over the batch, then W2 . mean.  All global loads (the sample
                // points, this thread's half row of W2) are issued up front: one latency epoch, not a chain.
                const int c2 = tid & 127, hk = tid >> 7;            // matvec: channel x half of k
                float4 wv[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) wv[e] = *reinterpret_cast<const float4*>(p.conv[1].w + (size_t)c2 * C1 + hk * 32 + e * 4);
                const size_t M = (size_t)p.B * p.N;
                const size_t step = M > PRE_SAMPLE ? M / PRE_SAMPLE : 1;
                const int ns = (int)(M < PRE_SAMPLE ? M : PRE_SAMPLE);
                __shared__ float s_pt[3][PRE_SAMPLE];
                __shared__ float s_m[C1];
                {
                    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
                    if (tid < ns) {
                        const size_t P = (size_t)tid * step;
                        const int b = (int)(P / p.N), n = (int)(P % p.N);
                        const float* xb = p.x + (size_t)b * 3 * p.N + n;
                        const float p0 = xb[0], p1 = xb[p.N], p2 = xb[2 * (size_t)p.N];
                        t0 = p0; t1 = p1; t2 = p2;
                        if (p.trans) {
                            const float* T = p.trans + (size_t)b * 9;
                            t0 = T[0] * p0 + T[3] * p1 + T[6] * p2;
                            t1 = T[1] * p0 + T[4] * p1 + T[7] * p2;
                            t2 = T[2] * p0 + T[5] * p1 + T[8] * p2;
                        }
                    }
                    s_pt[0][tid] = t0; s_pt[1][tid] = t1; s_pt[2][tid] = t2;
                }
                __syncthreads();                                     // also: BatchNorm1 scale / shift written above are visible
                const int k = tid & 63, ln = tid >> 6;               // thread = channel x 4 point lanes
                const float w0 = p.conv[0].w[k * 3 + 0], w1 = p.conv[0].w[k * 3 + 1], w2 = p.conv[0].w[k * 3 + 2];
                const float sc = p.st[0].scale[k], sf = p.st[0].shift[k];
                float acc = 0.f;
#pragma unroll 8
                for (int i = ln; i < ns; i += 4)
                    acc += relu_nan(sc * (w0 * s_pt[0][i] + w1 * s_pt[1][i] + w2 * s_pt[2][i]) + sf);
                redf[tid] = acc;
                __syncthreads();
                if (tid < C1) s_m[tid] = (((redf[tid] + redf[tid + 64]) + redf[tid + 128]) + redf[tid + 192]) / (float)ns;
                __syncthreads();
                float s2 = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float* mm = s_m + hk * 32 + e * 4;
                    s2 += wv[e].x * mm[0] + wv[e].y * mm[1] + wv[e].z * mm[2] + wv[e].w * mm[3];
                }
                redf[tid] = s2;
                __syncthreads();
                if (tid < C2) p.centre2[tid] = redf[tid] + redf[tid + 128];
            }
        }
        return;
    }
    blk -= p.n_mom;
#ifndef PGPD_EMU
    if (blk < p.n_w2) {
        // ---- W2IMG: 4 rows of conv2.weight [128][64] per block (KD = 64: one k-block)
        const int r = blk * 4 + (tid >> 6), k = tid & 63;
        const float w = p.conv[1].w[(size_t)r * C1 + k];
        redf[tid] = fabsf(w);
        __syncthreads();
        for (int s = 32; s > 0; s >>= 1) {
            if (k < s) redf[tid] = fmaxf(redf[tid], redf[tid + s]);
            __syncthreads();
        }
        const float mx = redf[tid & ~63];
        const int e = tc::prepack_elem(w, mx, r, k, (__half*)p.wimg2, 0);
        if (k == 0) p.inv2[r] = ldexpf(1.f, -(e + p.act_shift));
        return;
    }
    blk -= p.n_w2;
    if (blk < p.n_w3) {
        // ---- W3IMG: 2 rows of conv3.weight [1024][128] per block; image block (mt = c/128): [mt][kb][part][128 rows][64 halves]
        const int c = blk * 2 + (tid >> 7), k = tid & 127;
        const float w = p.conv[2].w[(size_t)c * C2 + k];
        redf[tid] = fabsf(w);
        __syncthreads();
        for (int s = 64; s > 0; s >>= 1) {
            if (k < s) redf[tid] = fmaxf(redf[tid], redf[tid + s]);
            __syncthreads();
        }
        const float mx = redf[tid & ~127];
        const float sg = p.bn[2].gamma[c] >= 0.f ? 1.f : -1.f;
        const int mt = c >> 7, r = c & 127;
        const int e = tc::prepack_elem(w * sg, mx, r, k, (__half*)p.wimg3, (size_t)(mt * 4) * 8192);
        if (k == 0) {
            p.sgn[c] = sg * ldexpf(1.f, -(e + p.act_shift));
            if (!(mx < INFINITY)) p.bad[p.B] = 1u;           // NaN / Inf weight: every pooled value of this tower is poisoned
        }
        return;
    }
#else
    (void)redf;
#endif
    // ---- SIGN (CUDA-core path): the remaining blocks
    {
        const int c = blk * PRE_THREADS + tid;
        if (c < C3) p.sgn[c] = p.bn[2].gamma[c] >= 0.f ? 1.f : -1.f;
    }
}

// ================================================================================================
// k_a1: a1 = relu(scale1 * (W1 T^T x) + shift1), stored [M][64]  (+ train: sum of a1 -> S1a, mean(u2) = W2 mean(a1))
// grid = (ceil(N / 512), clouds), block = 512 threads = 64 channels x 8 point slots; the block's points are staged
// (transformed) in shared memory once.  Train mode: every block writes one partial row of the a1 sums; the LAST block sums the
// rows (64 columns x 8 lanes, fixed order, double) and propagates the mean through conv2: mean(u2) = W2 mean(a1) -- the
// centre of layer 2's sum of squares.
// `limit`: activations above it (or NaN) flag the cloud in bad[] (tensor-core path: the fp16 operand range).
// Rule for every tail in this file: a serial loop over partial rows is a chain of L2 round trips (~0.4 us each), so rows are
// spread over many lanes and the loops are unrolled 8-fold (8 loads in flight per thread).
// ================================================================================================
constexpr int A1_THREADS = 512;
struct A1Params {
    const float* x; const float* trans; int B, N;
    const float* W1; BnState st; float* A1;
    double* part;                 // [gridDim.y * gridDim.x][64] or null (eval)
    unsigned* counter; const float* W2; double count; float* mean_u2; double* S1a;
    unsigned* bad; float limit;
};

__global__ void __launch_bounds__(A1_THREADS) k_a1(A1Params p) {
    __shared__ float xs[3][A1_CHUNK * A1_CPB];
    __shared__ double sh[A1_THREADS];
    __shared__ double vs[C1];
    const int tid = (int)threadIdx.x, k = tid & 63, q = tid >> 6;      // q = 0..7
    const int b = (int)blockIdx.y;
    const float w0 = p.W1[k * 3 + 0], w1 = p.W1[k * 3 + 1], w2 = p.W1[k * 3 + 2];
    const float sc = p.st.scale[k], sh_ = p.st.shift[k];
    const int n0 = (int)blockIdx.x * (A1_CHUNK * A1_CPB);
    const int nv = (p.N - n0 < A1_CHUNK * A1_CPB) ? p.N - n0 : A1_CHUNK * A1_CPB;
    {
        // the block's (up to) 1024 points, transformed, staged once
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
        if (tid < nv) {
            const float* xb = p.x + (size_t)b * 3 * p.N + n0 + tid;
            const float p0 = xb[0], p1 = xb[p.N], p2 = xb[2 * p.N];
            t0 = p0; t1 = p1; t2 = p2;
            if (p.trans) {
                const float* T = p.trans + (size_t)b * 9;
                t0 = T[0] * p0 + T[3] * p1 + T[6] * p2;
                t1 = T[1] * p0 + T[4] * p1 + T[7] * p2;
                t2 = T[2] * p0 + T[5] * p1 + T[8] * p2;
            }
        }
        xs[0][tid] = t0; xs[1][tid] = t1; xs[2][tid] = t2;
    }
    __syncthreads();
    float acc = 0.f;
    bool flag = false;
    float* out = p.A1 + ((size_t)b * p.N + n0) * C1 + k;
    // thread (k, q): points q, q+8, ... : a warp writes one 128-byte segment of a point's row
#pragma unroll 8
    for (int pp = q; pp < nv; pp += 8) {
        const float u = w0 * xs[0][pp] + w1 * xs[1][pp] + w2 * xs[2][pp];
        const float a = relu_nan(sc * u + sh_);
        flag = flag || !(a <= p.limit);
        out[(size_t)pp * C1] = a;
        acc += a;
    }
    if (flag) p.bad[b] = 1u;
    if (!p.part) return;
    const unsigned nblk = gridDim.x * gridDim.y;
    sh[tid] = (double)acc;
    __syncthreads();
    if (tid < 64) {
        double t = 0.0;
#pragma unroll
        for (int l = 0; l < A1_THREADS / 64; ++l) t += sh[l * 64 + tid];
        p.part[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * C1 + tid] = t;
    }
    if (!last_block_done(p.counter, nblk)) return;
    // ---- last block: S1a = sum of the partial rows (64 columns x 16 lanes), mean(u2) = W2 S1a / count
    {
        double s = 0.0;
#pragma unroll 8
        for (unsigned r = (unsigned)q; r < nblk; r += A1_THREADS / 64) s += p.part[(size_t)r * C1 + k];
        sh[tid] = s;
        __syncthreads();
        if (tid < 64) {
            double t = 0.0;
#pragma unroll
            for (int l = 0; l < A1_THREADS / 64; ++l) t += sh[l * 64 + tid];
            vs[tid] = t;
            if (p.S1a) p.S1a[tid] = t;
        }
        __syncthreads();
        if (tid < C2) {
            double s2 = 0.0;
#pragma unroll 8
            for (int kk = 0; kk < C1; ++kk) s2 += (double)p.W2[(size_t)tid * C1 + kk] * vs[kk];
            p.mean_u2[tid] = (float)(s2 / p.count);
        }
    }
}

// ================================================================================================
// T3: k_tail_l2 (train) -- between the layer-2 GEMM and the layer-3 GEMM
//   every block: BatchNorm2 batch statistics from the partial rows of centred squares -> scale/shift (block 0 also
//                writes the state and updates the running statistics);
//   then the sum of a2 = relu(bn2(u2)): either from the caller's partial rows (CUDA-core path: k_a2_sum over all points,
//                the exact mean) or, tensor-core path, a PILOT estimate from `nsample` points spread over the batch
//                (each block sums its share; the layer-3 kernel accumulates the exact sum while it stages the tiles);
//   tensor-core path (mu_s != null): every block forms the centres of 1024 / TL2_BLOCKS channels from ITS OWN share of the
//                sample, centre[c] = W3[c] . (block sum / block samples): any value near the mean will do (the statistics are
//                corrected exactly afterwards), so no cross-block step is needed; st3.mean = centre, mu_s = centre / inv;
//   CUDA-core path: the last block adds the partial rows (exact sum of a2 -> S1) and forms mean(u3) = W3 S1 / nsample.
// grid = TL2_BLOCKS, block = 1024 = 128 channels x 8 lanes.
// ================================================================================================
constexpr int TL2_BLOCKS = 32;
constexpr int TL2_SPB = 256;                  // sample points per block of the pilot estimate of mean(a2)
struct TailL2Params {
    const float* css; int n_css;              // partial rows [n_css][128]: sum (u2 - c)^2, c = mean_u2
    const float* s1a_part; int n_s1a; double* S1a;     // tensor-core path (c is a pilot): partial rows [n_s1a][64] of sum a1 * 2^4 -> S1a; else null
    double s1a_scale; const float* W2;        //   ... their scale (2^-4), conv2.weight [128][64]
    int bn_done;                              // 1: BatchNorm2 is already finalised (second call of the CUDA-core path): only read st2
    const float* mean_u2; double count; const float* bias2; pgpd_bn bn2; BnState st2;
    const float* Y2; size_t npoints, pstride;  // pilot (a2part == null): block b samples points ((b * TL2_SPB + j) * pstride) mod npoints
    size_t nsample;                           // CUDA-core path: number of points the partial rows cover
    const double* a2part; int n_a2part;       // or: exact partial rows [n][128] (nsample = number of points they cover)
    double* part;                             // [TL2_BLOCKS][128] scratch
    unsigned* counter;
    double* S1;                               // [128] sum of a2 over the nsample points (CUDA-core path)
    const float* W3; float* mean_u3;          // [1024] (pilot) mean of u3
    const float* inv3; float* mu_s;           // tensor-core path: accumulator scale / the centres in accumulator units
};

__global__ void __launch_bounds__(1024) k_tail_l2(TailL2Params p) {
    __shared__ double sh[1024];
    __shared__ float s_sc[C2], s_sf[C2];
    __shared__ double vs[C2];
    const int tid = (int)threadIdx.x, c = tid & 127, q = tid >> 7;      // q = 0..7
    // ---- BatchNorm2 statistics (every block, identically)
    if (p.bn_done) {
        if (tid < C2) { s_sc[tid] = p.st2.scale[tid]; s_sf[tid] = p.st2.shift[tid]; }
        __syncthreads();
    } else {
        // Tensor-core path: the layer-2 kernel centred its squares on a pilot c and summed a1 exactly; the exact mean of u2 is
        // W2 (sum a1) / M and var = sum (u-c)^2 / M - (mean - c)^2 (an identity).  CUDA-core path: c is the exact mean already.
        const bool pilot = p.s1a_part != nullptr;
        __shared__ double s_a[C1];
        __shared__ double s_dm[C2];
        // every global load of this section is issued before the first barrier (one latency epoch): the css rows, and on the
        // tensor-core path the a1-sum rows and this thread's eight weights of W2
        double s = 0.0;
#pragma unroll 8
        for (int r = q; r < p.n_css; r += 8) s += (double)p.css[(size_t)r * C2 + c];
        if (pilot) {
            const int k = tid & 63, ln = tid >> 6;          // 64 columns x 16 lanes, lanes added in order
            float4 wa = *reinterpret_cast<const float4*>(p.W2 + (size_t)c * C1 + q * 8);
            float4 wb = *reinterpret_cast<const float4*>(p.W2 + (size_t)c * C1 + q * 8 + 4);
            double t = 0.0;
#pragma unroll 8
            for (int r = ln; r < p.n_s1a; r += 16) t += (double)p.s1a_part[(size_t)r * C1 + k];
            sh[tid] = t;
            __syncthreads();
            if (tid < C1) {
                double tt = 0.0;
#pragma unroll
                for (int l = 0; l < 16; ++l) tt += sh[l * C1 + tid];
                tt *= p.s1a_scale;
                s_a[tid] = tt;
                if (blockIdx.x == 0) p.S1a[tid] = tt;       // the backward's dW2 needs it
            }
            __syncthreads();
            const double* sa = s_a + q * 8;
            const double m = (double)wa.x * sa[0] + (double)wa.y * sa[1] + (double)wa.z * sa[2] + (double)wa.w * sa[3]
                           + (double)wb.x * sa[4] + (double)wb.y * sa[5] + (double)wb.z * sa[6] + (double)wb.w * sa[7];
            sh[tid] = m;
            __syncthreads();
            if (tid < C2) {
                double mm = 0.0;
#pragma unroll
                for (int l = 0; l < 8; ++l) mm += sh[l * C2 + tid];
                s_dm[tid] = mm / p.count - (double)p.mean_u2[tid];
            }
            __syncthreads();
        }
        sh[tid] = s;
        __syncthreads();
        if (tid < C2) {
            double tsq = 0.0;
#pragma unroll
            for (int l = 0; l < 8; ++l) tsq += sh[l * C2 + tid];
            const double dm = pilot ? s_dm[tid] : 0.0;
            const double mean = (double)p.mean_u2[tid] + dm;
            const double var0 = tsq / p.count - dm * dm;
            const double var = var0 < 0.0 ? 0.0 : var0;
            const float rstd = (float)(1.0 / sqrt(var + (double)BN_EPS));
            const float sc = p.bn2.gamma[tid] * rstd;
            s_sc[tid] = sc;
            s_sf[tid] = p.bn2.beta[tid] - sc * (float)mean;
            if (blockIdx.x == 0) bn_finalize_train(tid, mean, var0, p.count, p.bias2, p.bn2, p.st2);
        }
        __syncthreads();
    }
    // ---- this block's share of the sum of a2
    double acc = 0.0;
    if (p.a2part) {
#pragma unroll 4
        for (int r = (int)blockIdx.x * 8 + q; r < p.n_a2part; r += 8 * (int)gridDim.x) acc += p.a2part[(size_t)r * C2 + c];
    } else {
        // pilot: this block's own TL2_SPB sample points, spread over the batch (indices wrap around when the batch is small)
        const float sc = s_sc[c], sf = s_sf[c];
        float f = 0.f;
#pragma unroll 8
        for (int j = q; j < TL2_SPB; j += 8) {
            const size_t P = (((size_t)blockIdx.x * TL2_SPB + j) * p.pstride) % p.npoints;
            f += relu_nan(sc * p.Y2[P * C2 + c] + sf);
        }
        acc = (double)f;
    }
    sh[tid] = acc;
    __syncthreads();
    if (tid < C2) {
        double t = 0.0;
#pragma unroll
        for (int l = 0; l < 8; ++l) t += sh[l * C2 + tid];
        p.part[(size_t)blockIdx.x * C2 + tid] = t;
        vs[tid] = t;
    }
    const int warp = tid >> 5, lane = tid & 31;
    if (p.mu_s) {
        // tensor-core path: centres of my 32 channels from my own samples (one warp per row of W3: coalesced, fixed shuffle tree)
        __syncthreads();
        const double inv_n = 1.0 / (double)TL2_SPB;
        const int r = (int)blockIdx.x * (C3 / TL2_BLOCKS) + warp;
        if (warp < C3 / TL2_BLOCKS) {
            double s = 0.0;
#pragma unroll
            for (int kk = lane; kk < C2; kk += 32) s += (double)p.W3[(size_t)r * C2 + kk] * vs[kk];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) {
                const float m = (float)(s * inv_n);
                p.mean_u3[r] = m;
                p.mu_s[r] = m / p.inv3[r];
            }
        }
        return;
    }
    if (!last_block_done(p.counter, gridDim.x)) return;
    // ---- CUDA-core path, last block: total of the partial rows (exact sum of a2)
    {
        double s = 0.0;
        for (unsigned r = (unsigned)q; r < gridDim.x; r += 8) s += p.part[(size_t)r * C2 + c];
        __syncthreads();
        sh[tid] = s;
        __syncthreads();
        if (tid < C2) {
            double t = 0.0;
#pragma unroll
            for (int l = 0; l < 8; ++l) t += sh[l * C2 + tid];
            vs[tid] = t;
            if (p.S1) p.S1[tid] = t;
        }
        __syncthreads();
    }
    // mean(u3) = W3 S / nsample  (one warp per row of W3: coalesced, fixed shuffle tree)
    const double inv_n = 1.0 / (double)p.nsample;
    for (int r = warp; r < C3; r += 32) {
        double s = 0.0;
#pragma unroll
        for (int kk = lane; kk < C2; kk += 32) s += (double)p.W3[(size_t)r * C2 + kk] * vs[kk];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) p.mean_u3[r] = (float)(s * inv_n);
    }
}

// ================================================================================================
// T4: k_tail_l3 -- after the layer-3 GEMM + max-pool kernel; split by output channel (TL3_CH channels per block), so
// there is no cross-block step:
//   train: exact sum of a2 from the GEMM kernel's partial rows (tensor-core path with a pilot centre) -> exact mean(u3);
//          BatchNorm3 batch statistics from the partial rows of squares (centred on `centre`):
//          var = sum (u-c)^2 / count - (mean - c)^2;  finalisation + running statistics;
//   both:  decode the (max, arg-max) keys and apply BatchNorm3 (+ReLU) to the pooled [B][1024] values;
//          a cloud flagged in bad[] (NaN / out-of-range activation) or a pooled value outside the fp16 operand range of
//          the tensor-core heads gets NaN instead of a silently clamped number.
// grid = 1024 / TL3_CH, block = 1024 = TL3_CH channels x 64 lanes.
// ================================================================================================
constexpr int TL3_CH = 16;
constexpr int TL3_LANES = 64;
struct TailL3Params {
    int B; int relu_last; int train;
    const unsigned long long* keys; const float* sgn; BnState st3;
    float* pooled; float* uext; int* idx;     // uext / idx may be null (no backward)
    const unsigned* bad; float limit;
    // train
    const float* css; int n_css;              // [n_css][1024]
    const float* s1part; int n_s1; double s1scale;   // [n_s1][128] partial sums of a2 * 2^4, or null
    const float* W3; const float* bias3; pgpd_bn bn3; double count;
    double* S1;                               // [128] exact sum of a2 (written by block 0 when s1part != null)
};

__global__ void __launch_bounds__(1024) k_tail_l3(TailL3Params p) {
    __shared__ double sh[TL3_LANES][TL3_CH + 1];
    __shared__ double vs[C2];
    __shared__ float s_sc[TL3_CH], s_sf[TL3_CH];
    const int tid = (int)threadIdx.x, cx = tid & (TL3_CH - 1), ln = tid >> 4;      // ln = 0..63
    const int c0 = (int)blockIdx.x * TL3_CH, c = c0 + cx;
    if (p.train) {
        double mean_exact = 0.0;
        const bool pilot = p.s1part != nullptr;
        if (pilot) {
            // exact sum of a2: one partial row per CTA of the GEMM kernel; every block repeats the same sums
            // (128 columns x 8 lanes, lanes added in order)
            {
                const int k = tid & 127, l8 = tid >> 7;
                double t = 0.0;
#pragma unroll 8
                for (int r = l8; r < p.n_s1; r += 8) t += (double)p.s1part[(size_t)r * C2 + k];
                double* shf = &sh[0][0];
                shf[l8 * C2 + k] = t;              // 1024 doubles fit: sh is 64 x 17
            }
            __syncthreads();
            if (tid < C2) {
                const double* shf = &sh[0][0];
                double t = 0.0;
#pragma unroll
                for (int l = 0; l < 8; ++l) t += shf[l * C2 + tid];
                t *= p.s1scale;
                vs[tid] = t;
                if (blockIdx.x == 0 && p.S1) p.S1[tid] = t;
            }
            __syncthreads();
            // mean(u3)[c] = W3[c] . S / count : 64 lanes x 2 k each, lanes added in order
            double s = 0.0;
#pragma unroll
            for (int kk = ln * 2; kk < ln * 2 + 2; ++kk) s += (double)p.W3[(size_t)c * C2 + kk] * vs[kk];
            sh[ln][cx] = s;
            __syncthreads();
            if (ln == 0) {
                double t = 0.0;
#pragma unroll 8
                for (int l2 = 0; l2 < TL3_LANES; ++l2) t += sh[l2][cx];
                mean_exact = t / p.count;
            }
            __syncthreads();
        }
        double s = 0.0;
#pragma unroll 8
        for (int r = ln; r < p.n_css; r += TL3_LANES) s += (double)p.css[(size_t)r * C3 + c];
        sh[ln][cx] = s;
        __syncthreads();
        if (ln == 0) {
            double t = 0.0;
#pragma unroll 8
            for (int l2 = 0; l2 < TL3_LANES; ++l2) t += sh[l2][cx];
            const double centre = (double)p.st3.mean[c];         // what the GEMM kernel centred its squares on
            double var = t / p.count, mu = centre;
            if (pilot) { mu = mean_exact; const double d = mu - centre; var -= d * d; }
            bn_finalize_train(c, mu, var, p.count, p.bias3, p.bn3, p.st3);
            s_sc[cx] = p.st3.scale[c]; s_sf[cx] = p.st3.shift[c];
        }
        __syncthreads();
    } else {
        if (tid < TL3_CH) { s_sc[tid] = p.st3.scale[c0 + tid]; s_sf[tid] = p.st3.shift[c0 + tid]; }
        __syncthreads();
    }
    // ---- pooled values of my TL3_CH channels, all clouds
    const float sc = s_sc[cx], sf = s_sf[cx], sg = p.sgn[c];
    const bool bad_all = p.bad[p.B] != 0u;
    const float qnan = __uint_as_float(0x7FC00000u);
#pragma unroll 4
    for (int b = ln; b < p.B; b += TL3_LANES) {
        const size_t i = (size_t)b * C3 + c;
        const unsigned long long key = p.keys[i];
        const float u = sg * ord_decode((unsigned)(key >> 32));
        const int n = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
        float g = sc * u + sf;
        if (p.relu_last) g = relu_nan(g);
        if (bad_all || p.bad[b] != 0u || !(fabsf(g) <= p.limit)) g = qnan;
        p.pooled[i] = g;
        if (p.uext) { p.uext[i] = u; p.idx[i] = n; }
    }
}

// ================================================================================================
// backward tails
// ================================================================================================

// B2: k_q_uvec -- after k_pool_bwd:  Q = W3^T diag(d) W3 (128 x 128),  uvec = W3^T e,  and (tensor-core path) the hi/lo
// operand image of Q for the pass-A kernel.  grid = 32 blocks x 4 rows of Q, block = 1024 = 128 columns j x 8 lanes over the
// 1024 channels (128 channels per lane, summed in order; the 8 lanes are added in lane order).  The __global__ wrapper
// k_q_uvec (tower.cuh) runs this in blocks [0, 32) and the per-cloud sort of the arg-max pairs in the remaining blocks.
struct QuParams {
    const float* W3; const float* dvec; const float* evec;
    float* Q; float* uvec;
    void* qimg; float* inv_s; int act_shift;      // qimg == null: no image (CUDA-core path)
};

__device__ __forceinline__ void q_uvec_block(const QuParams& p) {
    __shared__ float s_q[8][4][128];
    __shared__ double s_u[4][32];              // [row][warp]: per-warp sums of the uvec products (fixed shuffle tree)
    __shared__ float s_mx[4][128];
    const int tid = (int)threadIdx.x, j = tid & 127, ln = tid >> 7, i0 = (int)blockIdx.x * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    double ue[4] = {0.0, 0.0, 0.0, 0.0};
    const int cbeg = ln * 128;
#pragma unroll 8
    for (int c = cbeg; c < cbeg + 128; ++c) {
        const float w = p.W3[(size_t)c * C2 + j];
        const float4 wi = *reinterpret_cast<const float4*>(p.W3 + (size_t)c * C2 + i0);    // same address for the whole block: broadcast
        const float d = p.dvec[c];
        acc[0] = fmaf(wi.x * d, w, acc[0]); acc[1] = fmaf(wi.y * d, w, acc[1]);
        acc[2] = fmaf(wi.z * d, w, acc[2]); acc[3] = fmaf(wi.w * d, w, acc[3]);
    }
    // uvec rows i0..i0+3: the same 128 channels, split over the 128 "columns": thread (ln, j) covers channel c = ln*128 + j
    {
        const int c = cbeg + j;
        const double e = (double)p.evec[c];
        const float4 wi = *reinterpret_cast<const float4*>(p.W3 + (size_t)c * C2 + i0);
        ue[0] = (double)wi.x * e; ue[1] = (double)wi.y * e; ue[2] = (double)wi.z * e; ue[3] = (double)wi.w * e;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        s_q[ln][r][j] = acc[r];
        double u = ue[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) u += __shfl_xor_sync(0xffffffffu, u, o);
        if ((tid & 31) == 0) s_u[r][tid >> 5] = u;
    }
    __syncthreads();
    float q[4] = {0.f, 0.f, 0.f, 0.f};
    if (ln == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float t = s_q[0][r][j];
#pragma unroll
            for (int l = 1; l < 8; ++l) t += s_q[l][r][j];
            q[r] = t;
            p.Q[(size_t)(i0 + r) * C2 + j] = t;
            s_mx[r][j] = fabsf(t);
        }
    } else if (ln == 1 && j < 4) {
        // uvec[i0 + j] = sum of the 32 warp sums, fixed order
        double t = 0.0;
        for (int wv = 0; wv < 32; ++wv) t += s_u[j][wv];
        p.uvec[i0 + j] = (float)t;
    }
#ifndef PGPD_EMU
    if (p.qimg) {
        __syncthreads();
        for (int st = 64; st > 0; st >>= 1) {
            if (ln == 0 && j < st) {
#pragma unroll
                for (int r = 0; r < 4; ++r) s_mx[r][j] = fmaxf(s_mx[r][j], s_mx[r][j + st]);
            }
            __syncthreads();
        }
        if (ln == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int e = tc::prepack_elem(q[r], s_mx[r][0], i0 + r, j, (__half*)p.qimg, 0);
                if (j == 0) p.inv_s[i0 + r] = ldexpf(1.f, -(e + p.act_shift));
            }
        }
    }
#endif
}

// T5: k_tail_ka -- after pass A of the backward (d a2 -> dz2).  Roles by block (block = 1024 threads):
//   [0, n_gram): column slices (256 columns x 4 lanes) of the Gram partial rows -> g2 (tensor-core path: the hi.hi and hi.lo
//               accumulators, 2 x 128 x 128; CUDA-core path: the plain Gram matrix, directly into `gram`);
//   [n_gram, +8): BatchNorm2-backward partial rows (sum dz, sum dz*yhat): 32 columns x 32 lanes each -> bnsum[256];
//   [.., +4) (tensor-core path): per-channel max |dz2| from the kernel's partial maxima: 32 columns x 32 lanes -> pmx[128];
//   last block: Gram = (hh + hl + hl^T) / 256 (tensor-core path); dgamma2, dbeta2, m1, m2.
// (The 64 x 64 precompute of the fused layer-2/1 pass, which needs m1 / m2, runs as extra blocks of k_dw3: kb_prep_row.)
struct TailKaParams {
    const float* gpart; int n_g; int gcols;       // [n_g][gcols] partial rows; gcols = 16384 (plain) or 32768 (hh, hl)
    float* g2;                                    // [gcols] reduced (scratch when sym)
    float* gram; int sym;                         // [128*128]
    const float* bnpart; int n_bn; double count;  // [n_bn][2][128] partial (sum dz, sum dz*yhat)
    double* bnsum;                                // [256] scratch
    float* dgamma; float* dbeta; float* m1; float* m2;
    unsigned* counter;
    float act_scale;
    const float* pmax; int n_pm; float* pmx;      // tensor-core path: [n_pm][2][128] partial maxima of |dz2| -> pmx [128]
};

__global__ void __launch_bounds__(1024) k_tail_ka(TailKaParams p) {
    __shared__ double sh[32][33];
    const int tid = (int)threadIdx.x;
    const int n_gram = p.gcols / 256;
    int blk = (int)blockIdx.x;
    if (blk < n_gram) {
        const int cl = tid & 255, ln = tid >> 8;          // 4 lanes
        const int col = blk * 256 + cl;
        double s = 0.0;
#pragma unroll 8
        for (int r = ln; r < p.n_g; r += 4) s += (double)p.gpart[(size_t)r * p.gcols + col];
        double* shf = &sh[0][0];                          // 1024 doubles
        shf[ln * 256 + cl] = s;
        __syncthreads();
        if (ln == 0) {
            const double t = ((shf[cl] + shf[256 + cl]) + shf[512 + cl]) + shf[768 + cl];
            (p.sym ? p.g2 : p.gram)[col] = (float)t;
        }
    } else if (blk < n_gram + 8) {
        const int cx = tid & 31, ry = tid >> 5;
        const int col = (blk - n_gram) * 32 + cx;         // 0..255 of [2][128]
        double s = 0.0;
#pragma unroll 8
        for (int r = ry; r < p.n_bn; r += 32) s += (double)p.bnpart[(size_t)r * 2 * C2 + col];
        sh[ry][cx] = s;
        __syncthreads();
        if (ry == 0) {
            double t = 0.0;
#pragma unroll 8
            for (int q = 0; q < 32; ++q) t += sh[q][cx];
            p.bnsum[col] = t;
        }
    } else {
        const int cx = tid & 31, ry = tid >> 5;
        const int col = (blk - n_gram - 8) * 32 + cx;     // 0..127
        float mx = 0.f;
#pragma unroll 8
        for (int r = ry; r < p.n_pm; r += 32) mx = fmaxf(mx, p.pmax[(size_t)r * 2 * C2 + col]);
        float* shf = reinterpret_cast<float*>(&sh[0][0]);
        shf[ry * 32 + cx] = mx;
        __syncthreads();
        if (ry == 0) {
#pragma unroll 8
            for (int q = 1; q < 32; ++q) mx = fmaxf(mx, shf[q * 32 + cx]);
            p.pmx[col] = mx;
        }
    }
    if (!last_block_done(p.counter, gridDim.x)) return;
    // ================================ last block ================================
    if (p.sym) {
        const float sc = 1.0f / (p.act_scale * p.act_scale);
#pragma unroll 4
        for (int i = tid; i < C2 * C2; i += 1024) {
            const int m = i >> 7, n = i & 127;
            p.gram[i] = (p.g2[i] + p.g2[C2 * C2 + i] + p.g2[C2 * C2 + n * C2 + m]) * sc;
        }
    }
    if (tid < 2 * C2) {
        const double s = p.bnsum[tid];
        if (tid < C2) { p.dbeta[tid] = (float)s; p.m1[tid] = (float)(s / p.count); }
        else { p.dgamma[tid - C2] = (float)s; p.m2[tid - C2] = (float)(s / p.count); }
    }
}

// Row r (0..127) of the precompute of the fused layer-2/1 backward pass, run as extra blocks of k_dw3 (512 threads):
//   r < 64:  K[r][k'] = sum_c W2[c][r] s_c r_c m2_c W2[c][k'],  cvec[r] = sum_c W2[c][r] s_c (r_c m2_c mu_c - m1_c)   (l2bwd.cuh)
//   tensor-core path: esc / einv (per-channel power-of-two scale of dz2, from the maxima pmx; written by row 0) and row r of the
//   two A-operand images of tc_kb.cuh: A1op[r][c] = W2[c][r] s_c einv_c 2^g_r, A2op[r][k'] = -K[r][k'] 2^g_r / 16 (zero rows for
//   r >= 64), ginv[r] = 2^-g_r.
struct KbPrepParams {
    const float* W2; BnState st2; const float* m1; const float* m2; float* Kmat; float* cvec;
    const float* pmx; float* esc; float* einv; void* img1; void* img2; float* ginv; float act_scale;   // img1 == null: CUDA-core path
};

__device__ __forceinline__ void kb_prep_row(const KbPrepParams& p, int r) {
    __shared__ double sd[C2], se[C2];
    __shared__ float s_einv[C2];
    __shared__ double s_part[8][C1];
    __shared__ float s_k[C1];
    __shared__ float s_red[C2 + C1];
    const int tid = (int)threadIdx.x;
    if (tid < C2) {
        const double sc = (double)p.st2.scale[tid], rm2 = (double)p.st2.rstd[tid] * (double)p.m2[tid];
        sd[tid] = sc * rm2;
        se[tid] = sc * (rm2 * (double)p.st2.mean[tid] - (double)p.m1[tid]);
        float ev = 1.f;
#ifndef PGPD_EMU
        if (p.img1) {
            const float mx = p.pmx[tid];
            int e = 139 - (int)((__float_as_uint(mx) >> 23) & 0xFFu);          // max|dz2[.,c]| 2^e in [2^12, 2^13)
            e = (mx > 0.f) ? (e > 100 ? 100 : (e < -100 ? -100 : e)) : 0;
            ev = __uint_as_float((uint32_t)(127 - e) << 23);
            if (r == 0) { p.esc[tid] = __uint_as_float((uint32_t)(127 + e) << 23); p.einv[tid] = ev; }
        }
#endif
        s_einv[tid] = ev;
    }
    __syncthreads();
    if (r < C1) {
        const int kp = tid & 63, ln = tid >> 6;          // 8 lanes x 16 channels
        double a = 0.0;
#pragma unroll 8
        for (int c = ln * 16; c < ln * 16 + 16; ++c) a += (double)p.W2[c * C1 + r] * sd[c] * (double)p.W2[c * C1 + kp];
        s_part[ln][kp] = a;
        __syncthreads();
        if (tid < C1) {
            double t = 0.0;
#pragma unroll
            for (int l = 0; l < 8; ++l) t += s_part[l][tid];
            const float kv = (float)t;
            p.Kmat[r * C1 + tid] = kv;
            s_k[tid] = kv;
        } else if (tid < C1 + 32) {
            const int lane = tid - C1;
            double cv = 0.0;
#pragma unroll
            for (int c = lane; c < C2; c += 32) cv += (double)p.W2[c * C1 + r] * se[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cv += __shfl_xor_sync(0xffffffffu, cv, o);
            if (lane == 0) p.cvec[r] = (float)cv;
        }
        __syncthreads();
    }
#ifndef PGPD_EMU
    if (!p.img1) return;
    // row r (< 64: the images hold only their real rows, tc_kb.cuh) of the two images: 128 + 64 values, common power-of-two scale
    if (r >= C1) return;
    float v = 0.f;
    if (tid < C2) v = p.W2[tid * C1 + r] * p.st2.scale[tid] * s_einv[tid];
    else if (tid < C2 + C1) v = -s_k[tid - C2] * (1.0f / p.act_scale);
    if (tid < C2 + C1) s_red[tid] = fabsf(v);
    __syncthreads();
    if (tid < C1) s_red[tid] = fmaxf(s_red[tid], s_red[tid + C2]);
    __syncthreads();
    for (int st = 64; st > 0; st >>= 1) {
        if (tid < st) s_red[tid] = fmaxf(s_red[tid], s_red[tid + st]);
        __syncthreads();
    }
    const float mx = s_red[0];
    if (tid < C2) {
        const int e = tc::prepack_elem(v, mx, r, tid, (__half*)p.img1, 0, 4096);
        if (tid == 0) p.ginv[r] = ldexpf(1.f, -e);
    } else if (tid < C2 + C1) {
        int ex = 0;
        if (mx > 0.f) frexpf(mx, &ex);
        const int e = (mx > 0.f && mx < INFINITY) ? 14 - ex : 0;
        const int c = tid - C2;
        const float ks = ldexpf(v, e);
        const __half hi = __float2half_rn(ks);
        const __half lo = __float2half_rn(ks - __half2float(hi));
        const int chunk = c >> 3, within = c & 7;
        const size_t off = (size_t)r * 64 + (size_t)((chunk ^ (r & 7)) << 3) + within;
        __half* img2 = (__half*)p.img2;
        img2[off] = hi;
        img2[4096 + off] = lo;
    }
#endif
}

// T7a: k_tail_kb -- after the fused layer-2/1 backward pass (block = 1024 threads):
//   blocks [0, 48): column slices (256 columns x 4 lanes) of the partial rows of C = sum dz2 a1^T (128 x 64) and
//                   Gram1 = sum a1 a1^T (64 x 64);
//   blocks [48, 52): BatchNorm1-backward partial rows (sum dz1, sum dz1 yhat1), 32 columns x 32 lanes each;
//   last block: dgamma1, dbeta1, m1, m2.   (dW2, which needs C and Gram1 whole, runs as extra blocks of k_kb_l1: dw2_row.)
struct TailKbParams {
    const float* Cpart; const float* G1part; int n_parts;     // [n_parts][8192], [n_parts][4096]
    float* Cm; float* G1;
    const float* bnpart; int n_bn; double count;              // [n_bn][2][64]
    double* bnsum;                                            // [128] scratch
    float* dgamma1; float* dbeta1; float* m1_1; float* m2_1;
    unsigned* counter;
};
constexpr int TKB_BLOCKS = 52;

__global__ void __launch_bounds__(1024) k_tail_kb(TailKbParams p) {
    __shared__ double sh[32][33];
    const int tid = (int)threadIdx.x, blk = (int)blockIdx.x;
    if (blk < 48) {
        const int cl = tid & 255, ln = tid >> 8;
        const int col = blk * 256 + cl;                       // 0 .. 12287
        const bool isC = col < C2 * C1;
        const float* src = isC ? p.Cpart + col : p.G1part + (col - C2 * C1);
        const size_t ld = isC ? (size_t)C2 * C1 : (size_t)C1 * C1;
        double s = 0.0;
#pragma unroll 8
        for (int r = ln; r < p.n_parts; r += 4) s += (double)src[(size_t)r * ld];
        double* shf = &sh[0][0];
        shf[ln * 256 + cl] = s;
        __syncthreads();
        if (ln == 0) {
            const double t = ((shf[cl] + shf[256 + cl]) + shf[512 + cl]) + shf[768 + cl];
            if (isC) p.Cm[col] = (float)t; else p.G1[col - C2 * C1] = (float)t;
        }
    } else {
        const int cx = tid & 31, ry = tid >> 5;
        const int col = (blk - 48) * 32 + cx;                 // 0 .. 127 of [2][64]
        double s = 0.0;
#pragma unroll 8
        for (int r = ry; r < p.n_bn; r += 32) s += (double)p.bnpart[(size_t)r * 2 * C1 + col];
        sh[ry][cx] = s;
        __syncthreads();
        if (ry == 0) {
            double t = 0.0;
#pragma unroll 8
            for (int l2 = 0; l2 < 32; ++l2) t += sh[l2][cx];
            p.bnsum[col] = t;
        }
    }
    if (!last_block_done(p.counter, gridDim.x)) return;
    if (tid < 2 * C1) {
        const double s = p.bnsum[tid];
        if (tid < C1) { p.dbeta1[tid] = (float)s; p.m1_1[tid] = (float)(s / p.count); }
        else { p.dgamma1[tid - C1] = (float)s; p.m2_1[tid - C1] = (float)(s / p.count); }
    }
}

// Row c (0..127) of dW2 = diag(s)[C - m1 S1a^T - diag(r m2)(W2 Gram1 - mu2 S1a^T)], run as extra blocks of k_kb_l1
// (first 256 threads: 64 columns x 4 lanes over the 64 terms of W2 Gram1); db2 = 0.
struct Dw2Params {
    const float* Cm; const float* G1; const double* S1a; const float* W2; BnState st2; const float* m1_2; const float* m2_2;
    float* dW2; float* db2;
};

__device__ __forceinline__ void dw2_row(const Dw2Params& p, int c) {
    __shared__ double s_p[4][C1];
    const int tid = (int)threadIdx.x;
    if (tid < 256) {
        const int k = tid & 63, ln = tid >> 6;
        double wg = 0.0;
#pragma unroll 8
        for (int kk = ln * 16; kk < ln * 16 + 16; ++kk) wg += (double)p.W2[c * C1 + kk] * (double)p.G1[kk * C1 + k];
        s_p[ln][k] = wg;
    }
    __syncthreads();
    if (tid < C1) {
        const int k = tid;
        const double wg = ((s_p[0][k] + s_p[1][k]) + s_p[2][k]) + s_p[3][k];
        const double sc = (double)p.st2.scale[c], rm2 = (double)p.st2.rstd[c] * (double)p.m2_2[c];
        const double v = (double)p.Cm[c * C1 + k] - (double)p.m1_2[c] * p.S1a[k] - rm2 * (wg - (double)p.st2.mean[c] * p.S1a[k]);
        p.dW2[c * C1 + k] = (float)(sc * v);
        if (k == 0 && p.db2) p.db2[c] = 0.f;     // bias feeding a train-mode BatchNorm: gradient is identically zero
    }
}

}  // namespace pgpd
