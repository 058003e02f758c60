// prep.cuh -- the two data-preparation steps immediately before the model, on the GPU:
//   * gripper-box crop of a cloud for many grasps   (reference: PointNetGPD/model/dataset.py:51-76, kinect2grasp.py:178-235)
//   * resampling every crop to exactly N points     (dataset.py:438-444, kinect2grasp.py:473-478)
// Both are HBM-bound integer/compare work: coalesced loads, ballot-based stable compaction, no tensor cores.
#pragma once
#include "common.cuh"

namespace pgpd {

// frame of one grasp: center[3], M[9] (rows approach, binormal, minor_normal), limits[3] (x, y, z half extents)
constexpr int FRAME_DOUBLES = 15;

// One block per grasp scans the whole cloud.  Arithmetic in double, like numpy in the reference (float32 points are
// promoted when the float64 centre is subtracted), so that the inside/outside decisions agree bit for bit.
// counts != nullptr : write the number of points inside the box of each grasp.
// out_idx/out_pts   : (optional) write the ascending indices / local coordinates at offsets[g] + running position.
__global__ void k_crop_box(const float* __restrict__ pc, int P, const double* __restrict__ frames,
                           const int* __restrict__ offsets, int* __restrict__ counts,
                           float* __restrict__ out_pts, int* __restrict__ out_idx) {
    __shared__ int warp_cnt[8];
    __shared__ int base_s;
    const int g = (int)blockIdx.x, tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double* f = frames + (size_t)g * FRAME_DOUBLES;
    const double cx = f[0], cy = f[1], cz = f[2];
    const double m00 = f[3], m01 = f[4], m02 = f[5], m10 = f[6], m11 = f[7], m12 = f[8], m20 = f[9], m21 = f[10], m22 = f[11];
    const double lx = f[12], ly = f[13], lz = f[14];
    if (tid == 0) base_s = 0;
    __syncthreads();
    const int out0 = offsets ? offsets[g] : 0;
    for (int i0 = 0; i0 < P; i0 += 256) {
        const int i = i0 + tid;
        bool in = false;
        double tx = 0, ty = 0, tz = 0;
        if (i < P) {
            const double px = (double)pc[(size_t)i * 3 + 0] - cx, py = (double)pc[(size_t)i * 3 + 1] - cy, pz = (double)pc[(size_t)i * 3 + 2] - cz;
            tx = m00 * px + m01 * py + m02 * pz;
            ty = m10 * px + m11 * py + m12 * pz;
            tz = m20 * px + m21 * py + m22 * pz;
            in = tx > -lx && tx < lx && ty > -ly && ty < ly && tz > -lz && tz < lz;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, in ? 1 : 0);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 8; ++w) { const int c = warp_cnt[w]; if (w < warp) before += c; total += c; }
        const int base = base_s;
        if (in && out_idx) {
            const int pos = out0 + base + before + __popc(bal & ((1u << lane) - 1u));
            out_idx[pos] = i;
            out_pts[(size_t)pos * 3 + 0] = (float)tx; out_pts[(size_t)pos * 3 + 1] = (float)ty; out_pts[(size_t)pos * 3 + 2] = (float)tz;
        }
        __syncthreads();
        if (tid == 0) base_s = base + total;
        __syncthreads();
    }
    if (tid == 0 && counts) counts[g] = base_s;
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Resample candidate c (points pts[offsets[c] .. offsets[c+1])) to exactly N points, `repeat` independent draws.
//   n >= N : a uniformly random SUBSET of N distinct points (np.random.choice(n, N, replace=False) up to order; the
//            model is invariant to the order of points), written in ascending index order;
//   n <  N : N independent uniform indices (replace=True).
// Selection without sorting: every point gets a pseudo-random 64-bit key (unique: the index is its low word); a binary
// search finds the threshold below which exactly N keys lie; a ballot compaction writes the selected points.
// grid = (candidates, repeat), block = 256.  Output x is channel-major [C*repeat][3][N] (the model's input layout).
__global__ void k_resample(const float* __restrict__ pts, const int* __restrict__ offsets, int N, unsigned long long seed,
                           float* __restrict__ out_x, int* __restrict__ out_idx) {
    __shared__ int red[256];
    __shared__ int warp_cnt[8];
    __shared__ int base_s;
    const int c = (int)blockIdx.x, rep = (int)blockIdx.y, R = (int)gridDim.y;
    const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int o0 = offsets[c], n = offsets[c + 1] - o0;
    const size_t row = (size_t)c * R + rep;
    float* x = out_x + row * 3 * N;
    int* oi = out_idx ? out_idx + row * N : nullptr;
    const unsigned long long stream = mix64(seed ^ mix64(((unsigned long long)c << 20) ^ (unsigned long long)rep));
    if (n <= 0) {
        for (int r = tid; r < N; r += 256) { x[r] = 0.f; x[N + r] = 0.f; x[2 * N + r] = 0.f; if (oi) oi[r] = -1; }
        return;
    }
    if (n < N) {
        for (int r = tid; r < N; r += 256) {
            const int j = (int)(mix64(stream + (unsigned long long)r) % (unsigned long long)n);
            const float* p = pts + (size_t)(o0 + j) * 3;
            x[r] = p[0]; x[N + r] = p[1]; x[2 * N + r] = p[2];
            if (oi) oi[r] = j;
        }
        return;
    }
    // ---- n >= N: threshold selection ----
    auto key = [&](int j) { return (mix64(stream + (unsigned long long)j) & 0xFFFFFFFF00000000ull) | (unsigned long long)(unsigned)j; };
    unsigned long long lo = 0ull, hi = ~0ull;          // invariant: count(key < lo) <= N <= count(key < hi) ... find smallest t with count(key <= t) >= N
    // binary search on t for: count(key <= t) >= N
    for (int it = 0; it < 64; ++it) {
        const unsigned long long mid = lo + ((hi - lo) >> 1);
        int cnt = 0;
        for (int j = tid; j < n; j += 256) cnt += key(j) <= mid ? 1 : 0;
        red[tid] = cnt;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) { if (tid < s) red[tid] += red[tid + s]; __syncthreads(); }
        const int total = red[0];
        __syncthreads();
        if (total >= N) hi = mid; else lo = mid + 1;
        if (lo >= hi) break;
    }
    const unsigned long long thr = hi;                  // exactly N keys are <= thr (keys are distinct)
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int j0 = 0; j0 < n; j0 += 256) {
        const int j = j0 + tid;
        const bool sel = j < n && key(j) <= thr;
        const unsigned bal = __ballot_sync(0xffffffffu, sel ? 1 : 0);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 8; ++w) { const int cc = warp_cnt[w]; if (w < warp) before += cc; total += cc; }
        const int base = base_s;
        if (sel) {
            const int r = base + before + __popc(bal & ((1u << lane) - 1u));
            if (r < N) {
                const float* p = pts + (size_t)(o0 + j) * 3;
                x[r] = p[0]; x[N + r] = p[1]; x[2 * N + r] = p[2];
                if (oi) oi[r] = j;
            }
        }
        __syncthreads();
        if (tid == 0) base_s = base + total;
        __syncthreads();
    }
}

}  // namespace pgpd
