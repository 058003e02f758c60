// mma_time.cu -- tcgen05.mma micro-benchmark (dev tool): cycles per MMA for the tile shapes / operand majorness the
// kernels use, and the TMEM lane mapping of an M = 64 accumulator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I pointnetgpd_b200/csrc -I include -o build/mma_time.bin tests/tc_probe/mma_time.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_accum.cuh"
using namespace pgpd::tc;

struct Cfg { int M, N, mn, reps; };

__global__ void __launch_bounds__(192, 1) k_time(Cfg c, long long* out, float* lanes) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // A at sbase (32 KB), B at sbase + 64 KB (64 KB)
    for (int i = tid; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    __syncthreads();
    if (tid < 128) {   // A[m][0] = m+1 (K-major SW128: k=0 is chunk 0 ^ (m&7))
        const int m = tid;
        *reinterpret_cast<__half*>(smem + m * 128 + (((0) ^ (m & 7)) << 4)) = __float2half((float)(m + 1));
    }
    for (int n = tid; n < 256; n += blockDim.x)
        *reinterpret_cast<__half*>(smem + 65536 + n * 128 + (((0) ^ (n & 7)) << 4)) = __float2half(1.0f);
    const uint32_t b32 = smem_u32(&bar);
    if (tid == 0) { mbar_init(b32, 1); mbar_fence_init(); }
    if (warp == 4) tmem_alloc<512>(smem_u32(&tslot));
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tslot;
    if (warp == 5 && lane == 0) {
        const uint32_t idesc = c.mn ? idesc_f16_mn(c.M, c.N) : idesc_f16(c.M, c.N);
        const uint64_t da = c.mn ? desc_sw128_mnmajor(sbase, 8192) : desc_sw128_kmajor(sbase);
        const uint64_t db = c.mn ? desc_sw128_mnmajor(sbase + 65536, 8192) : desc_sw128_kmajor(sbase + 65536);
        // warm-up
        mma_f16(tmem, da, db, idesc, 0u);
        mma_commit(b32);
        mbar_wait(b32, 0);
        const long long t0 = clock64();
        for (int r = 0; r < c.reps; ++r) mma_f16(tmem, da, db, idesc, r ? 1u : 0u);
        mma_commit(b32);
        mbar_wait(b32, 1);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    tc_fence_after_sync();
    if (warp < 4 && lanes && blockIdx.x == 0) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
        lanes[warp * 32 + lane] = v[0] / (float)c.reps;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc<512>(tmem);
}

int main() {
    long long* out; float* lanes;
    cudaMalloc(&out, 256 * 8); cudaMalloc(&lanes, 128 * 4);
    cudaFuncSetAttribute(k_time, cudaFuncAttributeMaxDynamicSharedMemorySize, 165 * 1024);
    const Cfg cfgs[] = {{128, 64, 0, 512}, {128, 128, 0, 512}, {128, 256, 0, 512}, {64, 64, 0, 512}, {64, 128, 0, 512}, {64, 256, 0, 512},
                        {128, 64, 1, 512}, {128, 128, 1, 512}, {128, 32, 0, 512}, {128, 16, 0, 512}};
    for (const Cfg& c : cfgs) {
        cudaMemset(lanes, 0, 512);
        k_time<<<148, 192, 165 * 1024>>>(c, out, lanes);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("M=%d N=%d mn=%d: CUDA error %s\n", c.M, c.N, c.mn, cudaGetErrorString(e)); return 1; }
        long long h[148]; float hl[128];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        cudaMemcpy(hl, lanes, sizeof(hl), cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < 148; ++i) s += (double)h[i];
        printf("M=%3d N=%3d %s: %.1f cycles/MMA (K=16)\n", c.M, c.N, c.mn ? "MN-major" : "K-major ", s / 148 / c.reps);
        if (c.M == 64 && c.N == 64) {
            printf("  M=64 accumulator: TMEM lane -> row+1 (col 0):");
            for (int l = 0; l < 128; ++l) { if (l % 16 == 0) printf("\n   lanes %3d..: ", l); printf("%3.0f ", hl[l]); }
            printf("\n");
        }
    }
    return 0;
}
