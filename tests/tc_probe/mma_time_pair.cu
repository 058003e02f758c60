// mma_time_pair.cu -- cycles per tcgen05.mma.cta_group::2 (dev tool; see mma_time.cu)
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_accum.cuh"
using namespace pgpd::tc;

struct Cfg { int M, N, reps; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1) k_time_pair(Cfg c, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    for (int i = tid; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    const uint32_t b32 = smem_u32(&bar);
    if (tid == 0) { mbar_init(b32, 1); mbar_fence_init(); }
    if (warp == 4) tmem_alloc_pair<512>(smem_u32(&tslot));
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem = tslot;
    if (warp == 5 && lane == 0) {
        long long t0 = 0;
        if (rank == 0) {
            const uint32_t idesc = idesc_f16(c.M, c.N);
            const uint64_t da = desc_sw128_kmajor(sbase), db = desc_sw128_kmajor(sbase + 65536);
            mma_f16_pair(tmem, da, db, idesc, 0u);
            mma_commit_pair(b32, (uint16_t)0x3);
        }
        mbar_wait(b32, 0);
        t0 = clock64();
        if (rank == 0) {
            const uint32_t idesc = idesc_f16(c.M, c.N);
            const uint64_t da = desc_sw128_kmajor(sbase), db = desc_sw128_kmajor(sbase + 65536);
            for (int r = 0; r < c.reps; ++r) mma_f16_pair(tmem, da, db, idesc, r ? 1u : 0u);
            mma_commit_pair(b32, (uint16_t)0x3);
        }
        mbar_wait(b32, 1);
        out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 4) tmem_dealloc_pair<512>(tmem);
}

int main() {
    long long* out;
    cudaMalloc(&out, 256 * 8);
    cudaFuncSetAttribute(k_time_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 165 * 1024);
    const Cfg cfgs[] = {{256, 256, 512}, {256, 128, 512}, {256, 64, 512}, {128, 256, 512}};
    for (const Cfg& c : cfgs) {
        k_time_pair<<<148, 192, 165 * 1024>>>(c, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("pair M=%d N=%d: CUDA error %s\n", c.M, c.N, cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < 148; ++i) s += (double)h[i];
        printf("cta_group::2 M=%3d N=%3d K-major: %.1f cycles/MMA (K=16)\n", c.M, c.N, s / 148 / c.reps);
    }
    return 0;
}
