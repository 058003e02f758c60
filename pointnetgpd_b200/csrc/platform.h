// platform.h -- one place where the CUDA build and the test-only SIMT emulator build differ.
//
// Product build: nvcc, sm_100a.  Test-only build (-DPGPD_EMU, g++): the CUDA-core kernels and the
// host orchestration compiled against tests/simt_emu/cuda_emu.h, so they can be checked against
// the oracle without a GPU.  The tcgen05 kernels exist only in the nvcc build.
#pragma once

#ifdef PGPD_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#endif

#include <cstddef>
#include <atomic>

namespace pgpd {

// number of kernels this library launched, process-wide (bench.py reports it; the backward runs on autograd's
// worker thread, so a per-thread counter would miss it)
inline std::atomic<unsigned long long>& launch_counter() { static std::atomic<unsigned long long> n{0}; return n; }

#ifdef PGPD_EMU
template <class T> __device__ __forceinline__ T* dyn_smem() { return reinterpret_cast<T*>(emu::dyn_smem()); }

template <class... KArgs, class... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t, Args... args) {
    ++launch_counter();
    emu::run_grid(grid, block, smem, [&]() { kernel(args...); });
}
#else
template <class T> __device__ __forceinline__ T* dyn_smem() {
    extern __shared__ __align__(1024) unsigned char pgpd_dyn_smem_[];
    return reinterpret_cast<T*>(pgpd_dyn_smem_);
}

template <class... KArgs, class... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    ++launch_counter();
    kernel<<<grid, block, smem, stream>>>(args...);
}
#endif

__host__ __device__ __forceinline__ int idiv_up(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ long long lldiv_up(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace pgpd
