// cuda_emu.h -- a tiny single-threaded CUDA *SIMT* emulator for g++.
//
// TEST INFRASTRUCTURE ONLY.  It exists so that the CUDA-core (SIMT) kernels and the C++ host
// orchestration of libpgpd can be exercised in the GPU-less build container, against the
// oracle, BEFORE spending GPU minutes.  It is compiled into tests/simt_emu/libpgpd_emu.so by
// tests/simt_emu/build.py and loaded only by tests (tests/emu_util.py).  The product library
// (pointnetgpd_b200/csrc -> libpgpd.so, built by nvcc for sm_100a) never includes this file
// and the product package never loads the emulator build.
//
// Model: one OS thread; every CUDA thread of a block is a ucontext fiber; blocks run one
// after another.  Fibers switch only at __syncthreads / warp shuffles, so execution is
// deterministic.  __shared__ becomes `static` (one block alive at a time).  A barrier that
// can never complete (divergent __syncthreads) is detected and aborts with a message.
#pragma once
#include <ucontext.h>
#include <sys/mman.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __restrict__ __restrict

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) int4 { int x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }

namespace emu {

constexpr size_t kStack = 256 * 1024;
constexpr int kMaxThreads = 1024;

struct Warp { int count = 0; int gen = 0; uint64_t slot[32]; };

struct State {
    ucontext_t sched;
    ucontext_t ctx[kMaxThreads];
    char* stacks = nullptr;
    bool done[kMaxThreads];
    uint3 tids[kMaxThreads];
    int nthreads = 0, live = 0, cur = 0;
    int bar_count = 0, bar_gen = 0;
    bool progress = false;
    Warp warps[kMaxThreads / 32];
    std::function<void()>* body = nullptr;
    uint3 blockIdx_{0, 0, 0};
    dim3 blockDim_, gridDim_;
    std::vector<unsigned char> dyn;
};

inline State& S() { static State s; return s; }

inline void yield() { State& s = S(); swapcontext(&s.ctx[s.cur], &s.sched); }

inline void fiber_entry() {
    State& s = S();
    (*s.body)();
    s.done[s.cur] = true;
    s.live--;
    s.progress = true;
    if (s.bar_count > 0 && s.bar_count == s.live) { s.bar_count = 0; s.bar_gen++; }
    swapcontext(&s.ctx[s.cur], &s.sched);
}

inline void syncthreads() {
    State& s = S();
    int gen = s.bar_gen;
    if (++s.bar_count == s.live) { s.bar_count = 0; s.bar_gen++; s.progress = true; }
    else while (s.bar_gen == gen) yield();
}

inline void warp_barrier(Warp& w, int members) {
    State& s = S();
    int gen = w.gen;
    if (++w.count == members) { w.count = 0; w.gen++; s.progress = true; }
    else while (w.gen == gen) yield();
}

inline int warp_members(int warp) {
    State& s = S();
    int rem = s.nthreads - warp * 32;
    return rem >= 32 ? 32 : rem;
}

template <class T>
inline T shfl_from(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload too large");
    State& s = S();
    int warp = s.cur / 32, lane = s.cur % 32;
    Warp& w = s.warps[warp];
    int members = warp_members(warp);
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    w.slot[lane] = bits;
    warp_barrier(w, members);
    T r = v;
    if (src_lane >= 0 && src_lane < members) memcpy(&r, &w.slot[src_lane], sizeof(T));
    warp_barrier(w, members);
    return r;
}

inline void run_block(dim3 block, const std::function<void()>& body) {
    State& s = S();
    int n = (int)(block.x * block.y * block.z);
    if (n > kMaxThreads || n <= 0) { fprintf(stderr, "emu: bad block size %d\n", n); abort(); }
    if (!s.stacks) {
        s.stacks = (char*)mmap(nullptr, kStack * kMaxThreads, PROT_READ | PROT_WRITE,
                               MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (s.stacks == (char*)MAP_FAILED) { perror("emu mmap"); abort(); }
    }
    s.nthreads = s.live = n;
    s.bar_count = 0;
    s.body = const_cast<std::function<void()>*>(&body);
    for (int w = 0; w < (n + 31) / 32; ++w) { s.warps[w].count = 0; }
    for (int i = 0; i < n; ++i) {
        s.done[i] = false;
        s.tids[i] = uint3{(unsigned)(i % block.x), (unsigned)((i / block.x) % block.y), (unsigned)(i / (block.x * block.y))};
        getcontext(&s.ctx[i]);
        s.ctx[i].uc_stack.ss_sp = s.stacks + kStack * i;
        s.ctx[i].uc_stack.ss_size = kStack;
        s.ctx[i].uc_link = &s.sched;
        makecontext(&s.ctx[i], (void (*)())fiber_entry, 0);
    }
    while (s.live > 0) {
        s.progress = false;
        for (int i = 0; i < n; ++i) {
            if (s.done[i]) continue;
            s.cur = i;
            swapcontext(&s.sched, &s.ctx[i]);
        }
        if (!s.progress && s.live > 0) {
            fprintf(stderr, "emu: deadlock -- %d threads wait at a barrier that cannot complete "
                            "(block %u,%u,%u)\n", s.live, s.blockIdx_.x, s.blockIdx_.y, s.blockIdx_.z);
            abort();
        }
    }
}

inline void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    State& s = S();
    s.gridDim_ = grid;
    s.blockDim_ = block;
    s.dyn.assign(smem + 1024, 0);
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                s.blockIdx_ = uint3{x, y, z};
                run_block(block, body);
            }
}

inline void* dyn_smem() {
    State& s = S();
    uintptr_t p = (uintptr_t)s.dyn.data();
    return (void*)((p + 1023) & ~(uintptr_t)1023);
}

}  // namespace emu

#define threadIdx (emu::S().tids[emu::S().cur])
#define blockIdx (emu::S().blockIdx_)
#define blockDim (emu::S().blockDim_)
#define gridDim (emu::S().gridDim_)

static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {      // a real barrier: fibers of a warp do not run in lockstep
    emu::State& s = emu::S();
    const int warp = s.cur / 32;
    emu::warp_barrier(s.warps[warp], emu::warp_members(warp));
}
static inline void __threadfence() {}

template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    int lane = emu::S().cur % 32;
    return emu::shfl_from(v, (lane / width) * width + (src % width));
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    int lane = emu::S().cur % 32;
    int src = lane ^ m;
    if (src / width != lane / width) src = lane;
    return emu::shfl_from(v, src);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
    int lane = emu::S().cur % 32;
    int src = lane + (int)d;
    if (src / width != lane / width) src = lane;
    return emu::shfl_from(v, src);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
    int lane = emu::S().cur % 32;
    int src = lane - (int)d;
    if (src < 0 || src / width != lane / width) src = lane;
    return emu::shfl_from(v, src);
}

static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned bit = pred ? 1u : 0u;
    unsigned out = 0;
    int members = emu::warp_members(emu::S().cur / 32);
    for (int l = 0; l < members; ++l) {
        unsigned b = emu::shfl_from(bit, l);
        out |= (b & 1u) << l;
    }
    return out;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }

template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
using std::max;
using std::min;
