// tc_ptx.cuh -- thin inline-PTX wrappers for the sm_100a features the tensor-core kernels use:
// mbarrier, bulk async copy (UBLKCP), TMEM allocation, tcgen05.mma / commit / ld, proxy fences.
// Every form here was validated on a B200 by tests/tc_probe/probe.cu (profiles/r1_tcgen05_probe.log).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace pgpd { namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Spins until the phase with the given parity has completed.  A wait that can never complete is a
// protocol bug; trap (error, no hang) instead of spinning forever.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 28); ++it)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}

// ---- fences -------------------------------------------------------------------------------------------
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core / bulk copy engine)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- bulk async copy global -> shared, completion on an mbarrier (SASS: UBLKCP) ----------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- thread-block clusters ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {      // every thread of every CTA in the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// bulk copy global -> the SAME shared-memory offset of every CTA in cta_mask; each destination CTA's mbarrier (same
// offset) receives the complete_tx for the bytes it got
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "h"(cta_mask) : "memory");
}
// tcgen05.commit that arrives on the mbarrier at the same offset in every CTA of cta_mask
__device__ __forceinline__ void mma_commit_multicast(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"(cta_mask) : "memory");
}

// ---- CTA pairs (cta_group::2): one MMA spans the tensor cores / shared memories / TMEMs of the two CTAs of a cluster ------
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem) {   // one whole warp in EACH CTA of the pair, same smem offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_smem), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(COLS) : "memory");
}
// D[256 x N] += A[256 x 16] B[N x 16]^T: issued by ONE thread of the leader CTA; each CTA supplies its 128 rows of A and
// its N/2 rows of B at the same shared-memory offsets, and receives its 128 rows of D in its own TMEM
__device__ __forceinline__ void mma_f16_pair(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (once every MMA issued so far has completed) on the mbarrier at the same offset in every CTA of cta_mask
__device__ __forceinline__ void mma_commit_pair(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"(cta_mask) : "memory");
}
// arrive on the mbarrier at local offset `bar` of CTA `rank` of this cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
                 "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" :: "r"(bar), "r"(rank) : "memory");
}
// wait on a LOCAL mbarrier whose arrivals may come from the peer CTA (acquire at cluster scope)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 28); ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}

// ---- warp-uniform single-thread issue --------------------------------------------------------------------------------------------
// tcgen05.mma / tcgen05.commit / TMA take their operands from UNIFORM registers.  Issued under `if (lane == 0)` the operands live in
// per-thread registers and the compiler wraps EVERY instruction in an elect / R2UR.BROADCAST x5 / branch "waterfall" (~130 cycles per
// MMA, measured: the issuing thread, not the tensor pipe, then paces the kernel).  Issued from warp-uniform code -- the whole warp
// runs the loop, operands are computed by all lanes from provably uniform values, one elected lane executes the instruction -- they
// are formed directly in uniform registers.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// a value every lane holds, made provably uniform for the compiler (warp index, a TMEM base address read from shared memory)
__device__ __forceinline__ uint32_t warp_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// ---- tensor-map (TMA) copy for CTA pairs ------------------------------------------------------------------------------------
// 2-D box of a tensor map -> THIS CTA's shared memory; the complete_tx goes to the mbarrier at local offset `bar` of the pair's
// LEADER CTA (rank 0), whichever CTA issues the copy (.cta_group::2 allows the barrier to live in the peer CTA): the MMA issuer
// then waits on ONE barrier for both halves of a weight stage, with no relay through the peer.
__device__ __forceinline__ void tma2d_g2s_pair_leaderbar(uint32_t dst_smem, const void* tmap, int c0, int c1, uint32_t bar) {
    asm volatile("{\n\t.reg .b32 rb;\n\tmapa.shared::cluster.u32 rb, %4, 0;\n\t"
                 "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [rb];\n\t}"
                 :: "r"(dst_smem), "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// pull [src, src+bytes) into L2 ahead of use (bytes multiple of 16, src 16-byte aligned); no completion tracking
__device__ __forceinline__ void l2_prefetch(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src), "r"(bytes) : "memory");
}

// ---- TMEM ---------------------------------------------------------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_smem), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {   // whole warp, the one that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(COLS) : "memory");
}

// ---- UMMA descriptors -----------------------------------------------------------------------------------
// K-major operand tile stored as rows of 128 bytes (64 fp16) with the 128-byte XOR swizzle:
//   byte(row, chunk16) = row*128 + ((chunk16 ^ (row & 7)) << 4);  8-row groups are 1024 B apart.
__device__ __forceinline__ uint64_t desc_sw128_kmajor(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address
    d |= (uint64_t)0 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: 8-row group pitch
    d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
// MN-major SWIZZLE_128B descriptor: atoms of 64 M/N-elements (128 B) x 8 K-rows; atom pitch `lbo` bytes
__device__ __forceinline__ uint64_t desc_sw128_mnmajor(uint32_t saddr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor for kind::f16, fp16 A/B (K-major), fp32 accumulate, M x N tile
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_f16_mn(int M, int N) { return idesc_f16(M, N) | (1u << 15) | (1u << 16); }   // both operands MN-major
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

// 32 lanes x 32 consecutive columns of fp32: thread `lane` gets row (lane of its warp's TMEM quadrant)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

// 32 lanes x 16 consecutive columns of fp32
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

// named barrier among `count` threads (count a multiple of 32); id 1..15 (0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
}

// split form for software pipelining: issue the load, and later wait for it.  The wait takes the destination
// registers as read-write operands so that the compiler cannot move their uses above it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}

// ---- packed fp32 pairs (sm_100 FADD2 / FFMA2): two lanes of fp32 arithmetic per issued instruction -----------------
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// O(1) activations are multiplied by 2^4 before the fp16 hi/lo split (keeps the lo parts in fp16's normal range)
constexpr float ACT_SCALE = 16.0f;
constexpr int ACT_SHIFT = 4;

// tuning aid (builds with -DPGPD_DEBUG only): when non-null, the streaming kernels add their pipeline cycle counters here
#ifdef PGPD_DEBUG
__device__ long long* g_stream_dbg = nullptr;
#else
constexpr long long* g_stream_dbg = nullptr;
#endif

// fp32 -> (hi, lo) fp16 pair with hi + lo == x to ~22 bits (x must be pre-scaled into fp16's normal range)
__device__ __forceinline__ void split2(float a, float b, __half2& hi, __half2& lo) {
    hi = __floats2half2_rn(a, b);
    float2 back = __half22float2(hi);
    lo = __floats2half2_rn(a - back.x, b - back.y);
}

}}  // namespace pgpd::tc
