// probe.cu -- stand-alone bring-up probe for the tcgen05 building blocks used by libpgpd's
// tensor-core kernels: TMEM alloc, hand-written shared-memory operand layouts + UMMA descriptors,
// tcgen05.mma (kind::f16 / kind::tf32), tcgen05.commit -> mbarrier, tcgen05.ld epilogue, and the
// 3-pass hi/lo fp16 split that gives fp32-grade GEMM accuracy.  TEST TOOL (not part of the product).
//
//   ./tc_probe.bin <variant>     prints one line: "variant V <name>: PASS|FAIL max_err=... ref_max=..."
// Each variant runs in its own process so that a trapping descriptor does not take the others down.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

struct Variant {
    const char* name;
    int layout;        // 0 = K-major SWIZZLE_128B (8 x 128 B atoms), 1 = K-major no swizzle (8 x 16 B core matrices),
                       // 2 = MN-major SWIZZLE_128B: rows indexed by k, 128 B (64 elements of M/N) per row, one 64-wide atom after another
    int kind;          // 0 = f16 (fp16 operands), 1 = tf32 (fp32 operands)
    int split;         // 0 = single pass, 1 = hi/lo 3-pass
    uint32_t lbo, sbo; // bytes; 0xFFFFFFFF = derive from layout
    uint32_t layout_type, version;
    float scaleA;      // magnitude of A (tests fp16 subnormal lo parts)
    int N;             // MMA N (128 or 256)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type, uint32_t version) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)(version & 0x3) << 46;
    d |= (uint64_t)(layout_type & 0x7) << 61;
    return d;
}

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// returns false on timeout
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (int it = 0; it < (1 << 22); ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

// operand tile storage: rows x K elements.  returns byte offset of the 16-byte chunk `c16` (index along K) of row r.
__device__ __forceinline__ uint32_t chunk_off(int layout, int r, int c16, int chunks_per_kblock, int rows, int total_chunks) {
    if (layout == 0) {
        int kb = c16 / chunks_per_kblock, c = c16 % chunks_per_kblock;       // chunks_per_kblock == 8 (128 B rows)
        return (uint32_t)(kb * rows * 128 + r * 128 + ((c ^ (r & 7)) << 4));
    }
    // no swizzle: core matrix (8 rows x 16 B) contiguous; K-adjacent core matrices 128 B apart; 8-row groups total_chunks*128 B apart
    return (uint32_t)((r >> 3) * total_chunks * 128 + c16 * 128 + (r & 7) * 16);
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K, Variant v, int* status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NB = v.N;                       // rows of the B operand
    const int esize = v.kind == 0 ? 2 : 4;
    const int epc = 16 / esize;               // elements per 16-byte chunk
    const int total_chunks = K / epc;
    const uint32_t a_bytes = 128u * K * esize, b_bytes = (uint32_t)NB * K * esize;
    unsigned char* sA_hi = smem;
    unsigned char* sA_lo = sA_hi + a_bytes;
    unsigned char* sB_hi = sA_lo + a_bytes;
    unsigned char* sB_lo = sB_hi + b_bytes;

    // ---- fill operands (generic proxy writes) ----
    if (v.layout == 2) {
        // element (mn, k) of an operand with `rows` M/N rows: atom = mn/64 at atom*(K*128) ; row k at k*128 ; chunk = (mn%64)/8 ^ (k&7)
        for (int idx = tid; idx < (128 + NB) * K; idx += 128) {
            const bool isA = idx < 128 * K;
            const int li = isA ? idx : idx - 128 * K;
            const int mn = li / K, k = li % K;
            const float f = (isA ? A : B)[(size_t)mn * K + k];
            const __half h = __float2half_rn(f);
            const __half l = __float2half_rn(f - __half2float(h));
            const uint32_t off = (uint32_t)((mn >> 6) * (K * 128) + k * 128 + ((((mn & 63) >> 3) ^ (k & 7)) << 4) + (mn & 7) * 2);
            *reinterpret_cast<__half*>((isA ? sA_hi : sB_hi) + off) = h;
            *reinterpret_cast<__half*>((isA ? sA_lo : sB_lo) + off) = l;
        }
    } else
    for (int r = tid; r < 128 + NB; r += 128) {
        const bool isA = r < 128;
        const int row = isA ? r : r - 128;
        const int rows = isA ? 128 : NB;
        const float* src = (isA ? A : B) + (size_t)row * K;
        unsigned char* hi = isA ? sA_hi : sB_hi;
        unsigned char* lo = isA ? sA_lo : sB_lo;
        for (int c = 0; c < total_chunks; ++c) {
            uint32_t off = chunk_off(v.layout, row, c, 8, rows, total_chunks);
            if (v.kind == 0) {
                __half h[8], l[8];
                for (int e = 0; e < 8; ++e) {
                    float f = src[c * 8 + e];
                    h[e] = __float2half_rn(f);
                    l[e] = __float2half_rn(f - __half2float(h[e]));
                }
                *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<uint4*>(h);
                *reinterpret_cast<uint4*>(lo + off) = *reinterpret_cast<uint4*>(l);
            } else {
                float h[4], l[4];
                for (int e = 0; e < 4; ++e) {
                    float f = src[c * 4 + e];
                    uint32_t u = __float_as_uint(f);
                    uint32_t rb = ((u >> 13) & 1u) + 0x0FFFu;       // round-to-nearest-even to 10 mantissa bits
                    h[e] = __uint_as_float((u + rb) & ~0x1FFFu);
                    float d = f - h[e];
                    uint32_t ud = __float_as_uint(d);
                    uint32_t rd = ((ud >> 13) & 1u) + 0x0FFFu;
                    l[e] = __uint_as_float((ud + rd) & ~0x1FFFu);
                }
                *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<uint4*>(h);
                *reinterpret_cast<uint4*>(lo + off) = *reinterpret_cast<uint4*>(l);
            }
        }
    }
    // ---- barrier init + TMEM alloc ----
    const uint32_t bar = smem_u32(&mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // operand writes -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    // ---- MMA issue (one thread) ----
    if (tid == 0) {
        uint32_t idesc = (1u << 4);                                  // D = f32
        if (v.kind == 0) idesc |= (0u << 7) | (0u << 10);            // A,B = f16
        else idesc |= (2u << 7) | (2u << 10);                        // A,B = tf32
        idesc |= ((uint32_t)(v.N >> 3) << 17) | ((128u >> 4) << 24);
        if (v.layout == 2) idesc |= (1u << 15) | (1u << 16);         // A and B are MN-major
        const int kper = v.kind == 0 ? 16 : 8;                       // K per MMA
        const int nk = K / kper;
        const uint32_t lboA = v.lbo != 0xFFFFFFFFu ? v.lbo : (v.layout == 0 ? 0u : 128u);
        const uint32_t sboA = v.sbo != 0xFFFFFFFFu ? v.sbo : (v.layout == 0 ? 1024u : (uint32_t)total_chunks * 128u);
        const int npass = v.split ? 3 : 1;
        uint32_t acc = 0;
        for (int pass = 0; pass < npass; ++pass) {
            const unsigned char* pa = (pass == 1) ? sA_lo : sA_hi;
            const unsigned char* pb = (pass == 2) ? sB_lo : sB_hi;
            for (int ks = 0; ks < nk; ++ks) {
                uint32_t offA, offB;
                if (v.layout == 0) {
                    int kb = ks / 4, within = ks % 4;                 // 4 MMAs per 128-byte K block
                    offA = kb * 128 * 128 + within * 32;
                    offB = kb * NB * 128 + within * 32;
                } else if (v.layout == 2) {
                    offA = offB = ks * 16 * 128;                      // 16 k-rows of 128 B per MMA
                } else {
                    offA = offB = ks * 2 * 128;                       // two 16-byte core-matrix columns per MMA
                }
                uint64_t da = make_desc(smem_u32(pa) + offA, lboA, sboA, v.layout_type, v.version);
                uint64_t db = make_desc(smem_u32(pb) + offB, lboA, sboA, v.layout_type, v.version);
                if (v.kind == 0) mma_f16(tmem, da, db, idesc, acc); else mma_tf32(tmem, da, db, idesc, acc);
                acc = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
    }
    // ---- epilogue ----
    bool ok = mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok) { if (tid == 0) *status = -1; }
    else {
        for (int c0 = 0; c0 < v.N; c0 += 32) {
            uint32_t r[32];
            uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                           "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                           "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) D[(size_t)(warp * 32 + lane) * v.N + c0 + j] = __uint_as_float(r[j]);
        }
        if (tid == 0) *status = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256));
}

static float lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 32768.0f - 1.0f; }

static double q16(double x) { return (double)__half2float(__float2half_rn((float)x)); }
static double qtf32(double x) { float f = (float)x; uint32_t u; memcpy(&u, &f, 4); uint32_t rb = ((u >> 13) & 1u) + 0x0FFFu; u = (u + rb) & ~0x1FFFu; memcpy(&f, &u, 4); return f; }

int main(int argc, char** argv) {
    const uint32_t X = 0xFFFFFFFFu;
    Variant vs[] = {
        {"sw128 f16 plain  lbo0 sbo1024 v1", 0, 0, 0, 0, 1024, 2, 1, 1.0f, 128},
        {"sw128 f16 plain  lbo16 sbo1024 v1", 0, 0, 0, 16, 1024, 2, 1, 1.0f, 128},
        {"nosw  f16 plain  lbo128 sboK v1", 1, 0, 0, X, X, 0, 1, 1.0f, 128},
        {"nosw  f16 plain  swapped lbo/sbo v1", 1, 0, 0, 0xFFFFFFFEu, 0xFFFFFFFEu, 0, 1, 1.0f, 128},
        {"sw128 f16 3pass  v1", 0, 0, 1, 0, 1024, 2, 1, 1.0f, 128},
        {"sw128 f16 3pass small A", 0, 0, 1, 0, 1024, 2, 1, 1e-3f, 128},
        {"sw128 tf32 plain v1", 0, 1, 0, 0, 1024, 2, 1, 1.0f, 128},
        {"sw128 tf32 3pass v1", 0, 1, 1, 0, 1024, 2, 1, 1.0f, 128},
        {"sw128 f16 plain  N=256", 0, 0, 0, 0, 1024, 2, 1, 1.0f, 256},
        {"sw128 f16 plain  version0", 0, 0, 0, 0, 1024, 2, 0, 1.0f, 128},
        {"nosw  f16 3pass", 1, 0, 1, X, X, 0, 1, 1.0f, 128},
        {"MN-major sw128 f16 plain lbo=K*128 sbo=1024", 2, 0, 0, 16384, 1024, 2, 1, 1.0f, 128},
        {"MN-major sw128 f16 plain lbo=1024 sbo=K*128", 2, 0, 0, 1024, 16384, 2, 1, 1.0f, 128},
        {"MN-major sw128 f16 3pass lbo=K*128 sbo=1024", 2, 0, 1, 16384, 1024, 2, 1, 1.0f, 128},
    };
    const int nv = (int)(sizeof(vs) / sizeof(vs[0]));
    if (argc < 2) { printf("%d\n", nv); return 0; }
    int vi = atoi(argv[1]);
    if (vi < 0 || vi >= nv) return 2;
    Variant v = vs[vi];
    const int K = (v.kind == 1) ? 64 : 128, M = 128, N = v.N;
    if (v.lbo == 0xFFFFFFFEu) {                 // "swapped" variant of the no-swizzle layout
        int esize = v.kind == 0 ? 2 : 4;
        int total_chunks = K / (16 / esize);
        v.lbo = (uint32_t)total_chunks * 128u; v.sbo = 128u;
    }
    std::vector<float> hA((size_t)M * K), hB((size_t)N * K), hD((size_t)M * N, 0.f);
    uint32_t seed = 12345u + vi;
    for (auto& x : hA) x = v.scaleA * lcg(seed);
    for (auto& x : hB) x = lcg(seed);
    float *dA, *dB, *dD; int* dS;
    cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dD, hD.size() * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, hD.size() * 4); cudaMemset(dS, 0, 4);
    int esize = v.kind == 0 ? 2 : 4;
    size_t smem = 2 * (size_t)(M + N) * K * esize + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, K, v, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d %s: CUDA_ERROR %s\n", vi, v.name, cudaGetErrorString(e)); return 1; }
    int st = 0;
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    if (st != 1) { printf("variant %d %s: TIMEOUT status=%d\n", vi, v.name, st); return 1; }
    double max_err = 0, ref_max = 0, max_err_exact = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0, exact = 0;
            for (int k = 0; k < K; ++k) {
                double a = hA[(size_t)m * K + k], b = hB[(size_t)n * K + k];
                exact += a * b;
                if (v.split) ref += a * b;
                else ref += (v.kind == 0 ? q16(a) * q16(b) : qtf32(a) * qtf32(b));
            }
            max_err = fmax(max_err, fabs(hD[(size_t)m * N + n] - ref));
            max_err_exact = fmax(max_err_exact, fabs(hD[(size_t)m * N + n] - exact));
            ref_max = fmax(ref_max, fabs(ref));
        }
    double tol = (v.split ? 2e-6 : 2e-5) * ref_max;
    printf("variant %d %s: %s max_err=%.3e (vs exact %.3e) ref_max=%.3e\n", vi, v.name, max_err <= tol ? "PASS" : "FAIL", max_err, max_err_exact, ref_max);
    return max_err <= tol ? 0 : 1;
}
