// Stand-alone probe for the tcgen05 (UMMA) building blocks used by the TF32 learn path:
// shared-memory matrix descriptors (K-major / MN-major, 128-byte swizzle), the kind::tf32
// instruction descriptor, TMEM alloc / ld, commit -> mbarrier, and 3xTF32 error compensation.
// One CTA computes D[128][N] = A[128][K] * B with fp32 operands.  Built by tools/probe.py.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    d |= (uint64_t)layout << 61;   // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B (tf32 MN-major)
    return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}

// hi = x rounded to tf32 (round-to-nearest, ties away), lo = x - hi (exact in fp32)
__device__ __forceinline__ float hi_part(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float lo_part(float x) { return x - hi_part(x); }
// three_pass == 2: the raw fp32 word is the "hi" operand (the tensor core drops the low 13 mantissa bits itself)
// and lo = x - truncate(x)
__device__ __forceinline__ float lo_trunc(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
// K-major SW64 tile: rows x 16-float atoms (64-byte rows, 8-row groups of 512 B); element (r, k)
__device__ __forceinline__ uint32_t off_kmajor64(int rows, int r, int k) {
    return (uint32_t)((k >> 4) * rows * 64 + r * 64 + ((((k & 15) >> 2) ^ ((r >> 1) & 3)) << 4) + ((k & 3) << 2));
}

// K-major SW128 tile: rows x 32-float atoms; element (r, k)
__device__ __forceinline__ uint32_t off_kmajor(int rows, int r, int k) {
    return (uint32_t)((k >> 5) * rows * 128 + r * 128 + ((((k & 31) >> 2) ^ (r & 7)) << 4) + ((k & 3) << 2));
}
// MN-major tf32 tile [K][N] ("SWIZZLE_128B_BASE32B"): atoms of 4 k x 32 n (512 B), rows of 128 B,
// 32-byte pieces of a row XOR-ed with the row index; element (k, n)
__device__ __forceinline__ uint32_t off_mnmajor(int n_total, int k, int n) {
    return (uint32_t)((k >> 2) * (n_total >> 5) * 512 + (n >> 5) * 512 + (k & 3) * 128 +
                      ((((n & 31) >> 3) ^ (k & 3)) << 5) + ((n & 7) << 2));
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K, int N,
             int b_mode, int three_pass, int a_mode, int* status, int reps) {
    const int b_mn = b_mode == 1;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t a_bytes = 128 * K * 4, b_bytes = (uint32_t)N * K * 4;
    uint8_t* sA = smem;
    uint8_t* sAl = sA + a_bytes;
    uint8_t* sB = sAl + a_bytes;
    uint8_t* sBl = sB + b_bytes;

    for (int i = tid; i < 128 * K; i += 128) {
        const int r = a_mode == 1 ? i % 128 : i / K, k = a_mode == 1 ? i / 128 : i % K;
        const float v = A[i];
        const uint32_t o = a_mode == 1 ? off_mnmajor(128, k, r) : off_kmajor(128, r, k);
        *reinterpret_cast<float*>(sA + o) = three_pass == 1 ? hi_part(v) : v;
        *reinterpret_cast<float*>(sAl + o) = three_pass == 2 ? lo_trunc(v) : lo_part(v);
    }
    for (int i = tid; i < N * K; i += 128) {
        float v = B[i];
        uint32_t o;
        if (b_mn) { const int k = i / N, n = i % N; o = off_mnmajor(N, k, n); }
        else      { const int n = i / K, k = i % K; o = b_mode == 2 ? off_kmajor64(N, n, k) : off_kmajor(N, n, k); }
        *reinterpret_cast<float*>(sB + o) = three_pass == 1 ? hi_part(v) : v;
        *reinterpret_cast<float*>(sBl + o) = three_pass == 2 ? lo_trunc(v) : lo_part(v);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (UMMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;

    if (a_mode == 2) {   // A_lo[row][k] -> TMEM lane row, column 256 + k (thread = row)
        for (int k0 = 0; k0 < K; k0 += 8) {
            uint32_t v[8];
            for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(three_pass == 2 ? lo_trunc(A[(size_t)tid * K + k0 + j]) : lo_part(A[(size_t)tid * K + k0 + j]));
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 256u + (uint32_t)k0;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
                         "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mode == 1 ? 1 : 0) << 15) |
                               ((uint32_t)(b_mn ? 1 : 0) << 16) |
                               ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        uint32_t accum = 0;
        const int passes = three_pass ? 3 : 1;
        const long long t0 = clock64();
        for (int rep = 0; rep < reps; ++rep)
        for (int p = 0; p < passes; ++p) {
            // small terms first: A_lo*B, A*B_lo, then A*B (the hardware truncates to tf32 itself)
            const uint8_t* a_src = (three_pass && p == 0) ? sAl : sA;
            const uint8_t* b_src = (three_pass && p == 1) ? sBl : sB;
            for (int k = 0; k < K; k += 8) {
                uint64_t adesc;
                if (a_mode == 1) adesc = make_desc(smem_u32(a_src) + (k >> 2) * 4 * 512, 512, 4 * 512, 1);
                else adesc = make_desc(smem_u32(a_src) + (k >> 5) * (128 * 128) + ((k & 31) >> 3) * 32, 16, 1024);
                uint64_t bdesc;
                if (b_mn) bdesc = make_desc(smem_u32(b_src) + (k >> 2) * (N >> 5) * 512, 512, (N >> 5) * 512, 1);
                else if (b_mode == 2) bdesc = make_desc(smem_u32(b_src) + (k >> 4) * (N * 64) + ((k & 15) >> 3) * 32, 16, 512, 4);
                else      bdesc = make_desc(smem_u32(b_src) + (k >> 5) * (N * 128) + ((k & 31) >> 3) * 32, 16, 1024);
                if (a_mode == 2 && three_pass && p == 0) mma_tf32_ts(tmem, tmem + 256u + (uint32_t)k, bdesc, idesc, accum);
                else mma_tf32(tmem, adesc, bdesc, idesc, accum);
                accum = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        uint32_t d2 = 0;
        for (int spin = 0; spin < (1 << 24) && !d2; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(d2) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
        status[1] = (int)(clock64() - t0);
        status[2] = reps * passes * (K / 8);
        // issue-rate test: 8 precomputed descriptor pairs, MMAs back to back
        uint64_t ad[8], bd[8];
        for (int q = 0; q < 8; ++q) {
            const int k = (q * 8) % K;
            ad[q] = a_mode == 1 ? make_desc(smem_u32(sA) + (k >> 2) * 4 * 512, 512, 4 * 512, 1)
                                : make_desc(smem_u32(sA) + (k >> 5) * (128 * 128) + ((k & 31) >> 3) * 32, 16, 1024);
            if (b_mn) bd[q] = make_desc(smem_u32(sB) + (k >> 2) * (N >> 5) * 512, 512, (N >> 5) * 512, 1);
            else if (b_mode == 2) bd[q] = make_desc(smem_u32(sB) + (k >> 4) * (N * 64) + ((k & 15) >> 3) * 32, 16, 512, 4);
            else bd[q] = make_desc(smem_u32(sB) + (k >> 5) * (N * 128) + ((k & 31) >> 3) * 32, 16, 1024);
        }
        const long long t1 = clock64();
        for (int rep = 0; rep < reps * 4; ++rep) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (a_mode == 2) mma_tf32_ts(tmem, tmem + 256u + (uint32_t)(q * 8 % K), bd[q], idesc, 1u);
                else mma_tf32(tmem, ad[q], bd[q], idesc, 1u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        uint32_t d3 = 0;
        for (int spin = 0; spin < (1 << 24) && !d3; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(d3) : "r"(smem_u32(&mbar)), "r"(1u) : "memory");
        status[3] = (int)(clock64() - t1);
    }
    // bounded wait (a broken descriptor must not hang the GPU)
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
    if (!done && tid == 0) *status = 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (done) {
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
                "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int row = warp * 32 + lane;
            for (int j = 0; j < 32 && c0 + j < N; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace

extern "C" int umma_probe(const float* A, const float* B, float* D, int K, int N, int b_mode, int three_pass, int a_mode,
                          int* status, int reps, void* stream) {
    const size_t smem = 2 * (size_t)(128 * K * 4) + 2 * (size_t)N * K * 4 + 1024;
    cudaError_t e = cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -2;
    probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, K, N, b_mode, three_pass, a_mode, status, reps);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
