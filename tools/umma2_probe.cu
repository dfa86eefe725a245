// Stand-alone probe for the 2-CTA tcgen05 path (cta_group::2) planned for K3 / K4a:
// a cluster of two CTAs computes D[256][256] = A[256][K] * B[K][256] in kind::tf32 with
//   * each CTA staging ITS 128 rows of A (K-major SW128) and ITS 128 columns of B (MN-major, 128B_BASE32B),
//   * tcgen05.alloc.cta_group::2 in both CTAs, one tcgen05.mma.cta_group::2 per k-step issued by the leader only,
//   * the peer CTA telling the leader "my operands are staged" with a remote mbarrier arrive (mapa + shared::cluster),
//   * tcgen05.commit.cta_group::2 ... multicast::cluster releasing the epilogue of both CTAs,
//   * each CTA reading its own 128 accumulator rows back with tcgen05.ld.
// Optional: A from TMEM (a_tmem = 1): each CTA writes its A rows into its own TMEM columns 256.. with tcgen05.st.
// Built and run by tools/probe2.py.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16 |
           (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32 | (uint64_t)1 << 46 | (uint64_t)layout << 61;
}
__device__ __forceinline__ uint32_t off_k128(int rows, int r, int k) {
    return (uint32_t)((k >> 5) * rows * 128 + r * 128 + ((((k & 31) >> 2) ^ (r & 7)) << 4) + ((k & 3) << 2));
}
__device__ __forceinline__ uint32_t off_mn(int mn_total, int k, int mn) {
    return (uint32_t)((k >> 2) * (mn_total >> 5) * 512 + (mn >> 5) * 512 + (k & 3) * 128 +
                      ((((mn & 31) >> 3) ^ (k & 3)) << 5) + ((mn & 7) << 2));
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 22) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe2_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K, int a_tmem, int* status) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full, bar_done;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sp = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t a_bytes = 128u * K * 4, sA = sbase, sB = sbase + a_bytes;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_full)), "r"(2));   // leader thread + peer thread
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_done)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    // my 128 rows of A (K-major SW128) and my 128 columns of B (MN-major, 128 columns wide)
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<float*>(sp + off_k128(128, r, k)) = A[(size_t)(rank * 128 + r) * K + k];
    }
    for (int i = tid; i < K * 128; i += 128) {
        const int k = i / 128, n = i % 128;
        *reinterpret_cast<float*>(sp + a_bytes + off_mn(128, k, n)) = B[(size_t)k * 256 + rank * 128 + n];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    if (a_tmem) {      // A also into TMEM columns 256..: thread = row, 32 columns per tcgen05.st
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        for (int c0 = 0; c0 < K; c0 += 32) {
            uint32_t v[32];
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(A[(size_t)(rank * 128 + tid) * K + c0 + j]);
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
                "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(tmem + lane_addr + 256u + (uint32_t)c0),
                "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
                "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
                "r"(v[31]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    cluster_sync();       // barriers initialised and TMEM allocated in both CTAs before anything crosses the pair

    // "my operands are staged": one arrival from each CTA on the LEADER's full barrier
    if (tid == 0) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&bar_full)), "r"(0));
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
    }
    bool ok = true;
    if (rank == 0 && tid == 0) {
        ok = mbar_wait(smem_u32(&bar_full), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // kind::tf32, fp32 accumulate, A K-major, B MN-major, M = 256 (both CTAs), N = 256
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        for (int k0 = 0; k0 < K; k0 += 8) {
            const uint64_t ad = make_desc(sA + (uint32_t)(k0 >> 5) * 128 * 128 + (uint32_t)((k0 & 31) >> 3) * 32, 16, 1024, 2);
            const uint64_t bd = make_desc(sB + (uint32_t)(k0 >> 2) * 4 * 512, 512, 4 * 512, 1);     // 4 n-atoms per k-group
            const uint32_t acc = k0 ? 1u : 0u;
            if (a_tmem)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                             ::"r"(tmem), "r"(tmem + 256u + (uint32_t)k0), "l"(bd), "r"(idesc), "r"(acc) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                             ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar_done)), "h"((uint16_t)3) : "memory");
    }
    const bool got = mbar_wait(smem_u32(&bar_done), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!got || !ok) atomicExch(status, 1 + (int)rank);
    // my 128 rows: thread = row, 32 columns per tcgen05.ld
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
            "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(tmem + lane_addr + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[(size_t)(rank * 128 + tid) * 256 + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();       // both CTAs are done with the pair's TMEM
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace

extern "C" int umma2_probe(const float* A, const float* B, float* D, int K, int a_tmem, int* status, void* stream) {
    const size_t smem = (size_t)(128 * K * 4) * 2 + 2048;
    cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe2_kernel<<<2, 128, smem, (cudaStream_t)stream>>>(A, B, D, K, a_tmem, status);
    return (int)cudaGetLastError();
}
