// Probe: how does the access pattern of K4b's Adam epilogue (a warp instruction = R rows x W bytes, rows 1 KB apart) affect the
// achieved read-modify-write bandwidth on theta / m / v?  148 persistent CTAs x 256 threads walk 128x256 tiles like K4b does.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rmw_pattern_probe tools/rmw_pattern_probe.cu && ./rmw_pattern_probe
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ float4 ldcg(const float* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void upd(float4& t, float4& m, float4& v) {
    m.x = m.x * 0.9f + 1e-3f; m.y = m.y * 0.9f + 1e-3f; m.z = m.z * 0.9f + 1e-3f; m.w = m.w * 0.9f + 1e-3f;
    v.x = v.x * 0.999f + 1e-6f; v.y = v.y * 0.999f + 1e-6f; v.z = v.z * 0.999f + 1e-6f; v.w = v.w * 0.999f + 1e-6f;
    t.x -= m.x * rsqrtf(v.x + 1e-7f); t.y -= m.y * rsqrtf(v.y + 1e-7f); t.z -= m.z * rsqrtf(v.z + 1e-7f); t.w -= m.w * rsqrtf(v.w + 1e-7f);
}
// LPR = lanes per row segment (8: 128 B, 16: 256 B, 32: 512 B).  A tile is 128 rows x 256 floats (1 KB per row).
// warp w: row quarter w & 3 (32 rows), column half w >> 2 (128 floats).  Per step the warp covers (32 / LPR) rows x (LPR * 4) floats;
// NB loads batches of NB row groups before computing.
template <int LPR, int NB>
__global__ void __launch_bounds__(256, 1) tile_rmw(float* th, float* am, float* av, int n_tiles) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rq = warp & 3, half = warp >> 2;
    constexpr int RPI = 32 / LPR;                // rows per instruction
    const int rsub = lane / LPR, c4 = (lane % LPR) * 4;
    constexpr int COLS = LPR * 4;                // floats per row segment
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const size_t base = (size_t)t * 128 * 256 + (size_t)(rq * 32) * 256 + half * 128;
        for (int cb = 0; cb < 128 / COLS; ++cb) {
            for (int r0 = 0; r0 < 32; r0 += RPI * NB) {
                float4 t4[NB], m4[NB], v4[NB];
                size_t off[NB];
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    off[u] = base + (size_t)(r0 + u * RPI + rsub) * 256 + cb * COLS + c4;
                    t4[u] = ldcg(th + off[u]); m4[u] = ldcg(am + off[u]); v4[u] = ldcg(av + off[u]);
                }
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    upd(t4[u], m4[u], v4[u]);
                    *reinterpret_cast<float4*>(th + off[u]) = t4[u];
                    *reinterpret_cast<float4*>(am + off[u]) = m4[u];
                    *reinterpret_cast<float4*>(av + off[u]) = v4[u];
                }
            }
        }
    }
}
// 16 warps: row quarter w & 3, column quarter w >> 2 (64 floats); LPR lanes per row segment, NB row groups per batch
template <int LPR, int NB>
__global__ void __launch_bounds__(512, 1) tile_rmw16(float* th, float* am, float* av, int n_tiles) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rq = warp & 3, quarter = warp >> 2;
    constexpr int RPI = 32 / LPR;
    const int rsub = lane / LPR, c4 = (lane % LPR) * 4;
    constexpr int COLS = LPR * 4;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const size_t base = (size_t)t * 128 * 256 + (size_t)(rq * 32) * 256 + quarter * 64;
        for (int cb = 0; cb < 64 / COLS; ++cb) {
            for (int r0 = 0; r0 < 32; r0 += RPI * NB) {
                float4 t4[NB], m4[NB], v4[NB];
                size_t off[NB];
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    off[u] = base + (size_t)(r0 + u * RPI + rsub) * 256 + cb * COLS + c4;
                    t4[u] = ldcg(th + off[u]); m4[u] = ldcg(am + off[u]); v4[u] = ldcg(av + off[u]);
                }
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    upd(t4[u], m4[u], v4[u]);
                    *reinterpret_cast<float4*>(th + off[u]) = t4[u];
                    *reinterpret_cast<float4*>(am + off[u]) = m4[u];
                    *reinterpret_cast<float4*>(av + off[u]) = v4[u];
                }
            }
        }
    }
}
// The same walk with K4b's real per-element work: gradient from a per-warp shared-memory tile (written as a "TMEM row" per
// lane, read back as 16-byte row pieces, like the epilogue's transpose) and the real Adam arithmetic (MUFU sqrt + divide).
__device__ __forceinline__ void adam_real(float g, float& th, float& m, float& v) {
    m = m + (g - m) * 0.1f;
    v = v + (g * g - v) * 0.001f;
    float sq;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(v));
    th = th - __fdividef(m * 5e-4f, sq + 1e-7f);
}
template <int WARPS, int LPR, int NB, int CW>   // CW = columns of a block (32 or 16); a warp owns 256 / (WARPS / 4) columns
__global__ void __launch_bounds__(WARPS * 32, 1) tile_real(float* th, float* am, float* av, int n_tiles) {
    constexpr int TLD = CW + 4;
    extern __shared__ __align__(16) float tiles_dyn[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rq = warp & 3, part = warp >> 2;
    constexpr int WCOLS = 256 / (WARPS / 4);
    constexpr int RPI = 32 / LPR;
    static_assert(LPR * 4 == CW, "a row segment is one block wide");
    const int rsub = lane / LPR, c4 = (lane % LPR) * 4;
    float* tile = tiles_dyn + warp * 32 * TLD;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const size_t base = (size_t)t * 128 * 256 + (size_t)(rq * 32) * 256 + part * WCOLS;
        for (int cb = 0; cb < WCOLS / CW; ++cb) {
#pragma unroll
            for (int j = 0; j < CW; j += 4)      // "accumulator row" of this lane -> tile
                *reinterpret_cast<float4*>(tile + lane * TLD + j) = make_float4(1e-3f * j, 2e-3f, 3e-3f, 1e-3f * lane);
            __syncwarp();
            for (int r0 = 0; r0 < 32; r0 += RPI * NB) {
                float4 t4[NB], m4[NB], v4[NB];
                size_t off[NB];
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    off[u] = base + (size_t)(r0 + u * RPI + rsub) * 256 + cb * CW + c4;
                    t4[u] = ldcg(th + off[u]); m4[u] = ldcg(am + off[u]); v4[u] = ldcg(av + off[u]);
                }
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    const float4 g = *reinterpret_cast<const float4*>(tile + (r0 + u * RPI + rsub) * TLD + c4);
                    adam_real(g.x, t4[u].x, m4[u].x, v4[u].x); adam_real(g.y, t4[u].y, m4[u].y, v4[u].y);
                    adam_real(g.z, t4[u].z, m4[u].z, v4[u].z); adam_real(g.w, t4[u].w, m4[u].w, v4[u].w);
                    *reinterpret_cast<float4*>(th + off[u]) = t4[u];
                    *reinterpret_cast<float4*>(am + off[u]) = m4[u];
                    *reinterpret_cast<float4*>(av + off[u]) = v4[u];
                }
            }
            __syncwarp();
        }
    }
}
__global__ void stream_rmw(float* th, float* am, float* av, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 t = ldcg(th + 4 * i), m = ldcg(am + 4 * i), v = ldcg(av + 4 * i);
        upd(t, m, v);
        *reinterpret_cast<float4*>(th + 4 * i) = t; *reinterpret_cast<float4*>(am + 4 * i) = m; *reinterpret_cast<float4*>(av + 4 * i) = v;
    }
}
template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < 20; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 20;
}
int main() {
    const int n_tiles = 768 - 768 % 1;          // cfg3: 256 networks x (2 dW2 tiles + ~0.75 dW1 tile)
    const size_t n = (size_t)n_tiles * 128 * 256;
    float *th, *am, *av;
    CK(cudaMalloc(&th, n * 4)); CK(cudaMalloc(&am, n * 4)); CK(cudaMalloc(&av, n * 4));
    CK(cudaMemset(th, 0, n * 4)); CK(cudaMemset(am, 0, n * 4)); CK(cudaMemset(av, 0, n * 4));
    const double bytes = 24.0 * n;
    auto rep = [&](const char* name, float ms) { printf("%-46s %7.1f us  %6.0f GB/s\n", name, ms * 1e3, bytes / ms / 1e6); };
    rep("streaming, 4736 x 256 threads", timeit([&] { stream_rmw<<<148 * 32, 256>>>(th, am, av, n / 4); }));
    rep("streaming, 148 x 256 threads", timeit([&] { stream_rmw<<<148, 256>>>(th, am, av, n / 4); }));
    rep("streaming, 148 x 1024 threads", timeit([&] { stream_rmw<<<148, 1024>>>(th, am, av, n / 4); }));
    rep("tiles, 4 rows x 128 B per instr, batch 4", timeit([&] { tile_rmw<8, 4><<<148, 256>>>(th, am, av, n_tiles); }));
    rep("tiles, 4 rows x 128 B per instr, batch 8", timeit([&] { tile_rmw<8, 8><<<148, 256>>>(th, am, av, n_tiles); }));
    rep("tiles, 2 rows x 256 B per instr, batch 8", timeit([&] { tile_rmw<16, 8><<<148, 256>>>(th, am, av, n_tiles); }));
    rep("tiles, 1 row x 512 B per instr, batch 8", timeit([&] { tile_rmw<32, 8><<<148, 256>>>(th, am, av, n_tiles); }));
    rep("tiles, 1 row x 512 B per instr, batch 16", timeit([&] { tile_rmw<32, 16><<<148, 256>>>(th, am, av, n_tiles); }));
    rep("16 warps, 8 rows x 64 B per instr, batch 4", timeit([&] { tile_rmw16<4, 4><<<148, 512>>>(th, am, av, n_tiles); }));
    rep("16 warps, 4 rows x 128 B per instr, batch 4", timeit([&] { tile_rmw16<8, 4><<<148, 512>>>(th, am, av, n_tiles); }));
    rep("16 warps, 4 rows x 128 B per instr, batch 8", timeit([&] { tile_rmw16<8, 8><<<148, 512>>>(th, am, av, n_tiles); }));
    rep("16 warps, 2 rows x 256 B per instr, batch 4", timeit([&] { tile_rmw16<16, 4><<<148, 512>>>(th, am, av, n_tiles); }));
    rep("REAL math+tile,  8 warps, 4 x 128 B, batch 8", timeit([&] { cudaFuncSetAttribute(tile_real<8, 8, 8, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 36864); tile_real<8, 8, 8, 32><<<148, 256, 36864>>>(th, am, av, n_tiles); }));
    rep("REAL math+tile, 16 warps, 4 x 128 B, batch 4", timeit([&] { cudaFuncSetAttribute(tile_real<16, 8, 4, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 73728); tile_real<16, 8, 4, 32><<<148, 512, 73728>>>(th, am, av, n_tiles); }));
    rep("REAL math+tile, 16 warps, 4 x 128 B, batch 8", timeit([&] { cudaFuncSetAttribute(tile_real<16, 8, 8, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 73728); tile_real<16, 8, 8, 32><<<148, 512, 73728>>>(th, am, av, n_tiles); }));
    rep("REAL math+tile, 16 warps, 8 x 64 B, batch 4", timeit([&] { cudaFuncSetAttribute(tile_real<16, 4, 4, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960); tile_real<16, 4, 4, 16><<<148, 512, 40960>>>(th, am, av, n_tiles); }));
    rep("REAL math+tile, 12 warps, 4 x 128 B, batch 8", timeit([&] { cudaFuncSetAttribute(tile_real<12, 8, 8, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 55296); tile_real<12, 8, 8, 32><<<148, 384, 55296>>>(th, am, av, n_tiles); }));
    CK(cudaDeviceSynchronize());
    return 0;
}
