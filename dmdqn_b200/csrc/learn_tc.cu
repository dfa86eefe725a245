// K3 / K4 on the 5th-generation tensor cores (tcgen05, kind::tf32, accumulators in TMEM).
//
// Same arithmetic contract as learn.cu (reference src/agents/dqn_agent.py:328-380), H = 256:
//   tc_target_kernel   K3   s' -> online(s'), target(s') -> TD target
//   tc_online_kernel   K4a  s  -> online(s), loss gradient, dh2, dh1 = (dh2 W2^T) * relu'(h1)
//   tc_wgrad_kernel    K4b  dW2 = h1^T dh2, dW1 = s^T dh1 with Adam + target sync in the epilogue
//
// tcgen05 has no fp32 MMA kind, so every fp32 operand x is split as x = hi + lo with
// hi = x rounded to tf32 and lo = x - hi (exact), and a product A*B is issued as three MMAs
// A_lo*B_hi + A_hi*B_lo + A_hi*B_hi accumulated in fp32 in TMEM ("3xTF32"; measured 3e-7 relative
// on B200, tools/umma_probe.cu).  PASSES = 1 issues only A*B (plain TF32, ~1e-3, toleranced apart).
//
// A CTA owns 128 batch rows of one network (UMMA M = 128, N = 256 = all output columns):
//   * activations: hi part in shared memory as the next layer's A operand (K-major, 128-byte
//     swizzle), lo part in TMEM columns 256..511 (A-from-TMEM MMA) -- 128 KB each, so the full
//     512 TMEM columns are used: 256 accumulator + 256 operand;
//   * weights: 8 x 256 (x.W) or 256 x 16 (d.W^T) chunks streamed from L2 through registers (loads two own
//     chunks ahead) into the UMMA canonical layout (MN-major "128B_BASE32B" / K-major 64-byte swizzle),
//     split hi/lo by the thread that loaded them; a full/empty mbarrier ring (4 / 2 stages, one producer
//     group per stage) couples the 8 producer warps to the single MMA-issuing lane (warp 8),
//     tcgen05.commit frees a stage;
//   * epilogues read the accumulator with tcgen05.ld (thread = batch row), so bias/ReLU, the
//     4-wide head (layer 3), TD target, loss gradient and dh2 are computed per row in registers.
#include "common.cuh"

namespace dmdqn {

namespace {

constexpr int H = 256;          // hidden width served by this path
constexpr int BM = 128;         // batch rows per CTA (UMMA M)
constexpr int NT = 256;         // producer / epilogue threads: 8 warps, two per TMEM sub-partition
constexpr int NT_F = NT + 32;   // K3 / K4a add one warp whose lane 0 only issues tcgen05.mma
constexpr int KC = 16;          // k-rows of the streamed operand per stage
constexpr uint32_t ATOM = BM * 128;           // bytes of one K-major SW128 atom column (128 rows x 32 floats)
constexpr uint32_t STAGE = 2 * KC * H * 4;    // hi + lo of a 16 x 256 chunk
constexpr uint32_t SPIN_LIMIT = 1u << 20;

// ------------------------------------------------------------------------------------------
// PTX wrappers (syntax as in the CUTLASS sm100 headers shipped with the image).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 1 = SWIZZLE_128B_BASE32B (the only MN-major layout for tf32)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16 |
           (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32 | (uint64_t)1 << 46 | (uint64_t)layout << 61;
}
// kind::tf32, fp32 accumulate, M = 128, N = 256
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(H >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc),
                 "r"(accum) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(b),
                 "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// Bounded wait: a wrong descriptor must end in an error flag, not in a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < SPIN_LIMIT && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");   // suspend-time hint (ns): sleep instead of re-polling
    return done != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void prod_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 producer warps only
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
        "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
        "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// x = hi + lo, hi = x rounded to tf32 (nearest, ties away from zero), lo = x - hi exact in fp32.
// The rounding is an integer add on the bit pattern (the carry walks into the exponent like the magnitude
// does) followed by the mask: two instructions where cvt.rna.tf32 expands to four on sm_100a (its inf/nan
// guard is dropped: a non-finite activation poisons the step either way).  Plain truncation (one LOP3)
// was tried and rejected: |lo| doubles and, worse, every truncated lo errs toward zero, so the error of a
// K = 512 gradient sum grows linearly instead of as a random walk (1.1e-5 of the tensor scale at B = 512).
__device__ __forceinline__ float tf32_hi(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
template <int PASSES>
__device__ __forceinline__ void split4(const float4& x, float4& hi, float4& lo) {
    if (PASSES == 1) { hi = x; lo = make_float4(0.f, 0.f, 0.f, 0.f); return; }   // the tensor core truncates itself
    hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
    lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
}

// Byte offsets inside UMMA canonical tiles (validated by tools/umma_probe.cu).
//   K-major, 128-byte swizzle: [k/32][row][128 B], 16-byte piece index ^= row % 8
__device__ __forceinline__ uint32_t off_k128(int rows, int r, int k) {
    return (uint32_t)((k >> 5) * rows * 128 + r * 128 + ((((k & 31) >> 2) ^ (r & 7)) << 4) + ((k & 3) << 2));
}
//   K-major, 64-byte swizzle: [k/16][row][64 B], piece index ^= (row / 2) % 4
__device__ __forceinline__ uint32_t off_k64(int rows, int r, int k) {
    return (uint32_t)((k >> 4) * rows * 64 + r * 64 + ((((k & 15) >> 2) ^ ((r >> 1) & 3)) << 4) + ((k & 3) << 2));
}
//   MN-major tf32 (128B_BASE32B): atoms of 4 k x 32 mn = 512 B, [k/4][mn/32][k%4][128 B], 32-byte piece ^= k % 4
__device__ __forceinline__ uint32_t off_mn(int mn_total, int k, int mn) {
    return (uint32_t)((k >> 2) * (mn_total >> 5) * 512 + (mn >> 5) * 512 + (k & 3) * 128 +
                      ((((mn & 31) >> 3) ^ (k & 3)) << 5) + ((mn & 7) << 2));
}

struct TcArgs {
    dmdqn_dims d;
    Layout L;
    dmdqn_replay rp;
    dmdqn_nets nets;
    float gamma;
    int loss, double_dqn, adam_form, freq, loss_batch, tiles;
    double lr, beta1, beta2, adam_eps, tau;
    const int32_t *rows, *act_b, *active, *step_t;
    const float *r_hat, *done_b;
    const float4* adam_sc;
    uint32_t* mask2;              // relu'(h2) bits [n_nets][B][H/32]: the eight words of a batch row are adjacent
    float2* ga;                   // [n_nets][B] {dL/dq of the taken action, action as int bits}: what K4b needs per row
    float* w3_copy;               // [n_nets][H][4]
    float *y, *gcoef, *q_all, *q_next, *tq_all, *h1, *dh1, *dh2, *part_loss, *part_b3, *part_w3, *part_b2, *metrics, *grads;
    int* error;
};

// Phase stamps (build with -DTC_TIMING, tools/phase_timing.py): thread 0 of a few CTAs prints clock64 deltas.
#ifndef TC_EXP
#define TC_EXP 0
#endif
#ifndef WG_EXP
#define WG_EXP 0      // K4b experiment bits (profiling aid; results are wrong by design)
#endif
#ifdef TC_TIMING
#define TS_DECL long long ts_[16]; int tsi_ = 0
#define TS() do { if (tsi_ < 16) ts_[tsi_++] = clock64(); } while (0)
#define TS_D(i) (i < tsi_ ? ts_[i] - ts_[i - 1] : 0LL)
#define TS_PRINT(name) do { if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 301 || blockIdx.x == gridDim.x - 1)) \
        printf("%s cta %d: %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld | total %lld\n", name, blockIdx.x, TS_D(1), TS_D(2), TS_D(3), \
               TS_D(4), TS_D(5), TS_D(6), TS_D(7), TS_D(8), TS_D(9), TS_D(10), ts_[tsi_ - 1] - ts_[0]); } while (0)
#else
#define TS_DECL
#define TS()
#define TS_PRINT(name)
#endif

#ifdef TC_TIMING
#define WG_TS_DECL long long wts_[20]; int wtsi_ = 0
#define WG_TS() do { if (wtsi_ < 20) wts_[wtsi_++] = clock64(); } while (0)
#define WG_D(i) ((i) < wtsi_ ? wts_[i] - wts_[(i) - 1] : 0LL)
#define WG_TS_PRINT(name, cond) do { if ((cond) && (blockIdx.x == 0 || blockIdx.x == 100)) \
        printf("%s cta %d: %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld\n", name, blockIdx.x, WG_D(1), WG_D(2), WG_D(3), WG_D(4), \
               WG_D(5), WG_D(6), WG_D(7), WG_D(8), WG_D(9), WG_D(10), WG_D(11), WG_D(12), WG_D(13), WG_D(14)); } while (0)
#else
#define WG_TS_DECL
#define WG_TS()
#define WG_TS_PRINT(name, cond)
#endif

// Shared-memory carve-up of the K3 / K4a kernels (offsets from a 1024-byte aligned base).
struct Fwd {
    static constexpr uint32_t R = 0;                          // 128 KB: X hi|lo during layer 1, then activation hi
    static constexpr uint32_t XLO = 3 * ATOM;                 // X lo (K <= 96) inside R
    static constexpr uint32_t WB = 8 * ATOM;                  // 2 weight stages
    static constexpr uint32_t BIAS1 = WB + 2 * STAGE;         // b1[256]
    static constexpr uint32_t BIAS2 = BIAS1 + 1024;           // b2[256]
    static constexpr uint32_t W3S = BIAS2 + 1024;             // W3[256][4]
    static constexpr uint32_t QP = W3S + 4096;                // q partials [2][128][4]
    static constexpr uint32_t ROWF = QP + 4096;               // per-row words [6][128]: loss term, q[4], action
    static constexpr uint32_t BARS = ROWF + 4096;             // full[4] | empty[4] mbarriers | tmem base
    static constexpr uint32_t TOTAL = BARS + 128;
};

// ------------------------------------------------------------------------------------------
// Streamed GEMM: D[128 x 256] (TMEM columns 0..255) = A[128 x K] * op(W), fp32 via PASSES MMAs.
//   A hi: shared memory, K-major SW128 at a_hi; A lo: shared memory (a_lo_smem != 0) or TMEM columns.
//   BT = false: W is [K][256] row-major (x.W),   staged MN-major, 8-row chunks, 4 stages;
//   BT = true : W is [256][K] row-major (d.W^T), staged K-major SW64, 16-column chunks, 2 stages.
// Warp-specialised: the 8 producer warps copy + split weight chunks and arrive on full[stage]; lane 0
// of warp 8 waits on full[stage], issues the MMAs and commits to empty[stage], which producers wait on
// before refilling.  No CTA-wide barrier inside the loop.  cnt[] = chunks that have gone through each
// stage so far (mbarrier phase bookkeeping; both roles run the same sequence).
// ------------------------------------------------------------------------------------------
// Barrier block of the K3 / K4a kernels (byte offsets from Fwd::BARS).
struct Bar {
    static constexpr uint32_t FULL_F = 0, EMPTY_F = 32;      // x.W pipe: 4 stages
    static constexpr uint32_t TMEM = 64;                     // TMEM base address (written by tcgen05.alloc)
    static constexpr uint32_t FULL_B = 72, EMPTY_B = 88;     // d.W^T pipe: 2 stages
    static constexpr uint32_t AREADY = 104, DONE = 112;      // A operand published (8 warps) / GEMM retired (1 commit)
};

template <bool BT>
struct Pipe {
    static constexpr int KCX = BT ? 16 : 8;                   // k extent of a chunk
    static constexpr int NST = BT ? 2 : 4;                    // shared-memory stages
    static constexpr int GROUPS = NST;                        // one producer group per stage: chunk c belongs to group c % NST
    static constexpr int WPC = (NT / 32) / GROUPS;            // warps of a group (2 forward, 4 backward)
    static constexpr uint32_t HALF = KCX * H * 4;             // bytes of the hi (or lo) part of a stage
    static constexpr int PIECES = KCX * H / 4 / (32 * WPC);   // 16-byte pieces per lane per chunk (8)
    static constexpr uint32_t FULL = BT ? Bar::FULL_B : Bar::FULL_F, EMPTY = BT ? Bar::EMPTY_B : Bar::EMPTY_F;
};

// Phase bookkeeping shared by the producer warps and the MMA lane (both walk the same GEMM sequence):
// how often every stage of each pipe has been used, and how many GEMMs have retired.
struct PipeState {
    uint32_t uses_f = 0, uses_b = 0, gemms = 0;
};

__device__ __forceinline__ float4 ldg_stream(const float* p) {   // volatile: issue order = program order
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_plain(const float* p) {    // coherent (the same kernel writes these arrays), issue order kept
    float4 v;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void sts4(uint32_t a, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Every producer thread calls this once its part of the next GEMM's A operand (shared memory and / or
// TMEM) is written: generic-proxy writes -> async proxy, tcgen05.st -> tcgen05.mma, one arrival per warp.
__device__ __forceinline__ void a_ready(uint32_t sbase) {
    fence_async_smem();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(sbase + Fwd::BARS + Bar::AREADY);
}

// Producer side of a streamed GEMM.  Every shared-memory stage is OWNED by one producer group (a warp pair
// forward, four warps backward): group g stages chunks g, g + NST, ... on its own -- global -> registers ->
// (hi | lo) -> shared memory -> one arrival per warp on full[g] -- and keeps its next TWO chunks in
// registers, i.e. the loads of a chunk are issued 2 * NST chunk-times (~3500 cycles of MMA work) before it
// is staged.  Groups never wait for each other, so the per-chunk latency chain (empty wait, stores, proxy
// fence, arrive) of one group overlaps with the chains of the others instead of pacing the whole CTA.
// Because a group sees every phase of its own stage's empty[] barrier, its parity waits are never more than
// one phase behind (a waiter that skips phases cannot tell "two ahead" from "not yet": that deadlocked an
// earlier version in which a stage was shared by two groups).
// begin() can be called ahead of the epilogue that produces the A operand, run() after a_ready().
template <int PASSES, bool BT>
struct WStream {
    using P = Pipe<BT>;
    float4 buf[2][P::PIECES];
    const float* src;             // this lane's piece 0 of chunk 0
    uint32_t dst;                 // its offset inside a stage
    int nchunks, grp;
    // forward : lane L of the pair (0..63), piece i -> k-row i, columns 4L: a pair load covers one whole 1 KB row
    // backward: lane L of the four warps (0..127), piece i -> row n = 32 i + L/4, k piece L%4

    __device__ __forceinline__ const float* piece_src(int i, int c) const {
        if (!BT) return src + (size_t)c * P::KCX * H + (size_t)i * H;
        return src + (size_t)c * P::KCX + (size_t)i * 32 * H;
    }
    __device__ __forceinline__ uint32_t piece_dst(int i) const {
        // forward: off_mn(H, k = i, mn): k/4 selects a 4 KB block of 8 atoms, k%4 the 128-byte row and the 32-byte xor
        if (!BT) return (dst + (uint32_t)(i >> 2) * (H >> 5) * 512 + (uint32_t)(i & 3) * 128) ^ ((uint32_t)(i & 3) << 5);
        return dst + (uint32_t)i * 32 * 64;                   // off_k64: 64-byte rows, the xor pattern repeats every 8 rows
    }
    __device__ __forceinline__ void load(int slot, int c) {
#if TC_EXP == 1
        return;
#endif
#pragma unroll
        for (int i = 0; i < P::PIECES; ++i) buf[slot][i] = ldg_stream(piece_src(i, TC_EXP == 3 ? 0 : c));
    }
    // W: [K][H] row-major (BT = false) or [H][K] row-major with K = H (BT = true)
    __device__ __forceinline__ void begin(const float* __restrict__ W, int K) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int L = (warp % P::WPC) * 32 + lane;
        nchunks = K / P::KCX;
        grp = warp / P::WPC;
        if (!BT) {
            src = W + (L << 2);
            dst = off_mn(H, 0, L << 2);                       // k = 0: no xor yet
        } else {
            src = W + (size_t)(L >> 2) * H + ((L & 3) << 2);
            dst = off_k64(H, L >> 2, (L & 3) << 2);
        }
        if (grp < nchunks) load(0, grp);
        if (grp + P::NST < nchunks) load(1, grp + P::NST);
    }
    __device__ __forceinline__ bool run(uint32_t sbase, PipeState& ps) {
        const uint32_t full = sbase + Fwd::BARS + P::FULL + 8 * grp, empty = sbase + Fwd::BARS + P::EMPTY + 8 * grp;
        const uint32_t st = sbase + Fwd::WB + grp * (2 * P::HALF);
        uint32_t u = BT ? ps.uses_b : ps.uses_f;              // use number of this group's stage
        bool ok = true;
        for (int c0 = grp; c0 < nchunks; c0 += 2 * P::NST) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c = c0 + j * P::NST;
                if (c < nchunks) {                            // uniform across the group
                    if (u && ok) ok = mbar_wait(empty, (u - 1) & 1);            // the MMAs of the previous use are done
                    ++u;
#if TC_EXP == 4
#pragma unroll
                    for (int i = 0; i < P::PIECES; ++i) asm volatile("" ::"f"(buf[j][i].x), "f"(buf[j][i].y), "f"(buf[j][i].z), "f"(buf[j][i].w));
#elif TC_EXP != 1
#pragma unroll
                    for (int i = 0; i < P::PIECES; ++i) {
                        float4 hi, lo;
                        split4<PASSES>(buf[j][i], hi, lo);
                        const uint32_t o = st + piece_dst(i);
                        sts4(o, hi);
                        if (PASSES == 3) sts4(o + P::HALF, lo);
                    }
#endif
                    if (c + 2 * P::NST < nchunks) load(j, c + 2 * P::NST);      // two own chunks ahead
                    fence_async_smem();
                    __syncwarp();
                    if ((threadIdx.x & 31) == 0) mbar_arrive(full);             // WPC arrivals complete the stage
                }
            }
        }
        if (ok) ok = mbar_wait(sbase + Fwd::BARS + Bar::DONE, ps.gemms & 1);     // every MMA of this GEMM has retired
        tc_fence_after();
        (BT ? ps.uses_b : ps.uses_f) += (uint32_t)(nchunks / P::NST);
        ps.gemms += 1;
        return ok;
    }
};

template <int PASSES, bool BT>
__device__ __forceinline__ bool gemm_mma(uint32_t sbase, uint32_t tmem, uint32_t a_hi, uint32_t a_lo_smem,
                                         uint32_t a_lo_tmem, int K, PipeState& ps) {
    using P = Pipe<BT>;
    const uint32_t full0 = sbase + Fwd::BARS + P::FULL, empty0 = sbase + Fwd::BARS + P::EMPTY;
    const uint32_t uses0 = BT ? ps.uses_b : ps.uses_f;
    const int nchunks = K / P::KCX;
    bool ok = true;
    constexpr uint32_t idesc = make_idesc(false, !BT);
    // Descriptors differ only in their 14-bit start-address field: build each once, then add (bytes >> 4).
    const uint64_t a_hi0 = make_desc(a_hi, 16, 1024, 2);
    const uint64_t a_lo0 = make_desc(a_lo_smem, 16, 1024, 2);
    const uint64_t b0 = BT ? make_desc(sbase + Fwd::WB, 16, 512, 4) : make_desc(sbase + Fwd::WB, 512, 4096, 1);
    constexpr uint32_t KSTEP_B = BT ? 32 : 2 * 4096;          // bytes between the k-steps of a chunk in the B stage
    if (ok) ok = mbar_wait(sbase + Fwd::BARS + Bar::AREADY, ps.gemms & 1);   // all 8 producer warps have published the A operand
    tc_fence_after();
    for (int c = 0; c < nchunks; ++c) {
        const int b = c % P::NST;
        const uint32_t u = uses0 + (uint32_t)(c / P::NST);
        if (ok) ok = mbar_wait(full0 + 8 * b, u & 1);         // the owning group has staged chunk c
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < P::KCX / 8; ++ks) {
            const int kg = c * P::KCX + ks * 8;
            const uint32_t a_off = ((uint32_t)(kg >> 5) * ATOM + (uint32_t)((kg & 31) >> 3) * 32) >> 4;
            const uint32_t b_off = ((uint32_t)b * (2 * P::HALF) + ks * KSTEP_B) >> 4;
            const uint64_t a_hi_d = a_hi0 + a_off;
            const uint64_t b_hi = b0 + b_off, b_lo = b0 + b_off + (P::HALF >> 4);
            uint32_t acc = (c | ks) ? 1u : 0u;
#if TC_EXP == 2
            continue;
#endif
            if (PASSES == 3) {                                // small terms first
                if (a_lo_smem) mma_ss(tmem, a_lo0 + a_off, b_hi, idesc, acc);
                else mma_ts(tmem, a_lo_tmem + (uint32_t)kg, b_hi, idesc, acc);
                mma_ss(tmem, a_hi_d, b_lo, idesc, 1u);
                acc = 1u;
            }
            mma_ss(tmem, a_hi_d, b_hi, idesc, acc);
        }
        umma_commit(empty0 + 8 * b);
    }
    umma_commit(sbase + Fwd::BARS + Bar::DONE);
    (BT ? ps.uses_b : ps.uses_f) += (uint32_t)(nchunks / P::NST);
    ps.gemms += 1;
    return ok;
}

// Gather 128 observation rows -> hi (R) and lo (R + XLO), K-major SW128; rows past the batch are zero.
// begin() issues every load of the thread (row ids first, then up to 12 independent 16-byte pieces, the
// pieces of a row on consecutive lanes); store() splits and writes them once R is free.
template <int PASSES>
struct XGather {
    static constexpr int MAXP = BM * 96 / 4 / NT;             // 12 pieces per thread at obs_stride = 96
    float4 x[MAXP];
    uint32_t off[MAXP];
    int n;

    __device__ __forceinline__ void begin(const float* __restrict__ ring, const int32_t* __restrict__ rows, int r0, int B, int Dp) {
        const int q = Dp >> 2;                                // 16-byte pieces per row
        n = BM * q / NT;
        const int dr = NT / q, dp = NT % q;
        int r = threadIdx.x / q, pc = threadIdx.x % q;
        int32_t rid[MAXP], col[MAXP];
#pragma unroll
        for (int j = 0; j < MAXP; ++j) {
            rid[j] = -1;
            if (j < n) {
                if (r0 + r < B) rid[j] = __ldg(rows + r0 + r);
                off[j] = off_k128(BM, r, pc << 2);
                col[j] = pc << 2;
                r += dr; pc += dp;
                if (pc >= q) { pc -= q; r += 1; }
            }
        }
#pragma unroll
        for (int j = 0; j < MAXP; ++j) {
            x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < n && rid[j] >= 0) x[j] = ldg_stream(ring + (size_t)rid[j] * Dp + col[j]);
        }
    }
    __device__ __forceinline__ void store(uint32_t sbase) {
#pragma unroll
        for (int j = 0; j < MAXP; ++j) {
            if (j < n) {
                float4 hi, lo;
                split4<PASSES>(x[j], hi, lo);
                sts4(sbase + Fwd::R + off[j], hi);
                if (PASSES == 3) sts4(sbase + Fwd::XLO + off[j], lo);
            }
        }
    }
};

// Per-thread epilogue coordinates: thread = batch row of the tile, two warps share a TMEM
// sub-partition and split the 256 columns in halves.
struct Epi {
    int row, half;
    uint32_t lane_addr;
    __device__ __forceinline__ Epi() {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        row = (warp & 3) * 32 + lane;
        half = warp >> 2;
        lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    }
};

// Epilogue of a hidden layer feeding another GEMM: a = relu(D + bias) (or D * mask for dh2 built by
// the caller); hi -> shared memory R (next A operand), lo -> TMEM columns 256.., raw -> global (optional),
// returns the relu mask bits of this thread's 128 columns.
// Scratch activations are stored TRANSPOSED, [feature][batch]: a warp's 32 lanes are 32 consecutive
// batch rows, so every store below is one full 128-byte line, and K4b can stream them K-major.
template <int PASSES>
__device__ __forceinline__ void epi_hidden(uint32_t sbase, uint32_t tmem, const Epi& e, uint32_t bias_off,
                                           float* __restrict__ gout_t, int ldt, uint32_t (&mask)[4]) {
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
        const int c0 = e.half * 128 + cc * 32;
        float v[32], lo[32];
        float4 bq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)       // every bias load of the block before the first (volatile) store below
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bq[j].x), "=f"(bq[j].y), "=f"(bq[j].z), "=f"(bq[j].w)
                         : "r"(sbase + bias_off + (uint32_t)(c0 + 4 * j) * 4));
        tmem_ld32(tmem + e.lane_addr + (uint32_t)c0, v);
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b = bq[j >> 2];
            float4 x = make_float4(fmaxf(v[j] + b.x, 0.f), fmaxf(v[j + 1] + b.y, 0.f), fmaxf(v[j + 2] + b.z, 0.f),
                                   fmaxf(v[j + 3] + b.w, 0.f));
            m |= (x.x > 0.f ? 1u : 0u) << j | (x.y > 0.f ? 1u : 0u) << (j + 1) | (x.z > 0.f ? 1u : 0u) << (j + 2) |
                 (x.w > 0.f ? 1u : 0u) << (j + 3);
            if (gout_t) {
                gout_t[(size_t)(c0 + j) * ldt] = x.x; gout_t[(size_t)(c0 + j + 1) * ldt] = x.y;
                gout_t[(size_t)(c0 + j + 2) * ldt] = x.z; gout_t[(size_t)(c0 + j + 3) * ldt] = x.w;
            }
            float4 hi, l4;
            split4<PASSES>(x, hi, l4);
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sbase + Fwd::R + off_k128(BM, e.row, c0 + j)), "f"(hi.x),
                         "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
            lo[j] = l4.x; lo[j + 1] = l4.y; lo[j + 2] = l4.z; lo[j + 3] = l4.w;
        }
        mask[cc] = m;
        if (PASSES == 3) tmem_st32(tmem + e.lane_addr + 256u + (uint32_t)c0, lo);
    }
    if (PASSES == 3) tmem_st_wait();
    // visibility to the MMA warp comes with this thread's next arrive on a full[] barrier
}

// Epilogue of layer 2 feeding the 4-wide head: q[a] = b3[a] + sum_j relu(D[j] + b2[j]) * W3[j][a]; the two
// column halves of a row are combined in fixed order through shared memory.  Optionally stores the raw
// h2 row (for dW3) and returns the relu mask.
__device__ __forceinline__ void epi_head(uint32_t sbase, uint32_t tmem, const Epi& e, bool keep_h2,
                                         uint32_t (&mask)[4], float (&q)[4], const float* __restrict__ b3) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
        const int c0 = e.half * 128 + cc * 32;
        float v[32];
        tmem_ld32(tmem + e.lane_addr + (uint32_t)c0, v);
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float b;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(b) : "r"(sbase + Fwd::BIAS2 + (uint32_t)(c0 + j) * 4));
            const float x = fmaxf(v[j] + b, 0.f);
            v[j] = x;
            m |= (x > 0.f ? 1u : 0u) << j;
            float4 w;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w)
                         : "r"(sbase + Fwd::W3S + (uint32_t)(c0 + j) * 16));
            acc.x = fmaf(x, w.x, acc.x); acc.y = fmaf(x, w.y, acc.y);
            acc.z = fmaf(x, w.z, acc.z); acc.w = fmaf(x, w.w, acc.w);
        }
        mask[cc] = m;
        if (keep_h2) {      // raw h2 tile -> R (free once layer 2 has run), read back column-wise for dW3 / db2
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sbase + Fwd::R + off_k128(BM, e.row, c0 + j)), "f"(v[j]),
                             "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
        }
    }
    float4* qp = reinterpret_cast<float4*>(__cvta_shared_to_generic((size_t)(sbase + Fwd::QP)));
    qp[e.half * BM + e.row] = acc;
    prod_sync();
    const float4 p0 = qp[e.row], p1 = qp[BM + e.row];
    q[0] = (p0.x + p1.x) + __ldg(b3 + 0); q[1] = (p0.y + p1.y) + __ldg(b3 + 1);
    q[2] = (p0.z + p1.z) + __ldg(b3 + 2); q[3] = (p0.w + p1.w) + __ldg(b3 + 3);
}

__device__ __forceinline__ void load_small_params(uint32_t sbase, const float* __restrict__ P, const Layout& L) {
    float* s = reinterpret_cast<float*>(__cvta_shared_to_generic((size_t)sbase));
    for (int j = threadIdx.x; j < H; j += NT) {
        s[Fwd::BIAS1 / 4 + j] = __ldg(P + L.b1 + j);
        s[Fwd::BIAS2 / 4 + j] = __ldg(P + L.b2 + j);
        reinterpret_cast<float4*>(s + Fwd::W3S / 4)[j] = __ldg(reinterpret_cast<const float4*>(P + L.w3) + j);
    }
}

__device__ __forceinline__ uint32_t tc_prologue(uint32_t sbase) {
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 4; ++b) {
            mbar_init(sbase + Fwd::BARS + Bar::FULL_F + 8 * b, Pipe<false>::WPC);    // full[b]: the owning producer group
            mbar_init(sbase + Fwd::BARS + Bar::EMPTY_F + 8 * b, 1);                  // empty[b]: one tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(sbase + Fwd::BARS + Bar::FULL_B + 8 * b, Pipe<true>::WPC);
            mbar_init(sbase + Fwd::BARS + Bar::EMPTY_B + 8 * b, 1);
        }
        mbar_init(sbase + Fwd::BARS + Bar::AREADY, NT / 32);
        mbar_init(sbase + Fwd::BARS + Bar::DONE, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + Fwd::BARS + Bar::TMEM), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(sbase + Fwd::BARS + Bar::TMEM));
    return tmem;
}
__device__ __forceinline__ void tc_epilogue(uint32_t tmem) {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ------------------------------------------------------------------------------------------
// K3: one CTA = (network, 128-row tile, which parameter set): Q(s') of the online net (-> q_next) or of the
// target net (-> tq_all).  The two halves are independent CTAs (twice as many, half as long: less tail on
// 148 SMs); K4a combines them into the TD target of its rows (reference :342-347).
// ------------------------------------------------------------------------------------------
template <int PASSES>
__global__ void __launch_bounds__(NT_F, 1) tc_target_kernel(const TcArgs A) {
    extern __shared__ uint8_t smem_raw[];
    const int item = blockIdx.x >> 1, pass = blockIdx.x & 1;          // 0: online(s')  1: target(s')
    const int g = item / A.tiles, rt = item % A.tiles;
    if (!A.active[g]) return;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int B = A.d.batch, Dp = A.d.obs_stride, r0 = rt * BM;
    const uint32_t tmem = tc_prologue(sbase);
    PipeState ps;
    bool ok = true;
    if (threadIdx.x >= NT) {
        // ---- MMA warp: lane 0 issues every tcgen05.mma of this CTA ----
        if (threadIdx.x == NT) {
            ok = ok && gemm_mma<PASSES, false>(sbase, tmem, sbase + Fwd::R, sbase + Fwd::XLO, 0, Dp, ps);
            ok = ok && gemm_mma<PASSES, false>(sbase, tmem, sbase + Fwd::R, 0, tmem + 256u, H, ps);
            if (!ok) atomicExch(A.error, 13);
        }
        __syncwarp();
    } else {
        const Epi e;
        const float* P = (pass == 0 ? A.nets.theta : A.nets.theta_tgt) + (size_t)g * A.L.stride;
        float q[4];
        TS_DECL;
        TS();
        WStream<PASSES, false> ws;
        ws.begin(P + A.L.w1, Dp);                      // first W1 chunks in flight before anything else
        {
            XGather<PASSES> xg;
            xg.begin(A.rp.next_obs, A.rows + (size_t)g * B, r0, B, Dp);
            load_small_params(sbase, P, A.L);
            xg.store(sbase);
        }
        a_ready(sbase);
        prod_sync();   // biases / head weights visible to every epilogue thread
        TS();
        ok = ok && ws.run(sbase, ps);
        TS();
        ws.begin(P + A.L.w2, H);                       // W2 chunks travel while the epilogue runs
        uint32_t mask[4];
        epi_hidden<PASSES>(sbase, tmem, e, Fwd::BIAS1, nullptr, 0, mask);
        a_ready(sbase);
        TS();
        ok = ok && ws.run(sbase, ps);
        TS();
        epi_head(sbase, tmem, e, false, mask, q, P + A.L.b3);
        TS();
        TS_PRINT("K3 gather L1 epi1 L2 epi2");
        const int gr = r0 + e.row;
        if (e.half == 0 && gr < B) {
            float* out = (pass == 0 ? A.q_next : A.tq_all) + ((size_t)g * B + gr) * 4;
            *reinterpret_cast<float4*>(out) = make_float4(q[0], q[1], q[2], q[3]);
        }
    }
    if (!ok && threadIdx.x == 0) atomicExch(A.error, 3);
    tc_epilogue(tmem);
}

// ------------------------------------------------------------------------------------------
// K4a
// ------------------------------------------------------------------------------------------
template <int PASSES>
__global__ void __launch_bounds__(NT_F, 1) tc_online_kernel(const TcArgs A) {
    extern __shared__ uint8_t smem_raw[];
    const int g = blockIdx.x / A.tiles, rt = blockIdx.x % A.tiles;
    if (!A.active[g]) return;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    float* sf = reinterpret_cast<float*>(__cvta_shared_to_generic((size_t)sbase));
    const int B = A.d.batch, Dp = A.d.obs_stride, r0 = rt * BM;
    const size_t sb = (size_t)g * B;
    const int32_t* rows = A.rows + sb;
    const float* P = A.nets.theta + (size_t)g * A.L.stride;
    const uint32_t tmem = tc_prologue(sbase);
    PipeState ps;
    bool ok = true;
    if (threadIdx.x >= NT) {
        // ---- MMA warp: lane 0 issues every tcgen05.mma of this CTA ----
        if (threadIdx.x == NT) {
            ok = ok && gemm_mma<PASSES, false>(sbase, tmem, sbase + Fwd::R, sbase + Fwd::XLO, 0, Dp, ps);
            ok = ok && gemm_mma<PASSES, false>(sbase, tmem, sbase + Fwd::R, 0, tmem + 256u, H, ps);
            ok = ok && gemm_mma<PASSES, true>(sbase, tmem, sbase + Fwd::R, 0, tmem + 256u, H, ps);
            if (!ok) atomicExch(A.error, 14);
        }
        __syncwarp();
        tc_epilogue(tmem);
        return;
    }
    const Epi e;
    const int gr = r0 + e.row;
    const bool valid = gr < B;

    TS_DECL;
    TS();
    WStream<PASSES, false> ws;
    ws.begin(P + A.L.w1, Dp);
    {
        XGather<PASSES> xg;
        xg.begin(A.rp.obs, rows, r0, B, Dp);
        load_small_params(sbase, P, A.L);
        xg.store(sbase);
    }
    // TD target of this row from the two halves of K3 (reference :342-347), ties -> lowest index
    float yi = 0.f;
    if (valid) {
        const float4 qo = *reinterpret_cast<const float4*>(A.q_next + (sb + gr) * 4);
        const float4 qt = *reinterpret_cast<const float4*>(A.tq_all + (sb + gr) * 4);
        const float q_on[4] = {qo.x, qo.y, qo.z, qo.w}, q_tg[4] = {qt.x, qt.y, qt.z, qt.w};
        float bmax = q_on[0], tq = q_tg[0], tmax = q_tg[0];
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            if (k < A.d.n_actions) {
                if (q_on[k] > bmax) { bmax = q_on[k]; tq = q_tg[k]; }
                tmax = fmaxf(tmax, q_tg[k]);
            }
        }
        tq = A.double_dqn ? tq : tmax;
        yi = A.r_hat[sb + gr] + (A.gamma * (1.0f - A.done_b[sb + gr])) * tq;
        if (e.half == 0) A.y[sb + gr] = yi;
    }
    a_ready(sbase);
    prod_sync();   // biases / head weights visible to every epilogue thread
    TS();
    ok = ok && ws.run(sbase, ps);
    TS();
    ws.begin(P + A.L.w2, H);
    uint32_t mask1[4], mask2[4];
    epi_hidden<PASSES>(sbase, tmem, e, Fwd::BIAS1, valid ? A.h1 + sb * H + gr : nullptr, B, mask1);
    a_ready(sbase);
    TS();
    ok = ok && ws.run(sbase, ps);
    TS();
    float q[4];
    epi_head(sbase, tmem, e, true, mask2, q, P + A.L.b3);
    TS();

    // loss term and dL/dpred of this row (reference :349-352)
    float gi = 0.f, term = 0.f;
    int ai = 0;
    if (valid) {
        ai = A.act_b[sb + gr];
        const float qa = ai == 0 ? q[0] : ai == 1 ? q[1] : ai == 2 ? q[2] : q[3];
        const float err = qa - yi;
        if (A.loss == DMDQN_LOSS_MSE) {
            term = err * err;
            gi = (2.0f * err) / (float)A.loss_batch;
        } else {
            const float ae = fabsf(err);
            term = ae <= 1.0f ? 0.5f * err * err : ae - 0.5f;
            gi = fminf(fmaxf(err, -1.0f), 1.0f) / (float)A.loss_batch;
        }
        if (e.half == 0) {
            A.gcoef[sb + gr] = gi;
            A.ga[sb + gr] = make_float2(gi, __int_as_float(ai));
            for (int k = 0; k < 4; ++k) A.q_all[(sb + gr) * 4 + k] = q[k];
        }
    }
    float* rowf = sf + Fwd::ROWF / 4;                     // [1..4]: per row float4 g * onehot(action), [7]: warp partials
    if (e.half == 0) {
        // dL/dq of this row as a 4-vector (one non-zero): the column loop below then needs no selects
        reinterpret_cast<float4*>(rowf + BM)[e.row] = make_float4(ai == 0 ? gi : 0.f, ai == 1 ? gi : 0.f, ai == 2 ? gi : 0.f, ai == 3 ? gi : 0.f);
        // per-tile loss / metric / db3 partials: butterfly over the 32 rows of this warp (fixed order), then
        // the four warp partials are added in warp order by thread 0
        float red[11];
        red[0] = term;
        red[1] = 0.f; red[2] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float qq = (valid && k < A.d.n_actions) ? q[k] : 0.f;
            red[1] += qq;
            red[2] = fmaf(qq, qq, red[2]);
            red[3 + k] = (valid && ai == k) ? 1.f : 0.f;
            red[7 + k] = (valid && ai == k) ? gi : 0.f;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int v = 0; v < 11; ++v) red[v] += __shfl_xor_sync(0xffffffffu, red[v], o);
        if ((threadIdx.x & 31) == 0)
#pragma unroll
            for (int v = 0; v < 11; ++v) rowf[7 * BM + (threadIdx.x >> 5) * 11 + v] = red[v];
    }
    prod_sync();
    if (threadIdx.x < 11) {
        const float* wp = rowf + 7 * BM + threadIdx.x;
        const float tot = ((wp[0] + wp[11]) + wp[22]) + wp[33];
        const size_t pt = (size_t)g * A.tiles + rt;
        if (threadIdx.x < 7) A.part_loss[pt * 8 + threadIdx.x] = tot;      // loss, q sum, q^2 sum, action histogram
        else A.part_b3[pt * 4 + (threadIdx.x - 7)] = tot;
    }
    {   // dW3[j][a] = sum_i h2[i][j] g_i [a_i = a] and db2[j] = sum_i dh2[i][j] over this tile's rows (in order):
        // thread = column j, rows read back from the h2 tile left in R by epi_head
        const int j = threadIdx.x;
        float4 w3;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(w3.x), "=f"(w3.y), "=f"(w3.z), "=f"(w3.w)
                     : "r"(sbase + Fwd::W3S + (uint32_t)j * 16));
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, s2 = 0.f;
        // h2 column j of the tile: K-major SW128, the 16-byte piece index is xor-ed with row % 8
        const uint32_t hcol = sbase + Fwd::R + (uint32_t)(j >> 5) * ATOM + (uint32_t)((j & 3) << 2);
        const uint32_t jq = (uint32_t)((j & 31) >> 2);
        const uint32_t gsel = sbase + Fwd::ROWF + BM * 4;
#pragma unroll 1
        for (int i0 = 0; i0 < BM; i0 += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u;
                float h;
                float4 gs;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(h) : "r"(hcol + (uint32_t)i * 128u + ((jq ^ (uint32_t)u) << 4)));
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(gs.x), "=f"(gs.y), "=f"(gs.z), "=f"(gs.w) : "r"(gsel + (uint32_t)i * 16u));
                d0 = fmaf(h, gs.x, d0); d1 = fmaf(h, gs.y, d1); d2 = fmaf(h, gs.z, d2); d3 = fmaf(h, gs.w, d3);
                const float gw = fmaf(gs.w, w3.w, fmaf(gs.z, w3.z, fmaf(gs.y, w3.y, gs.x * w3.x)));   // = g * W3[j][a]: the other terms are exact zeros
                s2 += h > 0.f ? gw : 0.f;
            }
        }
        const size_t pt = (size_t)g * A.tiles + rt;
        reinterpret_cast<float4*>(A.part_w3 + pt * H * 4)[j] = make_float4(d0, d1, d2, d3);
        A.part_b2[pt * H + j] = s2;
    }
    prod_sync();          // R is rewritten with dh2 below
    TS();

    WStream<PASSES, true> wsb;
    wsb.begin(P + A.L.w2, H);
    // dh2[j] = relu'(h2[j]) * g * W3[j][a]  (dq has one non-zero per row): hi -> R, lo -> TMEM.  It is not
    // written to global memory: K4b rebuilds its dh2^T operand from relu'(h2) bits, g, the action and W3.
    if (valid) {
        reinterpret_cast<uint4*>(A.mask2)[(sb + gr) * 2 + e.half] = make_uint4(mask2[0], mask2[1], mask2[2], mask2[3]);
    }
    if (rt == 0)    // W3 as this step saw it (K4b updates W3 while other CTAs of the network still need the old values)
        reinterpret_cast<float4*>(A.w3_copy + (size_t)g * H * 4)[threadIdx.x] =
            *reinterpret_cast<const float4*>(sf + Fwd::W3S / 4 + threadIdx.x * 4);
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
        const int c0 = e.half * 128 + cc * 32;
        float lo[32], wv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j)      // all 32 loads first: the volatile stores below would otherwise serialise load -> store chains
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(wv[j]) : "r"(sbase + Fwd::W3S + (uint32_t)(c0 + j) * 16 + (uint32_t)ai * 4));
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            float x[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) x[t] = ((mask2[cc] >> (j + t)) & 1u) ? gi * wv[j + t] : 0.f;
            const float4 x4 = make_float4(x[0], x[1], x[2], x[3]);
            float4 hi, l4;
            split4<PASSES>(x4, hi, l4);
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sbase + Fwd::R + off_k128(BM, e.row, c0 + j)), "f"(hi.x),
                         "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
            lo[j] = l4.x; lo[j + 1] = l4.y; lo[j + 2] = l4.z; lo[j + 3] = l4.w;
        }
        if (PASSES == 3) tmem_st32(tmem + e.lane_addr + 256u + (uint32_t)c0, lo);
    }
    if (PASSES == 3) tmem_st_wait();
    a_ready(sbase);

    TS();
    // dh1 = (dh2 W2^T) * relu'(h1)
    ok = ok && wsb.run(sbase, ps);
    TS();
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
        const int c0 = e.half * 128 + cc * 32;
        float v[32];
        tmem_ld32(tmem + e.lane_addr + (uint32_t)c0, v);
        if (valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                A.dh1[sb * H + (size_t)(c0 + j) * B + gr] = ((mask1[cc] >> j) & 1u) ? v[j] : 0.f;
        }
    }
    TS();
    TS_PRINT("K4a gather L1 epi1 L2 epi2 loss+dW3 dh2 bwdGEMM dh1store");
    if (!ok && threadIdx.x == 0) atomicExch(A.error, 4);
    tc_epilogue(tmem);
}

// ------------------------------------------------------------------------------------------
// K4b
// ------------------------------------------------------------------------------------------
struct AdamK {
    float alpha, eps, omb1, omb2, tau;
    int sync;
};
__device__ __forceinline__ AdamK adam_k(const TcArgs& A, int g) {   // scalars prepared by the sample kernel
    AdamK k;
    const float4 sc = __ldg(A.adam_sc + g);
    k.alpha = sc.x;
    k.eps = sc.y;
    k.sync = __float_as_int(sc.z);
    k.omb1 = (float)(1.0 - A.beta1);
    k.omb2 = (float)(1.0 - A.beta2);
    k.tau = (float)A.tau;
    return k;
}
__device__ __forceinline__ void adam1(const AdamK& k, float g, float& th, float& m, float& v, float& tg) {
#if WG_EXP & 1
    th -= g; m += g; v += g; return;
#endif
    m = m + (g - m) * k.omb1;
    v = v + (g * g - v) * k.omb2;
    // MUFU square root and reciprocal (each within ~2 ulp): the quotient is a step of at most ~lr, so a few
    // ulp of it are far below one ulp of the weight it is subtracted from; the IEEE sequences cost ~60
    // dependent instructions per element and made the epilogue issue-bound
    float sq;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(v));
    th = th - __fdividef(m * k.alpha, sq + k.eps);
    if (k.sync == 1) tg = th;
    else if (k.sync == 2) tg = k.tau * th + (1.0f - k.tau) * tg;
}

// Adam on one element without the target sync (the caller applies it per 16-byte group)
__device__ __forceinline__ void adam_fast(const AdamK& k, float g, float& th, float& m, float& v) {
#if WG_EXP & 1
    th -= g; m += g; v += g; return;
#endif
    m = m + (g - m) * k.omb1;
    v = v + (g * g - v) * k.omb2;
    float sq;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(v));
    th = th - __fdividef(m * k.alpha, sq + k.eps);
}

struct Wg {     // persistent wgrad kernel, one CTA per SM
    static constexpr int STAGES = 3;
    static constexpr uint32_t A_BYTES = KC * BM * 4;          // 8 KB
    static constexpr uint32_t B_BYTES = KC * H * 4;           // 16 KB
    static constexpr uint32_t STG = 2 * A_BYTES + 2 * B_BYTES;   // A hi|lo, B hi|lo: 48 KB
    static constexpr int TLD = 36;                            // floats per row of an epilogue transpose tile
    static constexpr uint32_t TILES = STAGES * STG;           // 8 epilogue warps x 32 x TLD floats
    static constexpr uint32_t W3T = TILES + 8 * 32 * TLD * 4; // W3^T [2 item parities][4 actions][256] floats
    static constexpr uint32_t BARS = W3T + 2 * 4 * H * 4;
    static constexpr uint32_t FULL = 0, EMPTY = 24, ACC_FULL = 48, ACC_FREE = 64, TMEM = 80;   // byte offsets from BARS
    static constexpr uint32_t TOTAL = BARS + 128;
    static constexpr int NTW = 16 * 32;                       // 8 producer warps (lane 0 of warp 0 also issues the MMAs), 8 epilogue warps
};

// Work items of K4b: per network two dW2 tiles (rows t*128..) and one dW1 tile.  The dW2 tiles come first in
// the item order (they are the longer ones), so a static round-robin over the CTAs ends on the short items.
__device__ __forceinline__ void wg_item(int q, int G, int& g, int& t) {
    if (q < 2 * G) { g = q >> 1; t = q & 1; }
    else { g = q - 2 * G; t = 2; }
}

// dW tile: D[128 x 256] = A^T-operand * D-operand over K = batch, Adam in the epilogue.
//   t = 0,1: dW2 rows t*128.. : A = h1^T scratch [m][k] (K-major SW64), B = dh2 rebuilt [k][n] (MN-major)
//   t = 2  : dW1 rows 0..95   : A = s gathered from the ring [k][m] (MN-major),  B = dh1^T scratch (K-major SW64);
//            A column m = obs_stride is set to 1, so row obs_stride of the tile is db1 = sum_k dh1[k][:] for free;
//            the epilogue of this item also folds the per-row-tile partials of db2 / dW3 / db3 (from K4a) and
//            emits the metrics.
// Persistent and warp-specialised: a CTA walks its items q = blockIdx.x, + gridDim.x, ...;
//   warps 0-7  stage operand chunks (global -> registers -> hi | lo -> shared memory, 3 stages) and run ahead
//              into the next item;
//   lane 0 of warp 0 also issues the tcgen05.mma of the chunk staged one step earlier, alternating between two
//              TMEM accumulators (columns 0..255 / 256..511), and
//   warps 8-15 read a finished accumulator and do the Adam read-modify-write of theta / m / v / theta_tgt.
// The epilogue is pure HBM traffic (24 bytes per parameter) and the GEMM needs almost none, so running them
// concurrently on every SM keeps HBM busy for the whole kernel instead of only during the epilogue phase of a
// wave of lock-stepped CTAs.
template <int PASSES>
__global__ void __launch_bounds__(Wg::NTW, 1) tc_wgrad_kernel(const TcArgs A) {
    extern __shared__ uint8_t smem_raw[];
    const int B = A.d.batch, Dp = A.d.obs_stride, G = A.d.n_nets;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = sbase + Wg::BARS;
    const int n_items = 3 * G;
    const int nchunks = (B + KC - 1) / KC;

    if (tid == 0) {
        for (int s = 0; s < Wg::STAGES; ++s) {
            mbar_init(bars + Wg::FULL + 8 * s, 8);            // one arrival per producer warp
            mbar_init(bars + Wg::EMPTY + 8 * s, 1);           // one tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bars + Wg::ACC_FULL + 8 * b, 1);        // one tcgen05.commit
            mbar_init(bars + Wg::ACC_FREE + 8 * b, 8);        // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + Wg::TMEM), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(bars + Wg::TMEM));
    bool ok = true;

    if (warp < 8) {
        // ------------------------------------------------------------------ producers ------------------
        // Operand chunks (16 batch rows) travel global -> registers -> (hi | lo) -> shared memory; the loads of
        // chunk c + 2 are issued when chunk c is staged, so two chunks of latency are covered by registers and
        // the shared-memory stages only have to cover the MMAs.
        uint32_t stage = 0, ph = 0, used = 0;
        // Lane 0 of warp 0 is also the MMA issuer: after staging chunk c it issues the tcgen05.mma of chunk
        // c - 1 (which the other warps have normally finished staging by then), so no thread sits in a wait
        // that paces the others.  pend_* describe the chunk whose MMAs are still to be issued.
        bool pend = false, pend_w2 = false, pend_first = false, pend_last = false;
        uint32_t pend_stage = 0, pend_ph = 0;
        int pend_n = 0, n = 0;
        auto issue_pending = [&]() {
            if (lane == 0) {
                const uint32_t acc_buf = (uint32_t)(pend_n & 1);
                if (pend_first && pend_n >= 2 && ok)              // the epilogue has read item n - 2 out of this accumulator
                    ok = mbar_wait(bars + Wg::ACC_FREE + 8 * acc_buf, (uint32_t)((pend_n >> 1) - 1) & 1u);
                if (ok) ok = mbar_wait(bars + Wg::FULL + 8 * pend_stage, pend_ph);
                tc_fence_after();
                const uint32_t d_tmem = tmem + acc_buf * 256u;
                const uint32_t idesc = make_idesc(!pend_w2, pend_w2);
                const uint32_t st = sbase + pend_stage * Wg::STG;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    uint64_t a_hi, a_lo;
                    if (pend_w2) {
                        a_hi = make_desc(st + ks * 32, 16, 512, 4);
                        a_lo = make_desc(st + Wg::A_BYTES + ks * 32, 16, 512, 4);
                    } else {      // MN-major: 4 m-groups per k-group -> k-group stride 2048 B
                        a_hi = make_desc(st + ks * 2 * 2048, 512, 2048, 1);
                        a_lo = make_desc(st + Wg::A_BYTES + ks * 2 * 2048, 512, 2048, 1);
                    }
                    uint64_t b_hi, b_lo;
                    if (pend_w2) {    // MN-major: 8 n-groups per k-group -> k-group stride 4096 B, a k-step is two k-groups
                        b_hi = make_desc(st + 2 * Wg::A_BYTES + ks * 2 * 4096, 512, 4096, 1);
                        b_lo = make_desc(st + 2 * Wg::A_BYTES + Wg::B_BYTES + ks * 2 * 4096, 512, 4096, 1);
                    } else {
                        b_hi = make_desc(st + 2 * Wg::A_BYTES + ks * 32, 16, 512, 4);
                        b_lo = make_desc(st + 2 * Wg::A_BYTES + Wg::B_BYTES + ks * 32, 16, 512, 4);
                    }
                    uint32_t acc = (!pend_first || ks) ? 1u : 0u;
                    if (WG_EXP & 32) continue;
                    if (PASSES == 3) {
                        mma_ss(d_tmem, a_lo, b_hi, idesc, acc);
                        mma_ss(d_tmem, a_hi, b_lo, idesc, 1u);
                        acc = 1u;
                    }
                    mma_ss(d_tmem, a_hi, b_hi, idesc, acc);
                }
                umma_commit(bars + Wg::EMPTY + 8 * pend_stage);
                if (pend_last) umma_commit(bars + Wg::ACC_FULL + 8 * acc_buf);
            }
            __syncwarp();
        };
        // Called by every producer thread before / after it writes its pieces of a chunk into stage `stage`.
        auto stage_begin = [&]() -> uint32_t {
            if (used >= (uint32_t)Wg::STAGES && ok) ok = mbar_wait(bars + Wg::EMPTY + 8 * stage, ph ^ 1u);   // MMAs of the previous use are done
            return sbase + stage * Wg::STG;
        };
        auto stage_end = [&](bool is_w2, int c) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + Wg::FULL + 8 * stage);
            if (warp == 0) {
                if (pend) issue_pending();
                pend = true; pend_w2 = is_w2; pend_first = c == 0; pend_last = c == nchunks - 1;
                pend_stage = stage; pend_ph = ph; pend_n = n;
            }
            ++used;
            if (++stage == (uint32_t)Wg::STAGES) { stage = 0; ph ^= 1u; }
        };
        auto put = [&](uint32_t o, uint32_t lo_off, const float4& x) {
            if (WG_EXP & 64) { asm volatile("" ::"f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w)); return; }
            float4 hi, lo;
            split4<PASSES>(x, hi, lo);
            sts4(o, hi);
            if (PASSES == 3) sts4(o + lo_off, lo);
        };
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        WG_TS_DECL;
        for (int q = blockIdx.x; q < n_items; q += gridDim.x) {
            int g, t;
            wg_item(q, G, g, t);
            if (!A.active[g]) continue;
            WG_TS();
            const size_t sb = (size_t)g * B;
            if (t < 2) {
                // ---- dW2 rows t*128..: A = h1^T scratch [m][k] (K-major SW64), register prefetch DW2 chunks ahead;
                // B = dh2 rebuilt, staged MN-major ([k][n], n contiguous): dh2[k][n] = relu'(h2)[k][n] ? g_k * W3[n][a_k] : 0,
                // thread = (k-row bu of the chunk, 16 columns n = 64 i + 4 ng + 0..3, i = 0..3), so one (g, action) pair
                // and four mask words serve 16 values, W3^T comes from shared memory as four conflict-free 16-byte
                // loads, and every store is a 16-byte piece of the UMMA layout.
                constexpr int DW2 = 2;
                const int bu = tid >> 4, ng = tid & 15, hh = ng >> 3, l7 = ng & 7;
                const uint32_t w3t = sbase + Wg::W3T + (uint32_t)(n & 1) * (4 * H * 4);
                {
                    const float4 w3n = __ldg(reinterpret_cast<const float4*>(A.w3_copy + (size_t)g * H * 4) + tid);
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(w3t + (uint32_t)(0 * H + tid) * 4), "f"(w3n.x) : "memory");
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(w3t + (uint32_t)(1 * H + tid) * 4), "f"(w3n.y) : "memory");
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(w3t + (uint32_t)(2 * H + tid) * 4), "f"(w3n.z) : "memory");
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(w3t + (uint32_t)(3 * H + tid) * 4), "f"(w3n.w) : "memory");
                }
                prod_sync();      // every item: a warp two items ahead would overwrite the buffer a slower warp still reads
                // A: pieces p = tid, tid + 256: row m = p/4, k piece p%4.  B: piece i of the thread = columns 32 (4 hh + i) + 4 l7 + 0..3,
                // i.e. mask words 4 hh .. 4 hh + 3 of its batch row (one 16-byte load), W3^T at + 128 i bytes, stage atom + i.
                const float* pa = A.h1 + sb * H + (size_t)(t * BM + (tid >> 2)) * B + ((tid & 3) << 2);
                const size_t a_half = (size_t)64 * B;
                const uint32_t a_dst = off_k64(BM, tid >> 2, (tid & 3) << 2);       // second piece: + 64 rows = + 4096 B
                const float2* pga = A.ga + sb + bu;
                const uint4* pmk = reinterpret_cast<const uint4*>(A.mask2) + (sb + bu) * 2 + hh;
                const int mshift = l7 << 2, ka0 = (tid & 3) << 2;
                const uint32_t w3t_thr = w3t + (uint32_t)(128 * hh + 4 * l7) * 4;
                const uint32_t b_dst = 2 * Wg::A_BYTES + off_mn(H, bu, 128 * hh + 4 * l7);   // piece i: + i atoms of 512 B
                float4 ra[DW2][2];
                float2 gq[DW2];
                uint4 mq[DW2];
                auto load = [&](int slot, int c) {
                    ra[slot][0] = z4; ra[slot][1] = z4;
                    gq[slot] = make_float2(0.f, 0.f);
                    mq[slot] = make_uint4(0u, 0u, 0u, 0u);
                    if (WG_EXP & 4) return;
                    if (c * KC + ka0 < B) {
                        ra[slot][0] = ldg_stream(pa + c * KC);
                        ra[slot][1] = ldg_stream(pa + c * KC + a_half);
                    }
                    if (c * KC + bu < B) {
                        gq[slot] = __ldg(pga + c * KC);
                        mq[slot] = __ldg(pmk + c * (KC * 2));
                    }
                };
#pragma unroll
                for (int j = 0; j < DW2; ++j)
                    if (j < nchunks) load(j, j);
                for (int c0 = 0; c0 < nchunks; c0 += DW2) {
#pragma unroll
                    for (int j = 0; j < DW2; ++j) {
                        const int c = c0 + j;
                        if (c < nchunks) {                        // uniform across the CTA
                            const uint32_t st = stage_begin();
                            put(st + a_dst, Wg::A_BYTES, ra[j][0]);
                            put(st + a_dst + 4096u, Wg::A_BYTES, ra[j][1]);
                            const uint32_t wa = w3t_thr + (uint32_t)__float_as_int(gq[j].y) * (H * 4);
                            const float gg = gq[j].x;
                            const uint32_t mw[4] = {mq[j].x, mq[j].y, mq[j].z, mq[j].w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float4 w4;
                                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(w4.x), "=f"(w4.y), "=f"(w4.z), "=f"(w4.w)
                                             : "r"(wa + (uint32_t)i * 128u));
                                const uint32_t bits = mw[i] >> mshift;
                                const float4 x = make_float4((bits & 1u) ? gg * w4.x : 0.f, (bits & 2u) ? gg * w4.y : 0.f,
                                                             (bits & 4u) ? gg * w4.z : 0.f, (bits & 8u) ? gg * w4.w : 0.f);
                                put(st + b_dst + (uint32_t)i * 512u, Wg::B_BYTES, x);
                            }
                            if (c + DW2 < nchunks) load(j, c + DW2);
                            stage_end(true, c);
                        }
                    }
                }
            } else {
                // ---- dW1 (+ db1): A = s rows gathered from the ring [k][m] (MN-major; column m = Dp is the ones column),
                // B = dh1^T scratch [n][k] (K-major SW64); register prefetch DW1 chunks ahead.  The sampled row ids
                // of the item are read once (one per thread and pass) instead of once per gathered piece.
                constexpr int DW1 = 2;
                const float* b_src = A.dh1 + sb * H + (size_t)(tid >> 2) * B + ((tid & 3) << 2);   // pieces p = tid + 256 r: row n = p/4 = tid/4 + 64 r
                const uint32_t b_dst = 2 * Wg::A_BYTES + off_k64(H, tid >> 2, (tid & 3) << 2);  // + r * 64 rows * 64 B
                const int am = (tid & 31) << 2, ak = tid >> 5;                   // pieces p = tid, tid + 256: k = p/32 = ak, ak + 8
                const bool a_live = am < Dp, a_ones = am == Dp;
                const uint32_t a_dst = off_mn(BM, ak, am);                       // k + 8: two k-groups further = + 2 * 2048 B
                const int32_t* rows = A.rows + sb;
                float4 ra[DW1][2], rb[DW1][4];
                auto load = [&](int slot, int c) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int k = c * KC + ak + 8 * r;
                        ra[slot][r] = z4;
                        if (WG_EXP & 4) continue;
                        if (a_ones) ra[slot][r].x = k < B ? 1.f : 0.f;
                        else if (a_live && k < B) ra[slot][r] = ldg_stream(A.rp.obs + (size_t)__ldg(rows + k) * Dp + am);
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        rb[slot][r] = z4;
                        if (!(WG_EXP & 4) && c * KC + ((tid & 3) << 2) < B) rb[slot][r] = ldg_stream(b_src + (size_t)c * KC + (size_t)(64 * r) * B);
                    }
                };
#pragma unroll
                for (int j = 0; j < DW1; ++j)
                    if (j < nchunks) load(j, j);
                for (int c0 = 0; c0 < nchunks; c0 += DW1) {
#pragma unroll
                    for (int j = 0; j < DW1; ++j) {
                        const int c = c0 + j;
                        if (c < nchunks) {
                            const uint32_t st = stage_begin();
                            put(st + a_dst, Wg::A_BYTES, ra[j][0]);
                            put(st + a_dst + 4096u, Wg::A_BYTES, ra[j][1]);
#pragma unroll
                            for (int r = 0; r < 4; ++r) put(st + b_dst + (uint32_t)r * 4096u, Wg::B_BYTES, rb[j][r]);
                            if (c + DW1 < nchunks) load(j, c + DW1);
                            stage_end(false, c);
                        }
                    }
                }
            }
            ++n;
        }
        if (warp == 0 && pend) issue_pending();
        WG_TS();
        WG_TS_PRINT("K4b producer", tid == 0);
        if (!ok && tid == 0) atomicExch(A.error, 5);
    } else {
        // ------------------------------------------------------------------ epilogue -------------------
        const int et = tid - NT, ew = warp - 8;                   // 0..255 / 0..7
        const int ehalf = ew >> 2;                                // TMEM lane = weight row (ew & 3) * 32 + lane of the tile; column half
        const uint32_t lane_addr = (uint32_t)((ew & 3) * 32) << 16;
        float* tile = reinterpret_cast<float*>(__cvta_shared_to_generic((size_t)(sbase + Wg::TILES))) + ew * (32 * Wg::TLD);
        const int rsub = lane >> 3, c4 = (lane & 7) << 2;
        int n = 0;
        WG_TS_DECL;
        for (int q = blockIdx.x; q < n_items; q += gridDim.x) {
            int g, t;
            wg_item(q, G, g, t);
            if (!A.active[g]) {
                if (t == 2 && et < DMDQN_METRICS_STRIDE && A.metrics) A.metrics[g * DMDQN_METRICS_STRIDE + et] = 0.f;
                continue;
            }
            const size_t pb = (size_t)g * A.L.stride;
            float* th = A.nets.theta + pb;
            float* tg = A.nets.theta_tgt + pb;
            float* am = A.nets.adam_m + pb;
            float* av = A.nets.adam_v + pb;
            const AdamK k = adam_k(A, g);
            const bool is_w2 = t < 2;
            const int m0 = is_w2 ? t * BM : 0;
            if (t == 2) {
                // head / bias gradients: per-row-tile partials from K4a summed in tile order, then Adam
                auto upd = [&](int64_t off, float grad) {
                    if (A.grads) { A.grads[pb + off] = grad; return; }
                    float tgv = k.sync == 2 ? tg[off] : 0.f;
                    adam1(k, grad, th[off], am[off], av[off], tgv);
                    if (k.sync) tg[off] = tgv;
                };
                const size_t p0 = (size_t)g * A.tiles;
                float s2 = 0.f, w[4] = {0.f, 0.f, 0.f, 0.f};
                for (int r = 0; r < A.tiles; ++r) {
                    s2 += A.part_b2[(p0 + r) * H + et];
                    const float4 pw = reinterpret_cast<const float4*>(A.part_w3 + (p0 + r) * H * 4)[et];
                    w[0] += pw.x; w[1] += pw.y; w[2] += pw.z; w[3] += pw.w;
                }
                upd(A.L.b2 + et, s2);
                for (int a = 0; a < 4; ++a) upd(A.L.w3 + (int64_t)et * 4 + a, w[a]);
                if (et < 4) {
                    float s3 = 0.f;
                    for (int r = 0; r < A.tiles; ++r) s3 += A.part_b3[(p0 + r) * 4 + et];
                    upd(A.L.b3 + et, s3);
                }
                if (et == 0 && A.metrics) {
                    double ls = 0, qs = 0, qq = 0, hist[4] = {0, 0, 0, 0};
                    for (int r = 0; r < A.tiles; ++r) {
                        const float* pl = A.part_loss + (p0 + r) * 8;
                        ls += pl[0]; qs += pl[1]; qq += pl[2];
                        for (int a = 0; a < 4; ++a) hist[a] += pl[3 + a];
                    }
                    const double cnt = (double)B * A.d.n_actions, mean = qs / cnt, var = fmax(qq / cnt - mean * mean, 0.0);
                    float* m = A.metrics + g * DMDQN_METRICS_STRIDE;
                    m[0] = (float)(ls / A.loss_batch); m[1] = (float)mean; m[2] = (float)sqrt(var);
                    for (int a = 0; a < 4; ++a) m[3 + a] = (float)hist[a];
                    m[7] = 1.f;
                }
            }
            const uint32_t acc_buf = (uint32_t)(n & 1);
            WG_TS();
            if (ok) ok = mbar_wait(bars + Wg::ACC_FULL + 8 * acc_buf, (uint32_t)(n >> 1) & 1u);
            tc_fence_after();
            WG_TS();
            // TMEM (lane = weight row) -> per-warp smem tile -> 8 lanes per 128-byte row segment, so the
            // Adam read-modify-write of theta / m / v / theta_tgt is fully coalesced.  Offsets are 32-bit float
            // indices into the network's parameter block; all mode decisions are made per item, not per element.
            const int wbase = (int)(is_w2 ? A.L.w2 : A.L.w1) + (m0 + (ew & 3) * 32 + rsub) * H + c4;   // + 4 u' H + c0
            const int mrow0 = m0 + (ew & 3) * 32 + rsub;                                              // tile row of u' = 0
            const int sync = k.sync;
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                const int c0 = ehalf * 128 + cc * 32;
                float v[32];
                tmem_ld32(tmem + acc_buf * 256u + lane_addr + (uint32_t)c0, v);
                if (cc == 3) {                                    // this warp has read its whole part of the accumulator
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bars + Wg::ACC_FREE + 8 * acc_buf);
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(tile + lane * Wg::TLD + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                __syncwarp();
                const float* trow = tile + rsub * Wg::TLD + c4;   // row 4 u' + rsub of the tile
                if (A.grads) {                                    // shared-parameter mode: raw gradients out, no update
#pragma unroll
                    for (int up = 0; up < 8; ++up) {
                        const int m = mrow0 + 4 * up;
                        const bool bias_row = !is_w2 && m == Dp;
                        if (is_w2 || m < Dp || bias_row)
                            *reinterpret_cast<float4*>(A.grads + pb + (bias_row ? (int)A.L.b1 + c0 + c4 : wbase + 4 * up * H + c0)) =
                                *reinterpret_cast<const float4*>(trow + 4 * up * Wg::TLD);
                    }
                } else {
                    // two batches of four row groups: every load of a batch is issued before its first use, so twelve
                    // (sixteen with Polyak) 16-byte loads per thread are in flight while the HBM latency elapses
#pragma unroll 1
                    for (int it0 = 0; it0 < 8; it0 += 4) {
                        int off[4];
                        float4 gr4[4], t4[4], m4[4], v4[4], g4[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int m = mrow0 + 4 * (it0 + u);
                            off[u] = wbase + 4 * (it0 + u) * H + c0;
                            if (!is_w2) off[u] = m == Dp ? (int)A.L.b1 + c0 + c4 : (m < Dp ? off[u] : -1);   // db1 from the ones column
                            gr4[u] = *reinterpret_cast<const float4*>(trow + 4 * (it0 + u) * Wg::TLD);
                            if (is_w2 || off[u] >= 0) {
                                t4[u] = ldg_plain(th + off[u]);
                                m4[u] = ldg_plain(am + off[u]);
                                v4[u] = ldg_plain(av + off[u]);
                                if (sync == 2) g4[u] = ldg_plain(tg + off[u]);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (!is_w2 && off[u] < 0) continue;
                            adam_fast(k, gr4[u].x, t4[u].x, m4[u].x, v4[u].x); adam_fast(k, gr4[u].y, t4[u].y, m4[u].y, v4[u].y);
                            adam_fast(k, gr4[u].z, t4[u].z, m4[u].z, v4[u].z); adam_fast(k, gr4[u].w, t4[u].w, m4[u].w, v4[u].w);
                            *reinterpret_cast<float4*>(th + off[u]) = t4[u];
                            *reinterpret_cast<float4*>(am + off[u]) = m4[u];
                            *reinterpret_cast<float4*>(av + off[u]) = v4[u];
                            if (sync == 1) {
                                *reinterpret_cast<float4*>(tg + off[u]) = t4[u];
                            } else if (sync == 2) {
                                const float omt = 1.0f - k.tau;
                                g4[u] = make_float4(k.tau * t4[u].x + omt * g4[u].x, k.tau * t4[u].y + omt * g4[u].y,
                                                    k.tau * t4[u].z + omt * g4[u].z, k.tau * t4[u].w + omt * g4[u].w);
                                *reinterpret_cast<float4*>(tg + off[u]) = g4[u];
                            }
                        }
                    }
                }
                __syncwarp();                                     // the tile is rewritten by the next column block
            }
            ++n;
        }
        WG_TS();
        WG_TS_PRINT("K4b epilogue (wait, adam)*", et == 0);
        if (!ok && lane == 0) atomicExch(A.error, 6);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int PASSES>
int launch_tc(const TcArgs& A, int stages, cudaStream_t s) {
    const size_t smem_f = Fwd::TOTAL + 1024, smem_w = Wg::TOTAL + 1024;
    static size_t cfg_t[kMaxDevices] = {}, cfg_o[kMaxDevices] = {}, cfg_w[kMaxDevices] = {};    // per device, not per process
    if (int rc = opt_in_dynamic_smem(reinterpret_cast<const void*>(tc_target_kernel<PASSES>), smem_f, cfg_t)) return rc;
    if (int rc = opt_in_dynamic_smem(reinterpret_cast<const void*>(tc_online_kernel<PASSES>), smem_f, cfg_o)) return rc;
    if (int rc = opt_in_dynamic_smem(reinterpret_cast<const void*>(tc_wgrad_kernel<PASSES>), smem_w, cfg_w)) return rc;
    const int grid = A.d.n_nets * A.tiles;
    if (stages & DMDQN_STAGE_TARGET) {
        tc_target_kernel<PASSES><<<2 * grid, NT_F, smem_f, s>>>(A);
        DMDQN_CUDA(cudaGetLastError());
    }
    if (stages & DMDQN_STAGE_ONLINE) {
        tc_online_kernel<PASSES><<<grid, NT_F, smem_f, s>>>(A);
        DMDQN_CUDA(cudaGetLastError());
    }
    if (stages & DMDQN_STAGE_WGRAD) {
        int n_sm = 0;
        if (int rc = device_sm_count(&n_sm)) return rc;
        const int items = A.d.n_nets * 3;
        tc_wgrad_kernel<PASSES><<<items < n_sm ? items : n_sm, Wg::NTW, smem_w, s>>>(A);
        DMDQN_CUDA(cudaGetLastError());
    }
    return DMDQN_OK;
}

}  // namespace

bool tc_supported(const dmdqn_dims& d) {
    // batch % 4: the transposed scratch is streamed in 16-byte pieces along the batch axis
    return d.hidden == H && d.obs_stride % 32 == 0 && d.obs_stride <= 96 && d.batch % 4 == 0 && d.batch <= 4096;
}

int launch_learn_tc(const dmdqn_dims& d, const dmdqn_hparams& hp, const dmdqn_replay& rp, const dmdqn_nets& nets,
                    float* metrics, char* ws, const Workspace& w, int stages, float* grads, int loss_batch,
                    cudaStream_t s) {
    TcArgs A;
    A.d = d;
    A.L = make_layout(d.obs_stride, d.hidden);
    A.rp = rp;
    A.nets = nets;
    A.gamma = (float)hp.gamma;
    A.loss = hp.loss;
    A.double_dqn = hp.double_dqn;
    A.adam_form = hp.adam_form;
    A.freq = hp.target_update_frequency > 0 ? hp.target_update_frequency : 1;
    A.loss_batch = loss_batch > 0 ? loss_batch : d.batch;
    A.tiles = (d.batch + BM - 1) / BM;
    A.lr = hp.learning_rate; A.beta1 = hp.beta1; A.beta2 = hp.beta2; A.adam_eps = hp.adam_eps; A.tau = hp.tau;
    A.rows = reinterpret_cast<const int32_t*>(ws + w.rows);
    A.act_b = reinterpret_cast<const int32_t*>(ws + w.act_b);
    A.active = reinterpret_cast<const int32_t*>(ws + w.active);
    A.step_t = reinterpret_cast<const int32_t*>(ws + w.step_t);
    A.adam_sc = reinterpret_cast<const float4*>(ws + w.adam_sc);
    A.mask2 = reinterpret_cast<uint32_t*>(ws + w.mask2);
    A.ga = reinterpret_cast<float2*>(ws + w.ga);
    A.w3_copy = reinterpret_cast<float*>(ws + w.w3_copy);
    A.r_hat = reinterpret_cast<const float*>(ws + w.r_hat);
    A.done_b = reinterpret_cast<const float*>(ws + w.done_b);
    A.y = reinterpret_cast<float*>(ws + w.y);
    A.gcoef = reinterpret_cast<float*>(ws + w.gcoef);
    A.q_all = reinterpret_cast<float*>(ws + w.q_all);
    A.q_next = reinterpret_cast<float*>(ws + w.q_next);
    A.tq_all = reinterpret_cast<float*>(ws + w.tq_all);
    A.h1 = reinterpret_cast<float*>(ws + w.h1);
    A.dh1 = reinterpret_cast<float*>(ws + w.dh1);
    A.dh2 = reinterpret_cast<float*>(ws + w.dh2);
    A.part_loss = reinterpret_cast<float*>(ws + w.part_loss);
    A.part_b3 = reinterpret_cast<float*>(ws + w.part_b3);
    A.part_w3 = reinterpret_cast<float*>(ws + w.part_w3);
    A.part_b2 = reinterpret_cast<float*>(ws + w.part_b2);
    A.metrics = metrics;
    A.grads = grads;
    A.error = reinterpret_cast<int*>(ws + w.tc_error);
    return hp.precision == DMDQN_PRECISION_TF32 ? launch_tc<1>(A, stages, s) : launch_tc<3>(A, stages, s);
}

}  // namespace dmdqn
