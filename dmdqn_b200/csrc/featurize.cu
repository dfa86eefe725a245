// K0 -- observation + reward featurisation (one warp per intersection).
//
// Replaces get_own_state / _get_neighbor_info / build_state_vector
// (reference src/experimental/order_lanes.py:392-555) and the reward helpers
// (reference src/scripts/train.py:159-165,241,251-254).  HBM/latency bound: 64 B in and
// 89*4 + 17*8 + 8 B out per intersection; neighbour own-blocks are recomputed from the raw
// readings (17 values) instead of being exchanged, so one launch has no grid-wide
// dependency except the scalar global reward, finished by the last CTA to retire.
#include "common.cuh"

namespace dmdqn {

namespace {

struct FeatIn {
    const int32_t* halting;
    const int32_t* phase;
    const double* next_switch;
    const double* phase_dur;
    const uint8_t* signal_valid;
    const int32_t* phase_lut;
    double sim_time;
};

// Element e (0..16) of intersection a's own block, float64 (order_lanes.py:430-499).
__device__ __forceinline__ double own_elem(const FeatIn& in, int a, int e) {
    if (e < 12) return (double)in.halting[a * 12 + e];  // -1 = absent lane stays -1.0 (:439)
    const bool valid = in.signal_valid[a] != 0;         // traffic-light branch ran (:468)
    if (e < 16) {                                       // phase one-hot, zeros if unmapped (:471)
        const int p = in.phase[a];
        const int slot = (valid && p >= 0 && p < DMDQN_PHASE_LUT) ? in.phase_lut[p] : -1;
        return slot == e - 12 ? 1.0 : 0.0;
    }
    // time_spent = dur - (next_switch - now); a negative value trips the reference's assert
    // inside a swallowed try, leaving -1.0 (:478-486).
    const double calc = __dsub_rn(in.phase_dur[a], __dsub_rn(in.next_switch[a], in.sim_time));
    return (valid && calc >= 0.0) ? calc : -1.0;
}

constexpr int kWarpsPerCta = 4;

__global__ void __launch_bounds__(32 * kWarpsPerCta)
featurize_kernel(int n, FeatIn in, const int32_t* __restrict__ nbr_idx, const double* __restrict__ snapshot,
                 double lw, double gw, double* __restrict__ own_out, float* __restrict__ obs_out,
                 int obs_stride, double* reward_out, double* __restrict__ global_out,
                 unsigned long long* scratch) {
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    long long local = 0;
    if (a < n) {
        if (lane < DMDQN_OWN_DIM && own_out) own_out[a * DMDQN_OWN_DIM + lane] = own_elem(in, a, lane);
        for (int c = lane; c < obs_stride; c += 32) {
            double v = 0.0;  // pad columns
            if (c < DMDQN_OWN_DIM) {
                v = own_elem(in, a, c);                              // live own block (:526-531)
            } else if (c < DMDQN_OWN_DIM + 4) {
                v = nbr_idx[a * 4 + (c - DMDQN_OWN_DIM)] >= 0 ? 1.0 : 0.0;  // presence n,s,e,w
            } else if (c < DMDQN_OBS_DIM) {
                const int k = (c - DMDQN_OWN_DIM - 4) / DMDQN_OWN_DIM;
                const int e = (c - DMDQN_OWN_DIM - 4) % DMDQN_OWN_DIM;
                const int nb = nbr_idx[a * 4 + k];
                if (nb < 0) v = -1.0;                                // padding block (:516,550)
                else v = snapshot ? snapshot[nb * DMDQN_OWN_DIM + e] : own_elem(in, nb, e);
            }
            obs_out[(size_t)a * obs_stride + c] = (float)v;          // caller's fp32 cast (train.py:220)
        }
        if (lane < 12) local = in.halting[a * 12 + lane];
    }
    for (int off = 16; off; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
    __shared__ long long cta_sum[kWarpsPerCta];
    __shared__ bool is_last;
    if (lane == 0) {
        cta_sum[threadIdx.x >> 5] = (a < n) ? local : 0;
        // local_j = -1.0 * sum(own[:12]); parked in reward_out until the global sum is known
        if (a < n) reward_out[a] = __dmul_rn(-1.0, (double)local);
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long s = 0;
        for (int w = 0; w < kWarpsPerCta; ++w) s += cta_sum[w];
        atomicAdd(&scratch[0], (unsigned long long)s);   // integer sum: exact in any order
        __threadfence();
        is_last = atomicAdd(&scratch[1], 1ull) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const long long total = (long long)atomicAdd(&scratch[0], 0ull);
    const double glob = __dmul_rn(-1.0, (double)total);            // train.py:163-165
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double loc = __ldcg(&reward_out[j]);
        // two multiplies and one add, no FMA contraction (train.py:254)
        reward_out[j] = __dadd_rn(__dmul_rn(lw, loc), __dmul_rn(gw, glob));
    }
    if (threadIdx.x == 0) {
        *global_out = glob;
        scratch[0] = 0;
        scratch[1] = 0;
    }
}

// ---- alt contract: SumoTrafficEnvironment's 74-dim observation and queue-reduction reward
// (reference src/agents/sumo_env.py:532-580 local block, :582-631 assembly, :633-645 presence, :652-679 reward).
// local(14) = 12 queues in N,E,S,W order x 3 lanes (code -2 = PAD lane -> 0.0 (:545-549), -1 = failed read keeps
// the -1.0 padding (:555-557)) | phase index | max(0, nextSwitch - now); a junction without a readable signal keeps
// -1.0 in both signal slots (:560-575).
__device__ __forceinline__ double own_alt_elem(const int32_t* __restrict__ halting, const int32_t* __restrict__ phase,
                                               const double* __restrict__ next_switch, const uint8_t* __restrict__ signal_valid,
                                               double sim_time, int a, int e) {
    if (e < 12) {
        const int h = halting[a * 12 + e];
        return h == -2 ? 0.0 : (double)h;
    }
    if (!signal_valid[a]) return -1.0;
    if (e == 12) return (double)phase[a];
    return fmax(0.0, __dsub_rn(next_switch[a], sim_time));
}

__global__ void __launch_bounds__(32 * kWarpsPerCta)
featurize_alt_kernel(int n, const int32_t* __restrict__ halting, const int32_t* __restrict__ phase,
                     const double* __restrict__ next_switch, const uint8_t* __restrict__ signal_valid, double sim_time,
                     const int32_t* __restrict__ nbr_idx, const double* __restrict__ prev_own,
                     double* __restrict__ own_out, float* __restrict__ obs_out, int obs_stride, double* __restrict__ reward_out) {
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (a >= n) return;
    if (lane < DMDQN_OWN_ALT_DIM && own_out)
        own_out[a * DMDQN_OWN_ALT_DIM + lane] = own_alt_elem(halting, phase, next_switch, signal_valid, sim_time, a, lane);
    for (int c = lane; c < obs_stride; c += 32) {
        double v = 0.0;  // pad columns
        if (c < DMDQN_OWN_ALT_DIM) {
            v = own_alt_elem(halting, phase, next_switch, signal_valid, sim_time, a, c);
        } else if (c < DMDQN_OWN_ALT_DIM + 4) {
            v = nbr_idx[a * 4 + (c - DMDQN_OWN_ALT_DIM)] >= 0 ? 1.0 : 0.0;   // presence N,E,S,W (:633-645)
        } else if (c < DMDQN_OBS_ALT_DIM) {
            const int k = (c - DMDQN_OWN_ALT_DIM - 4) / DMDQN_OWN_ALT_DIM, e = (c - DMDQN_OWN_ALT_DIM - 4) % DMDQN_OWN_ALT_DIM;
            const int nb = nbr_idx[a * 4 + k];
            v = nb < 0 ? -1.0 : own_alt_elem(halting, phase, next_switch, signal_valid, sim_time, nb, e);   // (:606-617)
        }
        obs_out[(size_t)a * obs_stride + c] = (float)v;
    }
    if (reward_out) {   // sum max(0, prev q) - sum max(0, curr q) (:672-677); 0 on the first step (:655-656)
        double d = 0.0;
        if (prev_own && lane < 12) {
            const double cur = own_alt_elem(halting, phase, next_switch, signal_valid, sim_time, a, lane);
            d = fmax(0.0, prev_own[a * DMDQN_OWN_ALT_DIM + lane]) - fmax(0.0, cur);   // small integers: exact in any order
        }
        for (int off = 16; off; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
        if (lane == 0) reward_out[a] = d;
    }
}

}  // namespace

int launch_featurize_alt(int32_t n, const int32_t* halting, const int32_t* phase, const double* next_switch,
                         const uint8_t* signal_valid, double sim_time, const int32_t* nbr_idx, const double* prev_own,
                         double* own_out, float* obs_out, int32_t obs_out_stride, double* reward_out, cudaStream_t s) {
    const int grid = (n + kWarpsPerCta - 1) / kWarpsPerCta;
    featurize_alt_kernel<<<grid, 32 * kWarpsPerCta, 0, s>>>(n, halting, phase, next_switch, signal_valid, sim_time, nbr_idx,
                                                           prev_own, own_out, obs_out, obs_out_stride, reward_out);
    DMDQN_CUDA(cudaGetLastError());
    return DMDQN_OK;
}

int launch_featurize(int32_t n, const int32_t* halting, const int32_t* phase, const double* next_switch,
                     const double* phase_dur, double sim_time, const uint8_t* signal_valid,
                     const int32_t* nbr_idx, const int32_t* phase_lut, const double* snapshot,
                     double lw, double gw, double* own_out, float* obs_out, int32_t obs_out_stride,
                     double* reward_out, double* global_out, int64_t* scratch, cudaStream_t s) {
    FeatIn in{halting, phase, next_switch, phase_dur, signal_valid, phase_lut, sim_time};
    const int grid = (n + kWarpsPerCta - 1) / kWarpsPerCta;
    featurize_kernel<<<grid, 32 * kWarpsPerCta, 0, s>>>(n, in, nbr_idx, snapshot, lw, gw, own_out, obs_out,
                                                       obs_out_stride, reward_out, global_out,
                                                       reinterpret_cast<unsigned long long*>(scratch));
    DMDQN_CUDA(cudaGetLastError());
    return DMDQN_OK;
}

}  // namespace dmdqn
