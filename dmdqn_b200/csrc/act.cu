// K2 -- batched epsilon-greedy action selection: one CTA per agent runs the batch-1 MLP
// (89 -> H -> H -> 4) as a GEMV chain with argmax and the epsilon mask fused in.
//
// Replaces DQNAgent.select_action (reference src/agents/dqn_agent.py:263-274) and
// select_greedy_action (reference src/experimental/agent.py:148-152).
// HBM bound: every action reads the agent's 4*P weight bytes once (AI 0.5 flop/B); rows of
// W are streamed with 16-byte loads, 8 in flight per thread, K split over thread groups and
// combined in a fixed order (deterministic Q-values).  Exploring agents skip the forward
// pass, as the reference does.
#include "common.cuh"

namespace dmdqn {

namespace {

constexpr int kActThreads = 256;

// out[0..H) = relu?(bias + x[0..K) . W[K][H]) ; x, out, part in shared memory.
__device__ __forceinline__ void gemv_layer(const float* __restrict__ W, const float* __restrict__ bias,
                                           const float* x, float* out, float* part, int K, int H, bool relu) {
    const int c4 = H >> 2;                 // threads covering one row of W with float4
    const int groups = kActThreads / c4;   // K is split over this many thread groups
    const int col4 = threadIdx.x % c4, grp = threadIdx.x / c4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* Wv = reinterpret_cast<const float4*>(W) + col4;
#pragma unroll 8
    for (int k = grp; k < K; k += groups) {
        const float4 w = __ldg(Wv + (size_t)k * c4);
        const float xv = x[k];
        acc.x = fmaf(xv, w.x, acc.x);
        acc.y = fmaf(xv, w.y, acc.y);
        acc.z = fmaf(xv, w.z, acc.z);
        acc.w = fmaf(xv, w.w, acc.w);
    }
    reinterpret_cast<float4*>(part)[grp * c4 + col4] = acc;
    __syncthreads();
    for (int j = threadIdx.x; j < H; j += kActThreads) {
        float s = bias[j];
        for (int g = 0; g < groups; ++g) s += part[g * H + j];
        out[j] = relu ? fmaxf(s, 0.f) : s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kActThreads)
act_kernel(dmdqn_dims d, Layout L, const float* __restrict__ theta, const float* __restrict__ obs, int stride,
           const double* __restrict__ eps, const uint32_t* __restrict__ w_explore,
           const uint32_t* __restrict__ w_action, int32_t* __restrict__ actions, float* __restrict__ q_out) {
    extern __shared__ __align__(16) float smem[];
    const int a = blockIdx.x;
    const int H = d.hidden, Dp = d.obs_stride;
    // explore iff u < eps with u = w / 2^32 (dqn_agent.py:263), exact in float64
    const bool explore = (double)w_explore[a] < eps[a] * 4294967296.0;
    if (explore) {
        if (threadIdx.x == 0) actions[a] = (int32_t)__umulhi(w_action[a], (uint32_t)d.n_actions);
        return;  // dqn_agent.py:265: no forward pass
    }
    float* xs = smem;            // [Dp]
    float* h1 = xs + Dp;         // [H]
    float* h2 = h1 + H;          // [H]
    float* part = h2 + H;        // [groups][H] = [256*4]
    const float* P = theta + (size_t)(d.n_nets == 1 ? 0 : a) * L.stride;
    for (int c = threadIdx.x; c < Dp; c += kActThreads)
        xs[c] = c < d.obs_dim ? obs[(size_t)a * stride + c] : 0.f;
    __syncthreads();
    gemv_layer(P + L.w1, P + L.b1, xs, h1, part, Dp, H, true);
    gemv_layer(P + L.w2, P + L.b2, h1, h2, part, H, H, true);

    // layer 3: q[a] = b3[a] + sum_j h2[j] * W3[j][a]; warp butterfly, then warps in order
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = threadIdx.x; j < H; j += kActThreads) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(P + L.w3) + j);
        const float hv = h2[j];
        p.x = fmaf(hv, w.x, p.x); p.y = fmaf(hv, w.y, p.y);
        p.z = fmaf(hv, w.z, p.z); p.w = fmaf(hv, w.w, p.w);
    }
    for (int off = 16; off; off >>= 1) {
        p.x += __shfl_xor_sync(0xffffffffu, p.x, off);
        p.y += __shfl_xor_sync(0xffffffffu, p.y, off);
        p.z += __shfl_xor_sync(0xffffffffu, p.z, off);
        p.w += __shfl_xor_sync(0xffffffffu, p.w, off);
    }
    float4* wsum = reinterpret_cast<float4*>(part);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = p;
    __syncthreads();
    if (threadIdx.x == 0) {
        float q[4] = {P[L.b3 + 0], P[L.b3 + 1], P[L.b3 + 2], P[L.b3 + 3]};
        for (int w = 0; w < kActThreads / 32; ++w) {
            q[0] += wsum[w].x; q[1] += wsum[w].y; q[2] += wsum[w].z; q[3] += wsum[w].w;
        }
        int best = 0;                                   // ties -> lowest index (torch.argmax)
        for (int k = 1; k < d.n_actions; ++k) if (q[k] > q[best]) best = k;
        actions[a] = best;
        if (q_out) {
            for (int k = 0; k < 4; ++k) q_out[(size_t)a * 4 + k] = q[k];
        }
    }
}

}  // namespace

int launch_act(const dmdqn_dims& d, const dmdqn_nets& nets, const float* obs, int32_t stride,
               const double* eps, const uint32_t* w1, const uint32_t* w2, int32_t* actions, float* q_out,
               cudaStream_t s) {
    const Layout L = make_layout(d.obs_stride, d.hidden);
    const size_t smem = (size_t)(d.obs_stride + 2 * d.hidden + kActThreads * 4) * sizeof(float);
    act_kernel<<<d.n_agents, kActThreads, smem, s>>>(d, L, nets.theta, obs, stride, eps, w1, w2, actions, q_out);
    DMDQN_CUDA(cudaGetLastError());
    return DMDQN_OK;
}

}  // namespace dmdqn
