// Shared host/device helpers for the dmdqn_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/dmdqn_b200.h"

namespace dmdqn {

void set_error(const char* fmt, ...);

#define DMDQN_CHECK_ARG(cond, ...)                                  \
    do {                                                            \
        if (!(cond)) {                                              \
            ::dmdqn::set_error(__VA_ARGS__);                        \
            return DMDQN_ERR_ARG;                                   \
        }                                                           \
    } while (0)

#define DMDQN_CUDA(expr)                                                              \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            ::dmdqn::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));       \
            return DMDQN_ERR_CUDA;                                                    \
        }                                                                             \
    } while (0)

// Float offsets of W1|b1|W2|b2|W3|b3 inside one parameter block (include/dmdqn_b200.h).
struct Layout {
    int64_t w1, b1, w2, b2, w3, b3, stride;
};
__host__ __device__ inline Layout make_layout(int dp, int h) {
    Layout l;
    l.w1 = 0;
    l.b1 = (int64_t)dp * h;
    l.w2 = l.b1 + h;
    l.b2 = l.w2 + (int64_t)h * h;
    l.w3 = l.b2 + h;
    l.b3 = l.w3 + (int64_t)h * 4;
    l.stride = (l.b3 + 4 + 31) / 32 * 32;  // blocks start on 128-byte lines
    return l;
}

// Tile geometry of the fused MLP kernels for hidden width H (DESIGN.md "K3/K4 tiling").
// A CTA owns BM batch rows and all H output columns; each thread an 8x8 register tile.
template <int H>
struct Tile {
    static constexpr int WN = H / 64;               // warps along the output columns
    static constexpr int WM = (H >= 512) ? 1 : 2;   // warps along the batch rows
    static constexpr int NT = 32 * WN * WM;         // threads per CTA
    static constexpr int BM = 32 * WM;              // batch rows per CTA
    static constexpr int KC = 16;                   // K rows of the streamed operand per stage
    static constexpr int LDH = H + 4;               // smem row stride of activations
    static constexpr int LDW = H + 4;               // smem row stride of a weight chunk
};

inline int row_tiles(int batch, int bm) { return (batch + bm - 1) / bm; }

// Workspace carve-up (device scratch of sample/learn), offsets in bytes, 256-aligned.
struct Workspace {
    size_t rows, r_hat, act_b, done_b, active, step_t, y, gcoef, q_all, q_next, tq_all;
    size_t h1, dh1, dh2;              // [n_nets][B][H] floats each ([H][B] per network on the tcgen05 path)
    size_t tc_error;                  // int: set by a tcgen05 kernel whose mbarrier wait timed out
    size_t mask2;                     // tcgen05 path: relu'(h2) bits, uint32 [n_nets][B][H/32]
    size_t ga;                        // tcgen05 path: float2 [n_nets][B] {dL/dq of the taken action, action bits}
    size_t w3_copy;                   // tcgen05 path: W3 as K4a saw it, [n_nets][H][4] (K4b rebuilds dh2 while W3 is being updated)
    size_t adam_sc;                   // [n_nets] float4 {alpha_t, eps_eff, sync mode bits, -}: written by the sample kernel
    size_t part_loss;                 // [n_nets][T][8]
    size_t part_b3;                   // [n_nets][T][4]
    size_t part_w3;                   // [n_nets][T][H][4]
    size_t part_b2, part_b1;          // [n_nets][T][H]
    size_t sync;                      // tcgen05 path, int32: [0] K4b work counter, [1..3] spare, [4 ..] k4a_done[n_nets], then
                                      // k3_done[n_nets][T]; zeroed by the sample kernel (dataflow flags between the learn kernels)
    size_t total;
    int tiles;
};
inline int tile_bm(int h) { return h >= 512 ? 32 : 64; }
inline Workspace make_workspace(const dmdqn_dims& d) {
    Workspace w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t nb = (size_t)d.n_nets * d.batch;
    w.tiles = row_tiles(d.batch, tile_bm(d.hidden));
    const size_t nt = (size_t)d.n_nets * w.tiles;
    w.rows = take(nb * 4);
    w.r_hat = take(nb * 4);
    w.act_b = take(nb * 4);
    w.done_b = take(nb * 4);
    w.active = take((size_t)d.n_nets * 4);
    w.step_t = take((size_t)d.n_nets * 4);
    w.y = take(nb * 4);
    w.gcoef = take(nb * 4);
    w.q_all = take(nb * 16);
    w.q_next = take(nb * 16);
    w.tq_all = take(nb * 16);
    w.h1 = take(nb * d.hidden * 4);
    w.tc_error = take(4);
    w.adam_sc = take((size_t)d.n_nets * 16);
    w.mask2 = take(nb * (size_t)(d.hidden / 32) * 4);
    w.ga = take(nb * 8);
    w.w3_copy = take((size_t)d.n_nets * d.hidden * 16);
    w.dh1 = take(nb * d.hidden * 4);
    w.dh2 = take(nb * d.hidden * 4);
    w.part_loss = take(nt * 8 * 4);
    w.part_b3 = take(nt * 4 * 4);
    w.part_w3 = take(nt * d.hidden * 4 * 4);
    w.part_b2 = take(nt * d.hidden * 4);
    w.part_b1 = take(nt * d.hidden * 4);
    w.sync = take((4 + (size_t)d.n_nets + nt) * 4);
    w.total = off;
    return w;
}

int validate_dims(const dmdqn_dims* d);

// Per-DEVICE launch state.  The dynamic shared-memory opt-in (cudaFuncSetAttribute) and the SM count belong to a
// device, not to the process: a host that drives several GPUs from one process (one AgentGroup per device) must
// opt in on each of them.  `cache` is a zero-initialised static array owned by the call site, indexed by device
// ordinal; the worst a race between two host threads can do is set the same attribute twice.
constexpr int kMaxDevices = 64;
int device_sm_count(int* n_sm);
int opt_in_dynamic_smem(const void* kernel, size_t bytes, size_t (&cache)[kMaxDevices]);

// Launchers implemented in the .cu files.
int launch_featurize(int32_t n, const int32_t* halting, const int32_t* phase, const double* next_switch,
                     const double* phase_dur, double sim_time, const uint8_t* signal_valid,
                     const int32_t* nbr_idx, const int32_t* phase_lut, const double* snapshot,
                     double lw, double gw, double* own_out, float* obs_out, int32_t obs_out_stride,
                     double* reward_out, double* global_out, int64_t* scratch, cudaStream_t s);
int launch_featurize_alt(int32_t n, const int32_t* halting, const int32_t* phase, const double* next_switch,
                         const uint8_t* signal_valid, double sim_time, const int32_t* nbr_idx, const double* prev_own,
                         double* own_out, float* obs_out, int32_t obs_out_stride, double* reward_out, cudaStream_t s);
int launch_act(const dmdqn_dims& d, const dmdqn_nets& nets, const float* obs, int32_t stride,
               const double* eps, const uint32_t* w1, const uint32_t* w2, int32_t* actions, float* q_out,
               cudaStream_t s);
int launch_push(const dmdqn_dims& d, const dmdqn_replay& rp, const float* obs, const int32_t* act,
                const double* rew, const float* next_obs, const uint8_t* done, int32_t in_stride,
                const uint8_t* mask, cudaStream_t s);
int launch_sample(const dmdqn_dims& d, const dmdqn_hparams& hp, const dmdqn_replay& rp, const dmdqn_nets& nets,
                  const void* draws, const uint8_t* learn_mask, int advance, char* ws, const Workspace& w,
                  cudaStream_t s);
int launch_gather(const dmdqn_dims& d, const dmdqn_replay& rp, const char* ws, const Workspace& w,
                  float* states, int32_t* actions, float* rewards, float* next_states, float* dones,
                  int32_t* active_out, cudaStream_t s);
int launch_learn(const dmdqn_dims& d, const dmdqn_hparams& hp, const dmdqn_replay& rp, const dmdqn_nets& nets,
                 float* metrics, char* ws, const Workspace& w, int stages, float* grads, int loss_batch,
                 const float* apply_grads, cudaStream_t s);
bool tc_supported(const dmdqn_dims& d);
int launch_learn_tc(const dmdqn_dims& d, const dmdqn_hparams& hp, const dmdqn_replay& rp, const dmdqn_nets& nets,
                    float* metrics, char* ws, const Workspace& w, int stages, float* grads, int loss_batch,
                    cudaStream_t s);
int launch_sync_target(const dmdqn_dims& d, const dmdqn_nets& nets, const uint8_t* mask, double tau,
                       cudaStream_t s);
int launch_peer_adam(const dmdqn_dims& d, const dmdqn_hparams& hp, const dmdqn_nets& nets, const dmdqn_peers& peers,
                     const float* my_loss_src, float* loss_out, char* ws, const Workspace& w, cudaStream_t s);

}  // namespace dmdqn
