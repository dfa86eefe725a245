/*
 * dmdqn_b200.h -- C ABI of the B200-native agent-side hot path of dmdqn.
 *
 * The reference (pranshu-raj-211/dmdqn) is pure Python on TensorFlow/Keras and has NO
 * FFI / plugin interface; this header is the new, thin seam.
 * Every entry point names the reference function(s) whose arithmetic it replaces; the
 * Python side (dmdqn_b200/agent.py, group.py) keeps the reference's class and method
 * names on top of these calls.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain C types only; every pointer marked "device" is a CUDA device pointer owned by
 *     the caller (the Python host allocates them as torch tensors); the library keeps no
 *     state and allocates nothing.
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work (no host sync),
 *     so they can be captured into a CUDA graph.
 *   - return value: DMDQN_OK or a negative error code; dmdqn_last_error() gives the
 *     thread-local message.  Nothing aborts, nothing throws (the reference never raises
 *     on this path either: it returns None/0 or drops the sample -- dqn_agent.py:50-54,
 *     61-62,333-335).
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     DMDQN_ERR_CUDA.
 *
 * Device data layout (see DESIGN.md "Data layout in HBM")
 *   replay ring, per agent a (deque(maxlen=C), dqn_agent.py:27-57):
 *     obs, next_obs : float [n_agents][capacity][obs_stride]   (obs_stride = obs_dim rounded
 *                      up to 16 floats: 89 -> 96 = three 128-byte lines; pad columns are 0)
 *     act           : int32 [n_agents][capacity]
 *     rew           : double[n_agents][capacity]   (the reference keeps Python floats)
 *     done          : uint8 [n_agents][capacity]
 *     n_written     : int64 [n_agents]             (slot of next write = n_written % capacity)
 *   network parameters, per network g (Keras kernel layout [in,out], y = xW + b,
 *   dqn_agent.py:153-184), one contiguous block of dmdqn_layout.stride floats:
 *     W1[obs_stride][H] | b1[H] | W2[H][H] | b2[H] | W3[H][4] | b3[4] | pad
 *   theta, theta_tgt, adam_m, adam_v all use that block layout; learn_step: int32[n_nets].
 */
#ifndef DMDQN_B200_H
#define DMDQN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMDQN_ABI_VERSION 2

#define DMDQN_OK            0
#define DMDQN_ERR_ARG      -1   /* bad dims / null pointer / unsupported size          */
#define DMDQN_ERR_CUDA     -2   /* CUDA runtime error (no device, launch failure, ...) */
#define DMDQN_ERR_WORKSPACE -3  /* workspace too small                                  */

#define DMDQN_LOSS_MSE   0      /* dqn_agent.py:141,352 (reference default)  */
#define DMDQN_LOSS_HUBER 1      /* experimental/agent.py:99, delta = 1       */

#define DMDQN_SAMPLE_INDICES     0  /* host supplies logical indices (e.g. CPython random.sample) */
#define DMDQN_SAMPLE_FISHER_YATES 1 /* uint32 draws, without replacement (partial Fisher-Yates)   */
#define DMDQN_SAMPLE_REPLACEMENT 2  /* uint32 draws, with replacement (deviation, flagged)        */

#define DMDQN_ADAM_KERAS 0      /* eps = adam_eps                       (keras 3.9.2)      */
#define DMDQN_ADAM_TORCH 1      /* eps = adam_eps * sqrt(1 - beta2^t)   (torch.optim.Adam) */

#define DMDQN_PRECISION_FP32   0 /* FFMA, fp32 accumulate                                                   */
#define DMDQN_PRECISION_TF32   1 /* tcgen05 kind::tf32, one MMA per product (~1e-3; toleranced separately)   */
#define DMDQN_PRECISION_TF32X3 2 /* tcgen05 kind::tf32, error-compensated hi/lo split, 3 MMAs per product:   */
                                 /* fp32-class accuracy on the tensor cores (H = 256 only)                   */

#define DMDQN_MAX_ACTIONS 4
#define DMDQN_OWN_DIM 17        /* order_lanes.py:430-499 */
#define DMDQN_OBS_DIM 89        /* order_lanes.py:502-555 */
#define DMDQN_PHASE_LUT 16

typedef struct dmdqn_dims {
    int32_t n_agents;    /* replay rings (intersections) held by this GPU                   */
    int32_t n_nets;      /* independent networks: n_agents, or 1 when parameters are shared */
    int32_t obs_dim;     /* D  (89)                                                         */
    int32_t obs_stride;  /* floats per stored observation row; multiple of 16, >= obs_dim   */
    int32_t hidden;      /* H in {64,128,256,512}; nn_layers = [H, H]                       */
    int32_t n_actions;   /* A <= 4                                                          */
    int32_t batch;       /* B, transitions per network per learn step                       */
    int32_t capacity;    /* C, ring capacity per agent                                      */
} dmdqn_dims;

typedef struct dmdqn_layout {   /* float offsets inside one parameter block */
    int64_t w1, b1, w2, b2, w3, b3, stride;
} dmdqn_layout;

typedef struct dmdqn_replay {   /* all device pointers */
    float*   obs;
    float*   next_obs;
    int32_t* act;
    double*  rew;
    uint8_t* done;
    int64_t* n_written;
} dmdqn_replay;

typedef struct dmdqn_nets {     /* all device pointers */
    float*   theta;
    float*   theta_tgt;
    float*   adam_m;
    float*   adam_v;
    int32_t* learn_step;
} dmdqn_nets;

typedef struct dmdqn_hparams {
    double  gamma;                    /* agent_config.yaml: gamma (rounded to fp32 on use) */
    double  learning_rate;            /* agent_config.yaml: learning_rate                  */
    double  beta1, beta2, adam_eps;   /* Keras Adam defaults 0.9, 0.999, 1e-7; the step    */
                                      /* size alpha_t is evaluated in float64 on the device */
    double  tau;                      /* < 0: hard sync every target_update_frequency      */
                                      /* >= 0: Polyak every step (dqn_agent.py:389-399)    */
    int32_t target_update_frequency;  /* agent_config.yaml: target_update_frequency        */
    int32_t loss;                     /* DMDQN_LOSS_*                                      */
    int32_t normalize_rewards;        /* 1 = per-batch z-score (dqn_agent.py:66-69)        */
    int32_t double_dqn;               /* 1 = dqn_agent.py:342-345, 0 = experimental/agent.py:166-167 */
    int32_t adam_form;                /* DMDQN_ADAM_*                                      */
    int32_t sample_mode;              /* DMDQN_SAMPLE_*                                    */
    int32_t precision;                /* DMDQN_PRECISION_*                                 */
} dmdqn_hparams;

/* Per-network learn outputs, 8 floats each: loss, q_mean, q_std (population, over the
 * [B,A] online Q of the sampled states), action histogram[4], learned flag (0/1)
 * (dqn_agent.py:359-370). */
#define DMDQN_METRICS_STRIDE 8

const char* dmdqn_last_error(void);
int dmdqn_abi_version(void);

/* Parameter block layout for the given dims. */
int dmdqn_param_layout(const dmdqn_dims* dims, dmdqn_layout* out);

/* Bytes of device scratch dmdqn_sample / dmdqn_learn need for these dims. */
int dmdqn_workspace_bytes(const dmdqn_dims* dims, size_t* out_bytes);

/* K0 -- observation + reward featurisation for all intersections.
 * Replaces get_own_state / _get_neighbor_info / build_state_vector
 * (src/experimental/order_lanes.py:392-555) and calculate_local_reward /
 * calculate_global_reward + the 0.3/0.7 mix (src/scripts/train.py:159-165,241,251-254).
 *   halting      device int32 [n][12]   slot = dir*3+lane (n,s,e,w); -1 = lane absent
 *   phase        device int32 [n]       SUMO phase index
 *   next_switch  device double[n], phase_dur device double[n], sim_time: TraCI readings
 *   signal_valid device uint8 [n]       0 -> phase/time stay [0,0,0,0,-1] (shipped behaviour)
 *   nbr_idx      device int32 [n][4]    neighbour rows in n,s,e,w order, -1 = none
 *   phase_lut    device int32 [16]      phase index -> one-hot slot 0..3 or -1
 *   snapshot     device double[n][17] or NULL: neighbour blocks come from this
 *                global_state snapshot (order_lanes.py:547-548) instead of the live blocks
 *   own_out      device double[n][17]   (the caller's next global_state)
 *   obs_out      device float [n][obs_out_stride]  first 89 columns written, rest zeroed
 *   reward_out   device double[n], global_out device double[1]: reward from THESE readings
 *                used as the pre-step state (train.py:241).
 *   scratch      device int64 [2], zero on entry; left zero on return. */
int dmdqn_featurize(int32_t n, const int32_t* halting, const int32_t* phase,
                    const double* next_switch, const double* phase_dur, double sim_time,
                    const uint8_t* signal_valid, const int32_t* nbr_idx, const int32_t* phase_lut,
                    const double* snapshot, double local_weight, double global_weight,
                    double* own_out, float* obs_out, int32_t obs_out_stride,
                    double* reward_out, double* global_out, int64_t* scratch, void* stream);

/* K0, second layout -- replaces SumoTrafficEnvironment._get_local_observation / _get_observations /
 * _get_neighbor_presence_vector / _calculate_rewards (reference src/agents/sumo_env.py:532-580,582-631,633-645,
 * 652-679): the 74-dim observation local(14) | presence(4: N,E,S,W) | 4 x neighbour local(14), pad -1.0, and the
 * queue-reduction reward.
 *   halting      device int32 [n][12]   queues in N,E,S,W order x 3 lanes; -2 = PAD lane (-> 0.0), -1 = failed read (-> -1.0)
 *   phase        device int32 [n], next_switch device double[n], sim_time
 *   signal_valid device uint8 [n]       0 -> both signal slots keep the -1.0 padding value
 *   nbr_idx      device int32 [n][4]    neighbour rows in N,E,S,W order, -1 = none / not controlled
 *   prev_own     device double[n][14] or NULL (first step): the previous call's own_out
 *   own_out      device double[n][14] or NULL
 *   obs_out      device float [n][obs_out_stride]  first 74 columns written, rest zeroed
 *   reward_out   device double[n] or NULL: sum max(0, prev queues) - sum max(0, current queues); 0 when prev_own is NULL */
#define DMDQN_OWN_ALT_DIM 14
#define DMDQN_OBS_ALT_DIM 74
int dmdqn_featurize_alt(int32_t n, const int32_t* halting, const int32_t* phase, const double* next_switch,
                        const uint8_t* signal_valid, double sim_time, const int32_t* nbr_idx, const double* prev_own,
                        double* own_out, float* obs_out, int32_t obs_out_stride, double* reward_out, void* stream);

/* K2 -- batched epsilon-greedy action selection, one observation per agent.
 * Replaces DQNAgent.select_action (src/agents/dqn_agent.py:263-274; the epsilon schedule
 * :258-261 stays on the host) and select_greedy_action (experimental/agent.py:148-152).
 *   obs        device float [n_agents][obs_in_stride]
 *   eps        device double[n_agents]   host-computed epsilon per agent
 *   w_explore, w_action  device uint32[n_agents]  supplied draws: explore iff
 *              w_explore < eps * 2^32; random action = (w_action * A) >> 32
 *   actions_out device int32[n_agents];  q_out device float[n_agents][4] or NULL
 *              (q_out rows of exploring agents are left untouched: the reference skips
 *              the forward pass when exploring). */
int dmdqn_act(const dmdqn_dims* dims, const dmdqn_nets* nets, const float* obs, int32_t obs_in_stride,
              const double* eps, const uint32_t* w_explore, const uint32_t* w_action,
              int32_t* actions_out, float* q_out, void* stream);

/* K1a -- append one transition per agent to its ring.
 * Replaces ReplayBuffer.add / DQNAgent.remember / store_experience
 * (src/agents/dqn_agent.py:31-57,306-325).  mask (device uint8[n_agents] or NULL)
 * selects the agents that store. */
int dmdqn_push(const dmdqn_dims* dims, const dmdqn_replay* replay, const float* obs,
               const int32_t* act, const double* rew, const float* next_obs, const uint8_t* done,
               int32_t in_stride, const uint8_t* mask, void* stream);

/* K1b -- draw the batch of every network and z-score its rewards.
 * Replaces ReplayBuffer.sample (src/agents/dqn_agent.py:59-85) up to the gather.
 *   draws  device [n_nets][batch]: int32 logical indices (0 = oldest) or uint32 words,
 *          per hp->sample_mode
 *   learn_mask device uint8[n_nets] or NULL; a network is "active" iff its mask is set
 *          and its ring(s) hold >= batch transitions (dqn_agent.py:61-62,333-335).
 *   advance_step: 1 when called as the first stage of a learn step (bumps learn_step of
 *          active networks, dqn_agent.py:359); 0 for a stand-alone sample().
 * Results stay in the workspace (rows, r_hat, actions, dones, active flags). */
int dmdqn_sample(const dmdqn_dims* dims, const dmdqn_hparams* hp, const dmdqn_replay* replay,
                 const dmdqn_nets* nets, const void* draws, const uint8_t* learn_mask,
                 int32_t advance_step, void* workspace, size_t workspace_bytes, void* stream);

/* Materialise the sampled batch (the five tensors ReplayBuffer.sample returns,
 * dqn_agent.py:80-85) from the workspace of the last dmdqn_sample.
 *   states, next_states device float[n_nets][batch][obs_dim] (dense, no padding)
 *   actions int32, rewards float (z-scored), dones float: device [n_nets][batch]
 *   active_out device int32[n_nets] or NULL. */
int dmdqn_gather(const dmdqn_dims* dims, const dmdqn_replay* replay, const void* workspace,
                 size_t workspace_bytes, float* states, int32_t* actions, float* rewards,
                 float* next_states, float* dones, int32_t* active_out, void* stream);

/* K1b + K3 + K4 -- one Double-DQN learn step for every active network.
 * Replaces DQNAgent.learn / replay (src/agents/dqn_agent.py:328-380,428-434):
 * sample, target-net forward + online argmax + TD target, online forward, MSE/Huber,
 * backward, Adam, hard/Polyak target sync.
 *   metrics_out device float[n_nets][DMDQN_METRICS_STRIDE].
 * On the tcgen05 path the four kernels of one call form a dataflow chain (programmatic dependent launches:
 * K3 starts under the sample kernel's tail, K4a / K4b take over SMs while the previous kernel is still running
 * and wait on per-tile / per-network flags in the workspace, which the sample kernel resets).  Results are
 * bit-identical to the stage-by-stage form below; the workspace must not be shared by two calls in flight. */
int dmdqn_learn(const dmdqn_dims* dims, const dmdqn_hparams* hp, const dmdqn_replay* replay,
                const dmdqn_nets* nets, const void* draws, const uint8_t* learn_mask,
                float* metrics_out, void* workspace, size_t workspace_bytes, void* stream);

/* The same step, one stage per call, so a caller can bracket each kernel with CUDA events
 * (bench.py roofline): stages is a bit mask of DMDQN_STAGE_*; all four in order == dmdqn_learn (same bits;
 * kernels of different calls are ordered by the stream alone, only a call with all four stages is chained). */
#define DMDQN_STAGE_SAMPLE 1   /* K1b */
#define DMDQN_STAGE_TARGET 2   /* K3  */
#define DMDQN_STAGE_ONLINE 4   /* K4a */
#define DMDQN_STAGE_WGRAD  8   /* K4b */
int dmdqn_learn_stages(const dmdqn_dims* dims, const dmdqn_hparams* hp, const dmdqn_replay* replay,
                       const dmdqn_nets* nets, const void* draws, const uint8_t* learn_mask,
                       float* metrics_out, void* workspace, size_t workspace_bytes, int32_t stages,
                       void* stream);

/* remember + replay of every agent with HOST buffers -- what the reference's train loop does per step
 * (src/scripts/train.py:274-292: agent.remember(...) then agent.replay() for every agent), as one call:
 * the step's inputs sit in ONE pinned host block (struct of arrays at the byte offsets below, each a
 * multiple of 16), which is copied to its device mirror with a single cudaMemcpyAsync; then dmdqn_push,
 * dmdqn_learn and a copy of the metrics back to host memory are queued on `stream`.  Nothing is synchronised:
 * the caller waits on the stream before reading `metrics_host` (the losses replay() returns). */
typedef struct dmdqn_step_block {
    size_t bytes;                 /* size of the host / device block */
    size_t obs_off, next_obs_off; /* float [n_agents][in_stride] */
    size_t act_off;               /* int32 [n_agents] */
    size_t rew_off;               /* double[n_agents] */
    size_t done_off;              /* uint8 [n_agents] */
    size_t draws_off;             /* [n_nets][batch] words / indices per hp->sample_mode */
    int32_t in_stride;
} dmdqn_step_block;
int dmdqn_step_host(const dmdqn_dims* dims, const dmdqn_hparams* hp, const dmdqn_replay* replay,
                    const dmdqn_nets* nets, const dmdqn_step_block* blk, const void* host_block, void* device_block,
                    float* metrics_dev, float* metrics_host, void* workspace, size_t workspace_bytes, void* stream);

/* Shared-parameter mode across ranks (n_nets == 1; not in the reference, BASELINE.json cfg5):
 * the same step but K4b stores dL/dtheta into grads_out[n_nets][layout.stride] instead of
 * applying Adam, with the loss mean taken over global_batch (= batch * ranks), so that the
 * caller can sum the blocks over ranks (ncclAllReduce) and finish with dmdqn_adam_apply,
 * which reads the step counters dmdqn_learn_grads left in the workspace. */
int dmdqn_learn_grads(const dmdqn_dims* dims, const dmdqn_hparams* hp, const dmdqn_replay* replay,
                      const dmdqn_nets* nets, const void* draws, const uint8_t* learn_mask,
                      int32_t global_batch, float* grads_out, float* metrics_out, void* workspace,
                      size_t workspace_bytes, void* stream);
int dmdqn_adam_apply(const dmdqn_dims* dims, const dmdqn_hparams* hp, const dmdqn_nets* nets,
                     const float* grads, void* workspace, size_t workspace_bytes, void* stream);

/* K5, shared-parameter mode over NVLink peer memory (BASELINE.json cfg5): ONE kernel per
 * rank that (1) announces its gradient block of this step to every peer (release store of `epoch` into the peers'
 * flag arrays), (2) waits for the peers' announcements, (3) reads every rank's block through peer pointers, sums
 * them in rank order 0..world-1 (the same order on every rank: replicas stay bit-identical) and (4) applies Adam +
 * the target sync to the local replica -- the all-reduce and the optimizer step of dmdqn_learn_grads ->
 * ncclAllReduce -> dmdqn_adam_apply in one launch, with no host round trip and no NCCL on the path.
 * grads[p] / loss[p] / flags[p] are device pointers valid on THIS device: the rank's own buffers at [rank], the
 * peers' through dmdqn_ipc_open.  The caller double-buffers grads / loss by epoch parity (a peer may already be
 * writing step t+1 while this rank still reads step t); flags[p] is uint32[DMDQN_MAX_PEERS][DMDQN_PEER_BLOCKS],
 * zero-initialised once; epoch = 1, 2, 3, ... identical on every rank.  my_loss_src is this rank's share of the loss
 * (metrics_out[0] of dmdqn_learn_grads), published to loss[rank] by the kernel; loss_out receives the sum.
 * A peer that never arrives ends in DMDQN_PEER_TIMEOUT in the workspace error flag (no update applied), not in a hang. */
#define DMDQN_MAX_PEERS 8
#define DMDQN_PEER_TIMEOUT 77     /* value of the workspace error flag after a peer never arrived */
#define DMDQN_PEER_BLOCKS 128
typedef struct dmdqn_peers {
    const float* grads[DMDQN_MAX_PEERS];
    float* loss[DMDQN_MAX_PEERS];
    uint32_t* flags[DMDQN_MAX_PEERS];
    int32_t rank, world;
    uint32_t epoch;
} dmdqn_peers;
int dmdqn_allreduce_adam(const dmdqn_dims* dims, const dmdqn_hparams* hp, const dmdqn_nets* nets,
                         const dmdqn_peers* peers, const float* my_loss_src, float* loss_out, void* workspace,
                         size_t workspace_bytes, void* stream);
/* CUDA IPC plumbing for the peer pointers above: export the allocation that contains dev_ptr (64-byte handle +
 * offset of dev_ptr inside it), map a peer's export on the current device (peer access is enabled lazily), unmap. */
int dmdqn_ipc_export(const void* dev_ptr, void* handle64, uint64_t* offset);
int dmdqn_ipc_open(const void* handle64, uint64_t offset, void** dev_ptr);
int dmdqn_ipc_close(void* dev_ptr, uint64_t offset);

/* Debug / parity views into the workspace after dmdqn_learn (device pointers into it):
 * y[n_nets][B], q_all[n_nets][B][4] (online Q of s), q_next[n_nets][B][4] (online Q of s'),
 * tq_all[n_nets][B][4] (target Q of s'), rows int32[n_nets][B] (agent*C + slot). */
typedef struct dmdqn_debug_views {
    const float* y; const float* q_all; const float* q_next; const float* tq_all;
    const int32_t* rows; const float* r_hat; const int32_t* active;
    const int32_t* tc_error;   /* != 0: a tcgen05 kernel's bounded mbarrier wait expired (results invalid) */
    const float* dh1; const float* dh2;   /* [n_nets][B][H] activation gradients of the last step */
    /* tcgen05 path: dh2 is never materialised; relu'(h2) as bits, uint32 [n_nets][B][H/32]
     * (bit j of word w of row i = column 32 w + j), dh2 = bit ? g_i * W3[col][a_i] : 0 */
    const uint32_t* relu2_bits;
} dmdqn_debug_views;
int dmdqn_debug(const dmdqn_dims* dims, void* workspace, size_t workspace_bytes, dmdqn_debug_views* out);

/* theta_tgt <- theta for the networks whose mask is set (NULL = all).
 * Replaces DQNAgent.update_target_network (dqn_agent.py:382-384) when called explicitly. */
int dmdqn_sync_target(const dmdqn_dims* dims, const dmdqn_nets* nets, const uint8_t* mask,
                      double tau, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DMDQN_B200_H */
