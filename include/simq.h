/* simq — C-ABI of the B200-native DQN Q-map path (spatial-action-map Q-network).
 *
 * The reference (jimmyyhwu/spatial-intention-maps @ 336e03a) has no FFI: its boundary for this path
 * is the Python duck-type surface of networks.FCN / train.train.  Each entry point
 * below replaces one piece of that surface; the Python host side
 * (spatial_intention_maps_b200/{networks,train,policies,replay}.py) binds them with ctypes and mirrors
 * the reference names.  INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions: every function returns 0 on success, nonzero on error (message: simq_last_error(),
 * thread-local).  No exceptions, no ownership transfer: all tensors are caller-owned DEVICE
 * pointers that must stay valid until `stream` has executed the call's work.  The library owns
 * only the ctx workspace.  A ctx is bound to one device and is not thread-safe; DISTINCT contexts may be used from
 * different threads concurrently (scratch, launch counters and error words are per context, profiling state per thread).
 *
 * Flat parameter vector ("params"/"grads"/"momentum"): the 70 trainable tensors of networks.FCN in
 * reference state_dict order (networks.py:7-14, resnet.py:52-68; resnet18.fc.* excluded — it never
 * runs, resnet.py:93-104), each in PyTorch's native layout (OIHW for conv weights), fp32,
 * concatenated.  "bn" = for each of the 22 BatchNorm2d in the same order: running_mean[ch] then
 * running_var[ch], fp32.  "nbt" = 22 x int64 num_batches_tracked.  simq_layout() reports offsets.
 */
#ifndef SIMQ_H_
#define SIMQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct simq_ctx simq_ctx;
typedef void* simq_stream;            /* cudaStream_t */

#define SIMQ_N_BN 22
#define SIMQ_MAP 96                   /* envs.py:2010 state width */

/* conv back-end: the tcgen05/TMA kernels are the product; the FMA kernels are an on-device
 * fp32 comparator for tests/debug (never the default). */
enum { SIMQ_BACKEND_UMMA = 0, SIMQ_BACKEND_FMA = 1 };
/* tensor-core operand precision.  PARITY (default): every operand is a bf16 hi+lo pair and every product
 * three MMAs in every FORWARD pass (meets the 1e-3 Q-map / arg-max bar; backward GEMMs: see simq_set_backward_terms).  BF16: hi planes only, one MMA per product -- ~2.5x faster
 * conv kernels, Q-map error ~1e-2 (FAILS the parity bar; opt-in for users who train in bf16 anyway). */
enum { SIMQ_PRECISION_PARITY = 0, SIMQ_PRECISION_BF16 = 1 };
/* how the independent pieces of a step are scheduled.  LANES (default): two streams forked / joined with events -- the
 * forward on s beside the two forwards on s' (train.py:114 / :121-122), the weight-gradient GEMMs beside the
 * BatchNorm-backward + dgrad chain (train.py:132) -- so HBM-bound kernels run under tensor-bound ones; inside a
 * captured step the lanes are two branches of the CUDA graph.  SERIAL: one stream, launch order = reference order.
 * Both give bit-identical results (no atomics, per-lane scratch).  Env SIMQ_LANES=0 makes SERIAL the default. */
enum { SIMQ_SCHEDULE_SERIAL = 0, SIMQ_SCHEDULE_LANES = 1 };
/* x layouts accepted by the stem */
/* SIMQ_X_NHWC_PLUS1: [B,96,96,C+1] whose last channel is NOT a network input (train.py:145-146) */
enum { SIMQ_X_NCHW = 0, SIMQ_X_NHWC = 1, SIMQ_X_NHWC_PLUS1 = 2 };

const char* simq_last_error(void);
int simq_version(void);

/* Sizes/offsets of the flat vectors for FCN(C, A).  Any out pointer may be NULL.
 * param_offsets: 71 entries (70 tensors + end), bn_offsets: 23 entries (22 BNs + end; each BN holds
 * 2*ch floats). */
int simq_layout(int C, int A, int64_t* n_params, int64_t* n_bn, int64_t* param_offsets, int64_t* bn_offsets);

/* Replaces DQNPolicy.build_policy_nets' per-net device state (policies.py:35-42). */
int simq_ctx_create(simq_ctx** out, int device, int C, int A, int max_batch);
void simq_ctx_destroy(simq_ctx*);
int simq_set_backend(simq_ctx*, int backend);
int simq_set_precision(simq_ctx*, int mode);
int simq_set_schedule(simq_ctx*, int mode);
/* Operand terms of the BACKWARD GEMMs in parity mode.  3 = the forward's split-bf16 scheme.  2: the output
 * gradient dy contributes only its bf16 hi plane (hi*lo + hi*hi: two MMAs per product and no dy.lo loads) -- dgrad_terms for
 * the input-gradient convolutions of the residual blocks with at least dgrad2_min_planes planes (their rounding propagates down
 * the chain; the head's 1x1 convs always keep 3), wgrad_terms for the weight-gradient GEMMs (80 000-term sums: the rounding
 * averages out).  The forward passes -- everything the Q-map / arg-max parity bar covers -- are
 * never affected.  DEFAULT: (2, 2, 512) -- weight gradients and the layer-4 input gradients with two terms: measured against the
 * float64 twin of the reference the gradient error grows by <= 6 % (1.18e-2 -> 1.25e-2 flat rel-L2; the fp32 reference itself sits at
 * 0.4e-2) for -10 % step time; (3, 3, 0) restores three terms everywhere (env SIMQ_BWD_TERMS=3 makes that the default).  DESIGN.md
 * section 3 has the table. */
int simq_set_backward_terms(simq_ctx*, int dgrad_terms, int wgrad_terms, int dgrad2_min_planes);
size_t simq_workspace_bytes(const simq_ctx*);

/* Replaces FCN.forward (networks.py:16-26).  x: f32 [B,C,96,96] (NCHW) or [B,96,96,C] (NHWC);
 * q: f32 [B,A,96,96].  training!=0: batch-statistics BN, updates bn/nbt in place (also under
 * no_grad, like nn.BatchNorm2d).  save_for_backward!=0 keeps activations in the ctx for ONE
 * following simq_fcn_backward.  Weights are re-packed from `params` on every call unless
 * params_version equals the version of the previous call with the same pointer (0 = always). */
int simq_fcn_forward(simq_ctx*, const float* params, float* bn, int64_t* nbt, const float* x, int B,
                     int x_layout, int training, int save_for_backward, float* q, uint64_t params_version,
                     simq_stream stream);

/* Replaces loss.backward() through the FCN (train.py:132).  dq: f32 [B,A,96,96] dense gradient of
 * the loss w.r.t. the Q-map.  grads (flat, same layout as params) is OVERWRITTEN.  x must be the
 * tensor passed to the saving forward. */
int simq_fcn_backward(simq_ctx*, const float* params, const float* x, int x_layout, const float* dq, int B,
                      float* grads, simq_stream stream);

/* Replaces train.py:115-129: gather Q(s,a), Double-DQN target, SmoothL1, td_error, dL/dQ.
 * q_s [B,A,96,96]; q_next_online/q_next_target [Bn,A,96,96] rows = non-terminal samples in order;
 * nonfinal[B] in {0,1}; out2[0]=loss, out2[1]=mean|td|; dq (may be NULL) [B,A,96,96] is zero-filled
 * and receives the one-hot gradient. double_dqn==0 -> train.py:124. */
int simq_dqn_tail(simq_ctx*, const float* q_s, const float* q_next_online, const float* q_next_target,
                  const int64_t* action, const float* reward, const uint8_t* nonfinal, float gamma,
                  int B, int Bn, int double_dqn, float* out2, float* dq, simq_stream stream);

/* Input errors that only the device can see (an action index outside [0, A*96*96): the reference's gather raises, train.py:115)
 * never index out of bounds: the kernel sets a bit in the context's device error word and the step reports NaN loss / td_error.
 * This call synchronises `stream`, returns nonzero (message: simq_last_error) if a bit is set, and clears the word. */
int simq_check_device_errors(simq_ctx*, simq_stream stream);

/* Replaces clip_grad_norm_ + SGD.step (train.py:133-135, ctor :186) over the flat vectors.
 * first_step!=0: momentum := g (torch.optim.SGD first-step rule).  clip_norm<=0: no clipping.
 * grad_norm_out (device float, may be NULL) receives the pre-clip global L2 norm. */
int simq_sgd_step(simq_ctx*, float* params, float* grads, float* momentum, float lr, float mom, float wd,
                  float clip_norm, int first_step, float* grad_norm_out, simq_stream stream);

/* Replaces the whole of train.train (train.py:108-141) on device-resident inputs: online forward
 * on s (saving), online train-mode forward on s' (argmax), target eval forward on s', tail,
 * backward, clip, SGD.  out2 as in simq_dqn_tail.  apply_update==0 stops after the gradients
 * (data-parallel callers all-reduce `grads`, then call simq_sgd_step).  From the second call on, the ~290
 * launches of the step are replayed as ONE CUDA graph per distinct argument tuple (all pointers and scalars are part
 * of the key; SIMQ_GRAPH=0 in the environment keeps the eager launches; results are bitwise identical).  target_version: as params_version of
 * simq_fcn_forward, for the target network's packed weights (it changes only at target syncs). */
int simq_train_step(simq_ctx*, float* params, float* bn, int64_t* nbt, const float* target_params,
                    const float* target_bn, uint64_t target_version, float* grads, float* momentum,
                    const float* s, const float* s_next, int x_layout, const int64_t* action,
                    const float* reward, const uint8_t* nonfinal, int B, int Bn, float gamma, float lr,
                    float mom, float wd, float clip_norm, int first_step, int double_dqn, int apply_update,
                    float* out2, simq_stream stream);
/* The same step in two halves, for data-parallel callers that overlap the gradient all-reduce with the backward pass:
 * phase 1 = forwards, tail, and the backward through the head and layer 4 -- when it has run, every gradient from
 * resnet18.layer4.0.conv1.weight to the END of the flat vector (75 % of its bytes) is final; phase 2 = the rest of the
 * backward (layers 3..1, stem) [+ the update if apply_update].  Phase 1 followed by phase 2 with the same arguments launches
 * exactly the kernels of phase 0 (= simq_train_step): bit-identical results.  Each phase replays its own CUDA graph. */
int simq_train_step_phase(simq_ctx*, float* params, float* bn, int64_t* nbt, const float* target_params,
                          const float* target_bn, uint64_t target_version, float* grads, float* momentum,
                          const float* s, const float* s_next, int x_layout, const int64_t* action,
                          const float* reward, const uint8_t* nonfinal, int B, int Bn, float gamma, float lr,
                          float mom, float wd, float clip_norm, int first_step, int double_dqn, int apply_update,
                          float* out2, int phase, simq_stream stream);
/* Optional, one-shot: `event` (a cudaEvent_t the caller records after the host->device copy of s_next, e.g. on a copy
 * stream) is what the NEXT simq_train_step waits for before anything reads s_next -- on the lane that runs the s'
 * passes only, so the copy of s' overlaps the forward on s (the reference uploads s' with non_blocking=True for the same
 * reason, train.py:112).  Inside a captured step the wait is an external event-wait node; the handle is part of the graph
 * key.  NULL clears it. */
int simq_set_next_state_event(simq_ctx*, void* event);

/* Replay-batch assembly on the device (replaces the per-sample transform + torch.cat + H2D of
 * train.py:109-112 when the replay buffer is device-resident): dst[j][:] = src[idx[j]][:], rows of
 * row_floats fp32 (multiple of 4), idx on the device. */
int simq_gather_rows(const float* src, const int64_t* idx, int n, int64_t row_floats, float* dst, simq_stream stream);

/* Replaces nn.BCEWithLogitsLoss (mean) and its gradient, train.py:149-150.  q: n logits; target[i] at
 * target + i*target_stride; out1[0] = loss; dq[n] = dL/dq. */
int simq_bce_tail(simq_ctx*, const float* q, const float* target, int64_t target_stride, int64_t n, float* out1,
                  float* dq, simq_stream stream);

/* Replaces the whole of train.train_intention (train.py:143-158) for an FCN(C, 1) context: `state` is
 * the replay states [B,96,96,C+1] NHWC; channels 0..C-1 are the input, channel C the target intention
 * map.  Forward (train-mode BN), BCE-with-logits, backward, SGD (clip_norm<=0: none, as the reference). */
int simq_intention_step(simq_ctx*, float* params, float* bn, int64_t* nbt, float* grads, float* momentum,
                        const float* state, int B, float lr, float mom, float wd, float clip_norm, int first_step,
                        int apply_update, float* out1, simq_stream stream);

/* Replaces the greedy branch of DQNPolicy.step (policies.py:56-64): eval forward at batch B and
 * per-sample flat first-max argmax.  action_out: int64[B]; q (may be NULL) full Q-maps. */
int simq_greedy_action(simq_ctx*, const float* params, const float* bn, const float* x, int B, int x_layout,
                       int64_t* action_out, float* q, uint64_t params_version, simq_stream stream);

/* How many of this library's kernels were launched through this ctx so far. */
int64_t simq_launch_count(const simq_ctx*);

/* Per-kernel-class device timing for bench.py's roofline.  Collects what was recorded since the last
 * call into ms / flops / launches (arrays of 2: [0] tcgen05 conv+dgrad kernel, [1] tcgen05 wgrad
 * kernel incl. its split reduction; any may be NULL to discard), then enables or disables recording
 * (CUDA events around every launch of those classes, on the launching stream). */
int simq_profile(int enable, double* ms, double* flops, long long* launches);
/* Tensor-core FLOPs actually ISSUED by the launches the last simq_profile call collected (operand terms x all rows incl. the
 * pitch-25 halo rows), per class: array of 2. */
int simq_profile_issued(double* issued);

/* ---- test hooks (exercise single kernels through the C-ABI; used by tests/ only) ---- */
/* Copy an internal activation of the last forward (set 0 = saved set, 1 = scratch set) as dense
 * NCHW f32.  id: see simq_debug_tensor_name(). Returns channels*H*W per sample through *chw. */
int simq_debug_get(simq_ctx*, int set, int id, int B, float* out_nchw, int64_t* chw, simq_stream stream);
const char* simq_debug_tensor_name(int id);
/* Shifted-GEMM convolution on the pitch-25 layout with either back-end.
 * a: f32 NCHW [B,Cin,24,24]; w: f32 OIHW [Cout,Cin,k,k] (k=1|3); mode 0 = forward conv (out
 * [B,Cout,24,24]), 1 = dgrad (a is dY [B,Cout,24,24], out is dX [B,Cin,24,24]),
 * 2 = wgrad (a = X [B,Cin,24,24], a2 = dY [B,Cout,24,24], out = dW OIHW). */
int simq_test_conv(simq_ctx*, int backend, int mode, int B, int Cin, int Cout, int k, const float* a,
                   const float* a2, const float* w, float* out, simq_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* SIMQ_H_ */
