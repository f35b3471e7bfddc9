// Host-side launchers of the simq kernels.  All return 0 on success.
#pragma once
#include "common.cuh"

#define STAT_BLOCKS 296          // 2 x 148 SMs: partial-sum slots of every column reduction
#define MAX_CH 512

// ---- forward elementwise (kernels_elem.cu) ----
int k_colstats(const float* x, long long rows, int C, float* partials, cudaStream_t s);
int k_bn_finalize_train(const float* partials, int nparts, int C, double count, const float* gamma, const float* beta,
                        const float* conv_bias, float* rmean, float* rvar, long long* nbt, double* defer, float* mean,
                        float* invstd, float* scale, float* shift, cudaStream_t s);
// every BatchNorm of the network in one launch (block = one BN): scale/shift into bnstat[(idx*4 + 2|3) * MAX_CH]
struct BnEntry { long long gamma_off, bias_off /* -1: none */, bn_off; int ch, idx; };
// applies the running-statistics updates a pass stashed in `defer` (k_bn_finalize_train with defer != NULL)
int k_bn_running_update_all(const BnEntry* table_dev, int n, const double* defer, float* bn, long long* nbt, cudaStream_t s);
int k_bn_eval_affine_all(const float* params, const float* bn, const BnEntry* table_dev, int n, float* bnstat, cudaStream_t s);
int k_bn_eval_affine(int C, const float* gamma, const float* beta, const float* conv_bias, const float* rmean,
                     const float* rvar, float* scale, float* shift, cudaStream_t s);
// out = relu(raw*scale+shift + residual), residual: res_mode 0 none, 1 split tensor, 2 rawd*scaled+shiftd
int k_bn_apply(const float* raw, long long rows, int C, const float* scale, const float* shift, int res_mode,
               Split res, const float* rawd, const float* scaled, const float* shiftd, int pitch25, Split out,
               cudaStream_t s);
int k_stem_pool(const float* raw0, int B, const float* scale, const float* shift, Split a0, unsigned char* amax /* may be NULL */, cudaStream_t s);
int k_head_up1(const float* raw_h1, int B, const float* scale, const float* shift, Split u1, cudaStream_t s);
int k_head_t(const float* raw_h2, long long rows, const float* scale, const float* shift, const float* w3, int A,
             float* t, cudaStream_t s);
int k_head_up2(const float* t, int B, int A, const float* b3, float* q, cudaStream_t s);

// ---- DQN tail / optimiser ----
int k_dqn_tail(const float* q_s, const float* q_no, const float* q_nt, const long long* action, const float* reward,
               const unsigned char* nonfinal, float gamma, int B, int Bn, int A, int double_dqn, float* per_sample,
               long long* best_action, float* out2, float* dq, int* err_flag, cudaStream_t s);
#define SIMQ_DEVERR_ACTION_RANGE 1   // bit of the context's device error word: an action index outside [0, A*96*96)
int k_bce_tail(const float* q, const float* target, long long tstride, long long n, float* out1, float* dq, double* partials,
               cudaStream_t s);
int k_gather_rows(const float* src, const long long* idx, int n, long long row_floats, float* dst, cudaStream_t s);
int k_argmax_rows(const float* q, int B, long long row_len, long long* idx_out, cudaStream_t s);
int k_sgd_step(float* params, float* grads, float* momentum, long long n, float lr, float mom, float wd,
               float clip_norm, int first_step, double* partials, float* grad_norm_out, cudaStream_t s);

// ---- weight packing: OIHW fp32 -> split bf16 [tap][n][k] ----
// fwd: n = cout, k = cin.   bwd (dgrad): n = cin, k = cout, taps flipped.
int k_pack_weights(const float* w, int cout, int cin, int kk, Split fwd, Split bwd, cudaStream_t s);
// all convs of a network in ONE launch: table entry = one conv; `start` = its first 32 x 32 (cout x cin) tile, `total` = tiles
struct PackEntry { long long start; long long w_off; int cout, cin, kk, bk /* dgrad K stride, 0 = cout */; bf16 *fhi, *flo, *bhi, *blo; };
int k_pack_all(const float* params, const PackEntry* table_dev, int n, long long total, cudaStream_t s);

// ---- backward elementwise ----
int k_up2_adj(const float* dq, int B, int A, float* dt, cudaStream_t s);
// partials layout [STAT_BLOCKS][2+2A][32] : s1, s2, dW3[a] ; db3 in partials_b [STAT_BLOCKS][A]
int k_head2_reduce(const float* dt, const float* raw_h2, long long rows, int A, const float* scale,
                   const float* shift, const float* mean, const float* invstd, const float* w3, float* partials,
                   cudaStream_t s);
int k_reduce_partials(const float* partials, int nblk, int K, float* out, float mul, cudaStream_t s);
#define HEAD2_DY_STRIDE 64          // dy of head conv2: 32 channels zero-padded to the tensor-core tile
int k_head2_apply(const float* dt, const float* raw_h2, long long rows, int A, const float* scale, const float* shift,
                  const float* mean, const float* invstd, const float* w3, const float* sums, double count,
                  Split dy, cudaStream_t s);
int k_up1_adj(const float* du1, int B, float* g_h1, cudaStream_t s);
// mask_mode: 0 none, 1 split hi plane > 0, 2 recompute raw*scale+shift > 0
int k_bn_bwd_reduce(const float* G, long long rows, int C, int mask_mode, const bf16* mask_hi, const float* raw,
                    const float* scale, const float* shift, const float* mean, const float* invstd,
                    const float* rawd, const float* meand, const float* invstdd, float* partials, cudaStream_t s);
// sums: [3][C] = s1, s2, s2d (already reduced).  Writes dgamma/dbeta (and the ds pair) into grads.
int k_bn_bwd_apply(const float* G, long long rows, int C, int mask_mode, const bf16* mask_hi, const float* raw,
                   const float* scale, const float* shift, const float* mean, const float* invstd, const float* gamma,
                   const float* sums, double count, int pitch25, Split dy, float* dy_f32, const float* rawd,
                   const float* meand, const float* invstdd, const float* gammad, Split dyd, float* dgamma,
                   float* dbeta, float* dgammad, float* dbetad, int hi_only /* every consumer of dy is a two-term GEMM: skip the lo planes */,
                   cudaStream_t s);
int k_pool_bwd(const float* g_a0, const float* raw0, const unsigned char* amax, int B, const float* scale, const float* shift, float* dz0,
               cudaStream_t s);
int k_colsum_split(Split dy, long long rows, int C, float* partials, cudaStream_t s);   // bias grads: [STAT_BLOCKS][C]
// head2 parameter gradients out of the reduced sums [2+2A][32]: dbeta2, dgamma2, dW3, db3
int k_head2_scatter(const float* sums, int A, float* dgamma2, float* dbeta2, float* dW3, float* db3, cudaStream_t s);

// ---- debug export: internal layouts -> dense NCHW f32 ----
int k_export_p25(const float* raw, Split sp, int B, int C, float* out, cudaStream_t s);          // 24x24
int k_export_dense(const float* raw, Split sp, int B, int C, int HW, float* out, cudaStream_t s); // HWxHW NHWC
int k_import_p25(const float* nchw, int B, int C, Split sp, float* raw, cudaStream_t s);

// ---- FMA convolutions (conv_fma.cu): fp32 comparator + the stem ----
struct ConvEpilogue {
    int pitch25;              // zero the halo rows of the pitch-25 layout
    const float* add_prev;    // out += add_prev (may alias out)
    const float* add_g;       // out += (mask_hi > 0 ? add_g : 0)  (residual gradient)
    const bf16* add_g_mask;
    // ---- tcgen05 kernel only (the FMA comparator back-end takes the unfused route) ----
    const float* scale;       // v = v*scale[n] + shift[n]   (eval-mode BatchNorm folded into the epilogue)
    const float* shift;
    Split res;                // v += res.hi + res.lo        (identity shortcut)
    int relu;
    Split out_split;          // write v as split bf16 (out may then be NULL)
    float* stats;             // [stat rows][2][N] partial column sums / sums of squares of the raw accumulators
    int* stat_rows_out;       // (host) receives the number of partial rows the launch writes: one per CTA (<= 148) or one per m-tile
    // Fused reduction of those partial rows by the LAST CTA of the launch to finish (ticket counter + fences; fixed row order, so the
    // result does not depend on which CTA that is).  fin_mode 1: train-mode BatchNorm bookkeeping of the conv's output (what
    // bn_finalize_train_kernel does); 2: plain column totals into fin_out[2][N] (what reduce_partials_kernel does for the
    // BatchNorm-backward sums).  fin_ticket == NULL: no fused reduction, the caller launches those kernels.
    unsigned int* fin_ticket; int fin_mode; double fin_count;
    const float *fin_gamma, *fin_beta, *fin_bias; float *fin_rmean, *fin_rvar; long long* fin_nbt; double* fin_defer;
    float *fin_mean, *fin_invstd, *fin_scale, *fin_shift, *fin_out;
    // BatchNorm-backward statistics of the gradient this launch produces (dgrad): with dz = out * [bn_mask > 0],
    // stats rows become (sum dz, sum dz * (bn_raw - mean) * invstd) -- saves the separate reduction pass
    const float* bn_raw; const bf16* bn_mask; const float* bn_mean; const float* bn_invstd;
    int terms;                // 3: split-bf16 parity mode (lo*hi + hi*lo + hi*hi); 1: bf16 fast mode (hi*hi only)
    // scratch the split-K path of small problems may use (NULL disables it): per call, so that two contexts on two threads never share it
    float* splitk_scratch; size_t splitk_floats;
};
static inline ConvEpilogue conv_ep(int pitch25) {
    ConvEpilogue e;
    e.pitch25 = pitch25; e.add_prev = nullptr; e.add_g = nullptr; e.add_g_mask = nullptr; e.scale = nullptr; e.shift = nullptr;
    e.res.hi = nullptr; e.res.lo = nullptr; e.relu = 0; e.out_split.hi = nullptr; e.out_split.lo = nullptr; e.stats = nullptr;
    e.bn_raw = nullptr; e.bn_mask = nullptr; e.bn_mean = nullptr; e.bn_invstd = nullptr; e.terms = 3;
    e.splitk_scratch = nullptr; e.splitk_floats = 0; e.stat_rows_out = nullptr;
    e.fin_ticket = nullptr; e.fin_mode = 0; e.fin_count = 0; e.fin_gamma = e.fin_beta = e.fin_bias = nullptr; e.fin_rmean = e.fin_rvar = nullptr;
    e.fin_nbt = nullptr; e.fin_defer = nullptr; e.fin_mean = e.fin_invstd = e.fin_scale = e.fin_shift = e.fin_out = nullptr;
    return e;
}
// out[m][n] = sum_t sum_k A[m+off_t][k] * W[t][n][k]   (A, W split bf16; fp32 accumulate)
int k_conv_fma(Split A, long long rows, int K, Split W, int N, int ntaps, float* out, ConvEpilogue ep, cudaStream_t s);
// dW[co][ci][tap] (OIHW) = sum_p dY[p][co] * X[p+off_tap][ci]   (zero-inits dW itself)
int k_wgrad_fma(Split dY, Split X, long long rows, int Cout, int Cin, int ntaps, float* dW, float* scratch, cudaStream_t s);
int k_wgrad_reduce(const float* partial, int Cout, int Cin, int ntaps, int nsplit, float* dW, cudaStream_t s);
int k_stem_conv(const float* x, int x_layout, int B, int C, const float* w, float* raw0, cudaStream_t s);
int k_stem_wgrad(const float* x, int x_layout, int B, int C, const float* dy0, float* partials, float* dW,
                 cudaStream_t s);
size_t stem_wgrad_partial_floats(int C);
// tensor-core stem: im2col + GEMM (K = 49*C padded to Kp)
static inline int stem_kp(int C) { return (49 * C + 63) / 64 * 64; }
int k_stem_im2col(const float* x, int x_layout, int B, int C, int Kp, Split acol, cudaStream_t s);
int k_pack_stem(const float* w, int C, int Kp, Split out, cudaStream_t s);
int k_strip_stem(const float* tmp, int C, int Kp, float* dW, cudaStream_t s);

// ---- tcgen05 convolutions (conv_umma.cu) ----
struct UmmaTensor {          // a split tensor plus its row count / width, enough to build tensor maps
    Split t; long long rows; int cols;
};
int k_conv_umma(const UmmaTensor& A, const UmmaTensor& W, int N, int ntaps, float* out, ConvEpilogue ep, cudaStream_t s);
int k_wgrad_umma(const UmmaTensor& dY, const UmmaTensor& X, int ntaps, float* dW, float* scratch, int terms, cudaStream_t s);
int umma_init();             // resolves cuTensorMapEncodeTiled
bool umma_conv_supported(int K, int N);
int umma_conv_m_tiles(long long rows);     // rows of ConvEpilogue::stats written by k_conv_umma
bool umma_wgrad_supported(int Cout, int Cin);
size_t umma_wgrad_scratch_floats();

// ---- per-kernel-class device timing (bench.py's roofline): CUDA events around the launches of a class ----
enum { PROF_CONV = 0, PROF_WGRAD = 1, PROF_CLASSES = 2 };
void prof_enable(bool on);
void prof_mark(int cls, bool begin, double flops, cudaStream_t s, double issued = 0);   // issued: tensor-core FLOPs actually issued (terms x all rows)
int prof_collect(double* ms, double* flops, long long* launches);      // syncs the recorded events; arrays [PROF_CLASSES]
