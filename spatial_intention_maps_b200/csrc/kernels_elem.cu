// Elementwise / reduction kernels of the simq path: BatchNorm statistics and application, the stem
// max-pool, the bilinear decoder, the DQN tail, clipped momentum-SGD, weight packing, and their
// backward counterparts.  All of these are HBM-bound: one pass, 16-byte vector accesses along the
// channel (innermost, NHWC) axis, deterministic two-stage reductions (no atomics).
#include "kernels.h"
#include <math.h>

#define SGD_BLOCKS 592
#define BN_EPS 1e-5
#define BN_MOM 0.1

static inline int grid_for(long long n, int block) {
    long long g = (n + block - 1) / block;
    return (int)(g < 1 ? 1 : g);
}

// ------------------------------------------------------------------------------------------
// per-channel sum / sum of squares over the rows of an fp32 [rows][C] matrix
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colstats_kernel(const float* __restrict__ x, long long rows, int C,
                                                       float* __restrict__ partials) {
    __shared__ float sm[2][1024];
    const int C4 = C >> 2, rpi = 256 / C4;
    const int c4 = threadIdx.x % C4, rl = threadIdx.x / C4;
    float4 s = make_float4(0, 0, 0, 0), ss = make_float4(0, 0, 0, 0);
    for (long long r = (long long)blockIdx.x * rpi + rl; r < rows; r += (long long)gridDim.x * rpi) {
        float4 v = __ldg(reinterpret_cast<const float4*>(x + r * C) + c4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        ss.x += v.x * v.x; ss.y += v.y * v.y; ss.z += v.z * v.z; ss.w += v.w * v.w;
    }
    float* a = &sm[0][rl * C + c4 * 4];
    float* b = &sm[1][rl * C + c4 * 4];
    a[0] = s.x; a[1] = s.y; a[2] = s.z; a[3] = s.w;
    b[0] = ss.x; b[1] = ss.y; b[2] = ss.z; b[3] = ss.w;
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float t0 = 0, t1 = 0;
        for (int r = 0; r < rpi; ++r) { t0 += sm[0][r * C + c]; t1 += sm[1][r * C + c]; }
        partials[((size_t)blockIdx.x * 2 + 0) * C + c] = t0;
        partials[((size_t)blockIdx.x * 2 + 1) * C + c] = t1;
    }
}

int k_colstats(const float* x, long long rows, int C, float* partials, cudaStream_t s) {
    colstats_kernel<<<STAT_BLOCKS, 256, 0, s>>>(x, rows, C, partials);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// nn.BatchNorm2d train-mode bookkeeping (eps 1e-5, momentum 0.1, unbiased var into running_var,
// num_batches_tracked += 1).  `raw` excludes the conv bias: the batch mean of the true conv output is
// mean_raw + bias, and the bias cancels in the normalised value.
// One block = 32 channels x 32 partial-row groups (coalesced 128-byte reads of the partial rows, 4 loads in
// flight per thread: the kernel is latency-bound); the group sums are combined in a fixed order in shared
// memory, so the result is deterministic.
__global__ void __launch_bounds__(1024) bn_finalize_train_kernel(const float* __restrict__ partials, int nparts, int C, double count,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         const float* __restrict__ conv_bias, float* rmean, float* rvar,
                                         long long* nbt, double* defer, float* mean, float* invstd, float* scale, float* shift) {
    __shared__ double sS[32][33], sSS[32][33];
    const int cl = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    double S = 0, SS = 0;
    if (c < C) {
        int b = r;
        for (; b + 96 < nparts; b += 128) {
            float a0 = partials[((size_t)b * 2) * C + c], q0 = partials[((size_t)b * 2 + 1) * C + c];
            float a1 = partials[((size_t)(b + 32) * 2) * C + c], q1 = partials[((size_t)(b + 32) * 2 + 1) * C + c];
            float a2 = partials[((size_t)(b + 64) * 2) * C + c], q2 = partials[((size_t)(b + 64) * 2 + 1) * C + c];
            float a3 = partials[((size_t)(b + 96) * 2) * C + c], q3 = partials[((size_t)(b + 96) * 2 + 1) * C + c];
            S += ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
            SS += ((double)q0 + (double)q1) + ((double)q2 + (double)q3);
        }
        for (; b < nparts; b += 32) {
            S += (double)partials[((size_t)b * 2 + 0) * C + c];
            SS += (double)partials[((size_t)b * 2 + 1) * C + c];
        }
    }
    sS[r][cl] = S; sSS[r][cl] = SS;
    __syncthreads();
    if (r != 0 || c >= C) return;
    S = 0; SS = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) { S += sS[i][cl]; SS += sSS[i][cl]; }
    bn_finalize_channel(S, SS, count, c, gamma, beta, conv_bias, rmean, rvar, defer, MAX_CH, mean, invstd, scale, shift);
    if (defer) return;      // a concurrent pass owns the running statistics right now: bn_running_update_all applies these later
    if (c == 0 && nbt) nbt[0] += 1;
}

int k_bn_finalize_train(const float* partials, int nparts, int C, double count, const float* gamma, const float* beta,
                        const float* conv_bias, float* rmean, float* rvar, long long* nbt, double* defer, float* mean,
                        float* invstd, float* scale, float* shift, cudaStream_t s) {
    bn_finalize_train_kernel<<<ceil_div(C, 32), 1024, 0, s>>>(partials, nparts, C, count, gamma, beta, conv_bias, rmean, rvar,
                                                              nbt, defer, mean, invstd, scale, shift);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// The running-statistics update of every BatchNorm of one train-mode pass from the stash bn_finalize_train_kernel left
// in `defer` ([n][2][MAX_CH] doubles: batch mean incl. conv bias, unbiased batch variance): the same arithmetic, applied
// after the pass that ran concurrently (one block per BN).
__global__ void __launch_bounds__(MAX_CH) bn_running_update_all_kernel(const BnEntry* __restrict__ table, const double* __restrict__ defer,
                                                                       float* __restrict__ bn, long long* __restrict__ nbt) {
    const BnEntry E = table[blockIdx.x];
    const int c = threadIdx.x;
    if (c >= E.ch) return;
    const double* d = defer + (size_t)E.idx * 2 * MAX_CH;
    float* rmean = bn + E.bn_off;
    float* rvar = rmean + E.ch;
    rmean[c] = bn_running_mix(rmean[c], d[c]);
    rvar[c] = bn_running_mix(rvar[c], d[MAX_CH + c]);
    if (c == 0 && nbt) nbt[E.idx] += 1;
}

int k_bn_running_update_all(const BnEntry* table_dev, int n, const double* defer, float* bn, long long* nbt, cudaStream_t s) {
    bn_running_update_all_kernel<<<n, MAX_CH, 0, s>>>(table_dev, defer, bn, nbt);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

__global__ void bn_eval_affine_kernel(int C, const float* gamma, const float* beta, const float* conv_bias,
                                      const float* rmean, const float* rvar, float* scale, float* shift) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float is = 1.0f / sqrtf(rvar[c] + (float)BN_EPS);
    float sc = gamma[c] * is;
    float bias = conv_bias ? conv_bias[c] : 0.f;
    scale[c] = sc;
    shift[c] = beta[c] + (bias - rmean[c]) * sc;
}

__global__ void bn_eval_affine_all_kernel(const float* __restrict__ params, const float* __restrict__ bn, const BnEntry* __restrict__ table,
                                          float* __restrict__ bnstat) {
    const BnEntry E = table[blockIdx.x];
    const float* gamma = params + E.gamma_off;
    const float* beta = gamma + E.ch;
    const float* rmean = bn + E.bn_off;
    const float* rvar = rmean + E.ch;
    float* scale = bnstat + ((size_t)E.idx * 4 + 2) * MAX_CH;
    float* shift = bnstat + ((size_t)E.idx * 4 + 3) * MAX_CH;
    for (int c = threadIdx.x; c < E.ch; c += blockDim.x) {
        float is = 1.0f / sqrtf(rvar[c] + (float)BN_EPS);
        float sc = gamma[c] * is;
        float bias = E.bias_off >= 0 ? params[E.bias_off + c] : 0.f;
        scale[c] = sc;
        shift[c] = beta[c] + (bias - rmean[c]) * sc;
    }
}
int k_bn_eval_affine_all(const float* params, const float* bn, const BnEntry* table_dev, int n, float* bnstat, cudaStream_t s) {
    bn_eval_affine_all_kernel<<<n, 128, 0, s>>>(params, bn, table_dev, bnstat);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

int k_bn_eval_affine(int C, const float* gamma, const float* beta, const float* conv_bias, const float* rmean,
                     const float* rvar, float* scale, float* shift, cudaStream_t s) {
    bn_eval_affine_kernel<<<ceil_div(C, 128), 128, 0, s>>>(C, gamma, beta, conv_bias, rmean, rvar, scale, shift);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// BN apply (+ residual) + ReLU -> split bf16 activation   (resnet.py:35-36, 39-45)
// ------------------------------------------------------------------------------------------
// One thread = 8 consecutive channels of one row, blocks sweep the tensor linearly.  (A persistent variant -- a thread keeps its
// channels' scale / shift in registers and walks down the rows with a grid stride -- was measured SLOWER here: 0.65 vs 0.74 of the
// copy bandwidth over the 32 launches of a step; the same restructuring does pay for bn_bwd_apply_kernel, which needs 5-11
// per-channel vectors per element.)
template <int RES_MODE>
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ raw, long long rows, int C,
                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                       Split res, const float* __restrict__ rawd,
                                                       const float* __restrict__ scaled, const float* __restrict__ shiftd,
                                                       int pitch25, Split out) {
    const int C8 = C >> 3;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * C8) return;
    long long row = idx / C8;
    int c = (int)(idx % C8) * 8;
    size_t off = (size_t)row * C + c;
    float o[8];
    if (pitch25 && !p25_valid((int)(row % IMG25))) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 0.f;
        store8_split(out, off, o);
        return;
    }
    float v[8], sc[8], sh[8];
    load8(raw + off, v); load8(scale + c, sc); load8(shift + c, sh);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaf(v[i], sc[i], sh[i]);
    if (RES_MODE == 1) {
        float r[8];
        load8_split(res, off, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += r[i];
    } else if (RES_MODE == 2) {
        float r[8], s2[8], h2[8];
        load8(rawd + off, r); load8(scaled + c, s2); load8(shiftd + c, h2);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += fmaf(r[i], s2[i], h2[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaxf(o[i], 0.f);
    store8_split(out, off, o);
}

int k_bn_apply(const float* raw, long long rows, int C, const float* scale, const float* shift, int res_mode,
               Split res, const float* rawd, const float* scaled, const float* shiftd, int pitch25, Split out,
               cudaStream_t s) {
    const long long n = rows * (C / 8);
    const int grid = grid_for(n, 256);
    if (res_mode == 0) bn_apply_kernel<0><<<grid, 256, 0, s>>>(raw, rows, C, scale, shift, res, rawd, scaled, shiftd, pitch25, out);
    else if (res_mode == 1) bn_apply_kernel<1><<<grid, 256, 0, s>>>(raw, rows, C, scale, shift, res, rawd, scaled, shiftd, pitch25, out);
    else bn_apply_kernel<2><<<grid, 256, 0, s>>>(raw, rows, C, scale, shift, res, rawd, scaled, shiftd, pitch25, out);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// stem: BN + ReLU + 3x3/2 max-pool (pad 1, -inf) : raw0 [B,48,48,64] -> a0 split pitch-25 (resnet.py:95-97)
// amax (may be NULL): per output element the window position dy*3+dx of its FIRST maximum in scan order (ATen's max_pool2d rule),
// kept by the differentiated pass so that the backward need not recompute the windows.
__global__ void __launch_bounds__(256) stem_pool_kernel(const float* __restrict__ raw0, int B,
                                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                                        Split a0, unsigned char* __restrict__ amax) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * IMG25 * 8;
    if (idx >= total) return;
    int p = (int)(idx >> 3), c = (int)(idx & 7) * 8;
    int n = p / IMG25, q = p % IMG25, y = q / PITCH, x = q % PITCH;
    float o[8];
    if (y >= HW24 || x >= HW24) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 0.f;
        store8_split(a0, (size_t)p * 64 + c, o);
        if (amax) *reinterpret_cast<uint2*>(amax + (size_t)p * 64 + c) = make_uint2(0xffffffffu, 0xffffffffu);     // halo: matches no position
        return;
    }
    float sc[8], sh[8];
    load8(scale + c, sc); load8(shift + c, sh);
    unsigned char code[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i] = -INFINITY; code[i] = 0; }
    for (int dy = 0; dy < 3; ++dy) {
        int iy = 2 * y - 1 + dy;
        if (iy < 0 || iy >= 48) continue;
        for (int dx = 0; dx < 3; ++dx) {
            int ix = 2 * x - 1 + dx;
            if (ix < 0 || ix >= 48) continue;
            float v[8];
            load8(raw0 + ((size_t)(n * 48 + iy) * 48 + ix) * 64 + c, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float a = fmaxf(fmaf(v[i], sc[i], sh[i]), 0.f);
                if (a > o[i]) { o[i] = a; code[i] = (unsigned char)(dy * 3 + dx); }      // strict: the first maximum wins
            }
        }
    }
    store8_split(a0, (size_t)p * 64 + c, o);
    if (amax) *reinterpret_cast<uint2*>(amax + (size_t)p * 64 + c) = *reinterpret_cast<const uint2*>(code);
}

int k_stem_pool(const float* raw0, int B, const float* scale, const float* shift, Split a0, unsigned char* amax, cudaStream_t s) {
    long long n = (long long)B * IMG25 * 8;
    stem_pool_kernel<<<grid_for(n, 256), 256, 0, s>>>(raw0, B, scale, shift, a0, amax);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// head: BN1 + ReLU + bilinear x2 (24->48, align_corners) : raw_h1 pitch-25 [.,128] -> u1 split dense
// One block = one output row (n, oy): the two source rows it interpolates between (24 pixels x 128 channels each, contiguous in
// the pitch-25 layout) are staged in shared memory with BN + ReLU applied once, by coalesced float4 loads; then a thread produces
// 8 channels of 3 output pixels from shared memory.  (The first version gathered 4 x 32 bytes per thread from global memory and
// re-applied BN + ReLU per tap: 0.28 of the copy bandwidth, 100 us per launch at batch 128.)
__global__ void __launch_bounds__(256) head_up1_kernel(const float* __restrict__ raw, int B,
                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                       Split u1) {
    __shared__ __align__(16) float src[2][24 * 128];
    const int n = blockIdx.x / 48, oy = blockIdx.x % 48;
    int y0, y1; float wy0, wy1;
    bilin_src(oy, 24, 48, y0, y1, wy0, wy1);
    {
        const int c4 = (threadIdx.x & 31) * 4;                    // channel group of this thread: the same for all its loads
        const float4 sc = *reinterpret_cast<const float4*>(scale + c4), sh = *reinterpret_cast<const float4*>(shift + c4);
        const float* base = raw + (size_t)n * IMG25 * 128;
#pragma unroll
        for (int k = 0; k < 6; ++k) {                             // 2 rows x 24 pixels x 32 float4 = 1536 = 6 x 256
            const int i = k * 256 + threadIdx.x;
            const int r = i / 768, px = (i % 768) >> 5;
            const float4 v = *reinterpret_cast<const float4*>(base + (size_t)((r ? y1 : y0) * PITCH + px) * 128 + c4);
            float4 o;
            o.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f); o.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
            o.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f); o.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
            *reinterpret_cast<float4*>(&src[r][px * 128 + c4]) = o;
        }
    }
    __syncthreads();
    const int c = (threadIdx.x & 15) * 8;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int ox = k * 16 + (threadIdx.x >> 4);
        int x0, x1; float wx0, wx1;
        bilin_src(ox, 24, 48, x0, x1, wx0, wx1);
        float v00[8], v01[8], v10[8], v11[8], o[8];
        load8(&src[0][x0 * 128 + c], v00); load8(&src[0][x1 * 128 + c], v01);
        load8(&src[1][x0 * 128 + c], v10); load8(&src[1][x1 * 128 + c], v11);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = wy0 * (wx0 * v00[i] + wx1 * v01[i]) + wy1 * (wx0 * v10[i] + wx1 * v11[i]);
        store8_split(u1, ((size_t)(n * 48 + oy) * 48 + ox) * 128 + c, o);
    }
}

int k_head_up1(const float* raw_h1, int B, const float* scale, const float* shift, Split u1, cudaStream_t s) {
    head_up1_kernel<<<B * 48, 256, 0, s>>>(raw_h1, B, scale, shift, u1);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// head: BN2 + ReLU + conv3 (32->A, 1x1) evaluated at 48x48.  conv3 is linear and the bilinear weights
// sum to one, so conv3(upsample(h)) = upsample(conv3_nobias(h)) + bias: the (B,32,96,96) tensor of
// networks.py:25 never materialises.
__global__ void __launch_bounds__(256) head_t_kernel(const float* __restrict__ raw, long long rows,
                                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                                     const float* __restrict__ w3, int A, float* __restrict__ t) {
    __shared__ float s_sc[32], s_sh[32], s_w[2 * 32];
    if (threadIdx.x < 32) { s_sc[threadIdx.x] = scale[threadIdx.x]; s_sh[threadIdx.x] = shift[threadIdx.x]; }
    if (threadIdx.x < A * 32) s_w[threadIdx.x] = w3[threadIdx.x];
    __syncthreads();
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= rows) return;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float v[8];
        load8(raw + (size_t)p * 32 + g * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int c = g * 8 + i;
            float h = fmaxf(fmaf(v[i], s_sc[c], s_sh[c]), 0.f);
            acc0 = fmaf(h, s_w[c], acc0);
            if (A > 1) acc1 = fmaf(h, s_w[32 + c], acc1);
        }
    }
    t[(size_t)p * A] = acc0;
    if (A > 1) t[(size_t)p * A + 1] = acc1;
}

int k_head_t(const float* raw_h2, long long rows, const float* scale, const float* shift, const float* w3, int A,
             float* t, cudaStream_t s) {
    head_t_kernel<<<grid_for(rows, 256), 256, 0, s>>>(raw_h2, rows, scale, shift, w3, A, t);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// q[n][a][oy][ox] = bilinear x2 (48->96) of t[n][.][.][a] + b3[a]          (networks.py:25-26)
__global__ void __launch_bounds__(256) head_up2_kernel(const float* __restrict__ t, int B, int A,
                                                       const float* __restrict__ b3, float* __restrict__ q) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * A * 9216;
    if (idx >= total) return;
    int ox = (int)(idx % 96), oy = (int)((idx / 96) % 96), a = (int)((idx / 9216) % A), n = (int)(idx / (9216LL * A));
    int y0, y1, x0, x1; float wy0, wy1, wx0, wx1;
    bilin_src(oy, 48, 96, y0, y1, wy0, wy1);
    bilin_src(ox, 48, 96, x0, x1, wx0, wx1);
    const float* base = t + (size_t)n * 2304 * A + a;
    float v00 = base[(size_t)(y0 * 48 + x0) * A], v01 = base[(size_t)(y0 * 48 + x1) * A];
    float v10 = base[(size_t)(y1 * 48 + x0) * A], v11 = base[(size_t)(y1 * 48 + x1) * A];
    q[idx] = wy0 * (wx0 * v00 + wx1 * v01) + wy1 * (wx0 * v10 + wx1 * v11) + b3[a];
}

int k_head_up2(const float* t, int B, int A, const float* b3, float* q, cudaStream_t s) {
    long long n = (long long)B * A * 9216;
    head_up2_kernel<<<grid_for(n, 256), 256, 0, s>>>(t, B, A, b3, q);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// DQN tail (train.py:115-129)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void argmax_combine(float& v, long long& i, float v2, long long i2) {
    if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; }      // torch.max: first maximal index
}

__device__ void block_argmax(const float* __restrict__ row, long long len, float& best_v, long long& best_i) {
    __shared__ float sv[32];
    __shared__ long long si[32];
    float v = -INFINITY; long long i = 0x7fffffffffffffffLL;
    for (long long k = threadIdx.x; k < len; k += blockDim.x) argmax_combine(v, i, row[k], k);
    for (int o = 16; o > 0; o >>= 1) {
        float v2 = __shfl_down_sync(0xffffffffu, v, o);
        long long i2 = __shfl_down_sync(0xffffffffu, i, o);
        argmax_combine(v, i, v2, i2);
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sv[w] = v; si[w] = i; }
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
        v = l < nw ? sv[l] : -INFINITY;
        i = l < nw ? si[l] : 0x7fffffffffffffffLL;
        for (int o = 16; o > 0; o >>= 1) {
            float v2 = __shfl_down_sync(0xffffffffu, v, o);
            long long i2 = __shfl_down_sync(0xffffffffu, i, o);
            argmax_combine(v, i, v2, i2);
        }
        if (l == 0) { sv[0] = v; si[0] = i; }
    }
    __syncthreads();
    best_v = sv[0]; best_i = si[0];
    __syncthreads();
}

__global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ q, long long row_len,
                                                          long long* __restrict__ idx_out) {
    float v; long long i;
    block_argmax(q + (size_t)blockIdx.x * row_len, row_len, v, i);
    if (threadIdx.x == 0) idx_out[blockIdx.x] = i;
}

int k_argmax_rows(const float* q, int B, long long row_len, long long* idx_out, cudaStream_t s) {
    argmax_rows_kernel<<<B, 256, 0, s>>>(q, row_len, idx_out);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(256) dqn_tail_kernel(const float* __restrict__ q_s, const float* __restrict__ q_no,
                                                       const float* __restrict__ q_nt, const long long* __restrict__ action,
                                                       const float* __restrict__ reward, const unsigned char* __restrict__ nonfinal,
                                                       float gamma, int B, long long row_len, int double_dqn,
                                                       float* __restrict__ per_sample, long long* __restrict__ best_action,
                                                       float* __restrict__ dq, int* __restrict__ err_flag) {
    const int i = blockIdx.x;
    float next_v = 0.f;                                   // train.py:116
    if (nonfinal[i]) {
        __shared__ int s_j;
        if (threadIdx.x == 0) {                           // row of sample i in the compacted s' batch (train.py:112)
            int j = 0;
            for (int k = 0; k < i; ++k) j += nonfinal[k] ? 1 : 0;
            s_j = j;
        }
        __syncthreads();
        const int j = s_j;
        float v; long long idx;
        if (double_dqn) {                                 // train.py:121-122
            block_argmax(q_no + (size_t)j * row_len, row_len, v, idx);
            next_v = q_nt[(size_t)j * row_len + idx];
        } else {                                          // train.py:124
            block_argmax(q_nt + (size_t)j * row_len, row_len, v, idx);
            next_v = v;
        }
        if (threadIdx.x == 0 && best_action) best_action[j] = idx;
    }
    if (threadIdx.x == 0) {
        long long a = action[i];
        if (a < 0 || a >= row_len) {                      // the reference's gather would raise: report, never index out of bounds
            if (err_flag) atomicOr(err_flag, SIMQ_DEVERR_ACTION_RANGE);
            per_sample[i * 2 + 0] = per_sample[i * 2 + 1] = __int_as_float(0x7fc00000);      // NaN loss / td_error
            return;
        }
        float q = q_s[(size_t)i * row_len + a];           // train.py:115
        float y = reward[i] + gamma * next_v;             // train.py:126
        float d = q - y, ad = fabsf(d);
        per_sample[i * 2 + 0] = ad < 1.f ? 0.5f * d * d : ad - 0.5f;      // smooth_l1, beta = 1 (train.py:129)
        per_sample[i * 2 + 1] = ad;                                        // td_error (train.py:127)
        if (dq) dq[(size_t)i * row_len + a] = fminf(fmaxf(d, -1.f), 1.f) / (float)B;
    }
}

__global__ void dqn_tail_finalize_kernel(const float* __restrict__ per_sample, int B, float* out2) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double l = 0, t = 0;
        for (int i = 0; i < B; ++i) { l += per_sample[i * 2]; t += per_sample[i * 2 + 1]; }
        out2[0] = (float)(l / B);
        out2[1] = (float)(t / B);
    }
}

int k_dqn_tail(const float* q_s, const float* q_no, const float* q_nt, const long long* action, const float* reward,
               const unsigned char* nonfinal, float gamma, int B, int Bn, int A, int double_dqn, float* per_sample,
               long long* best_action, float* out2, float* dq, int* err_flag, cudaStream_t s) {
    long long row_len = (long long)A * 9216;
    (void)Bn;
    if (dq) { SIMQ_CUDA(cudaMemsetAsync(dq, 0, sizeof(float) * (size_t)B * row_len, s)); }
    dqn_tail_kernel<<<B, 256, 0, s>>>(q_s, q_no, q_nt, action, reward, nonfinal, gamma, B, row_len, double_dqn,
                                      per_sample, best_action, dq, err_flag);
    SIMQ_LAUNCH_CHECK();
    dqn_tail_finalize_kernel<<<1, 32, 0, s>>>(per_sample, B, out2);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// BCEWithLogitsLoss (mean) + its gradient (train.py:149-150): q [n] logits, target strided
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bce_tail_kernel(const float* __restrict__ q, const float* __restrict__ target, long long tstride,
                                                       long long n, float inv_n, float* __restrict__ dq, double* __restrict__ partials) {
    __shared__ double sm[8];
    double acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float x = q[i], t = target[i * tstride];
        float e = expf(-fabsf(x));
        acc += (double)(fmaxf(x, 0.f) - x * t + log1pf(e));
        float sig = x >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
        dq[i] = (sig - t) * inv_n;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += sm[w];
        partials[blockIdx.x] = t;
    }
}
__global__ void bce_finalize_kernel(const double* __restrict__ partials, int nblk, double inv_n, float* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0;
        for (int b = 0; b < nblk; ++b) t += partials[b];
        out[0] = (float)(t * inv_n);
    }
}
int k_bce_tail(const float* q, const float* target, long long tstride, long long n, float* out1, float* dq, double* partials,
               cudaStream_t s) {
    bce_tail_kernel<<<SGD_BLOCKS, 256, 0, s>>>(q, target, tstride, n, (float)(1.0 / (double)n), dq, partials);
    SIMQ_LAUNCH_CHECK();
    bce_finalize_kernel<<<1, 32, 0, s>>>(partials, SGD_BLOCKS, 1.0 / (double)n, out1);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// clip_grad_norm_ + SGD(momentum, weight decay)  (train.py:133-135, ctor :186)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, long long n, double* __restrict__ partials) {
    __shared__ double sm[8];
    double acc = 0;
    long long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = g4[i];
        acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long i = n4 << 2; i < n; ++i) acc += (double)g[i] * g[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += sm[w];
        partials[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) sgd_update_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                         long long n, float lr, float mom, float wd, float clip_norm,
                                                         int first_step, const double* __restrict__ partials,
                                                         float* grad_norm_out) {
    __shared__ float s_coef;
    if (threadIdx.x == 0) {
        double t = 0;
        for (int b = 0; b < SGD_BLOCKS; ++b) t += partials[b];
        float norm = (float)sqrt(t);
        float coef = 1.f;
        if (clip_norm > 0.f) { coef = clip_norm / (norm + 1e-6f); if (coef > 1.f) coef = 1.f; }
        s_coef = coef;
        if (blockIdx.x == 0 && grad_norm_out) *grad_norm_out = norm;
    }
    __syncthreads();
    const float coef = s_coef;
    auto upd = [&](float& pi, float& gi, float& mi) {
        gi = gi * coef;                          // clip_grad_norm_ rescales .grad in place
        const float d = gi + wd * pi;
        const float b = first_step ? d : mom * mi + d;
        mi = b;
        pi = pi - lr * b;
    };
    const long long n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    float4* g4 = reinterpret_cast<float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pv = p4[i], gv = g4[i], mv = first_step ? make_float4(0.f, 0.f, 0.f, 0.f) : m4[i];
        upd(pv.x, gv.x, mv.x); upd(pv.y, gv.y, mv.y); upd(pv.z, gv.z, mv.z); upd(pv.w, gv.w, mv.w);
        g4[i] = gv; m4[i] = mv; p4[i] = pv;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (long long i = n4 << 2; i < n; ++i) upd(p[i], g[i], m[i]);
}

int k_sgd_step(float* params, float* grads, float* momentum, long long n, float lr, float mom, float wd,
               float clip_norm, int first_step, double* partials, float* grad_norm_out, cudaStream_t s) {
    sqnorm_kernel<<<SGD_BLOCKS, 256, 0, s>>>(grads, n, partials);
    SIMQ_LAUNCH_CHECK();
    sgd_update_kernel<<<SGD_BLOCKS, 256, 0, s>>>(params, grads, momentum, n, lr, mom, wd, clip_norm, first_step,
                                                 partials, grad_norm_out);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// weight packing OIHW fp32 -> split bf16 [tap][n][k]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, int cout, int cin, int kk,
                                                           Split fwd, Split bwd) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)kk * cout * cin;
    if (idx >= total) return;
    {   // forward pack: [t][co][ci]
        int ci = (int)(idx % cin), co = (int)((idx / cin) % cout), t = (int)(idx / ((long long)cin * cout));
        float v = w[((size_t)co * cin + ci) * kk + t];
        split_store(v, fwd.hi[idx], fwd.lo[idx]);
    }
    if (bwd.hi) {   // dgrad pack: [t][ci][co] with the taps rotated by 180 degrees
        int co = (int)(idx % cout), ci = (int)((idx / cout) % cin), t = (int)(idx / ((long long)cin * cout));
        float v = w[((size_t)co * cin + ci) * kk + (kk - 1 - t)];
        split_store(v, bwd.hi[idx], bwd.lo[idx]);
    }
}

int k_pack_weights(const float* w, int cout, int cin, int kk, Split fwd, Split bwd, cudaStream_t s) {
    long long n = (long long)kk * cout * cin;
    pack_weights_kernel<<<grid_for(n, 256), 256, 0, s>>>(w, cout, cin, kk, fwd, bwd);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

#define PACK_MAX 32
#define PACK_T 32                 // tile: 32 output channels x 32 input channels x all taps of one conv
// All conv weights of a network in ONE launch.  `start` of a table entry counts 32 x 32 tiles.  A block stages its tile of
// the OIHW tensor in shared memory with coalesced reads (for one co the 32 ci x kk taps are contiguous), then writes the
// forward shadow [t][co][ci] with ci fastest and the dgrad shadow [kk-1-t][ci][co] with co fastest: both coalesced.  (The
// element-per-thread version read with a stride of kk floats and ran at 9 % of the copy bandwidth, at the head of every step.)
__global__ void __launch_bounds__(256) pack_all_kernel(const float* __restrict__ params, const PackEntry* __restrict__ table, int n,
                                                       long long total) {
    __shared__ PackEntry tb[PACK_MAX];
    __shared__ float tile[PACK_T][PACK_T * 9 + 1];
    for (int i = threadIdx.x; i < n; i += blockDim.x) tb[i] = table[i];
    __syncthreads();
    const long long blk = blockIdx.x;
    if (blk >= total) return;
    int e = 0;
    while (e + 1 < n && blk >= tb[e + 1].start) ++e;
    const PackEntry& E = tb[e];
    const int cin = E.cin, cout = E.cout, kk = E.kk;
    const int tiles_ci = cin / PACK_T;
    const int lt = (int)(blk - E.start), co0 = (lt / tiles_ci) * PACK_T, ci0 = (lt % tiles_ci) * PACK_T;
    const float* w = params + E.w_off;
    const int row = PACK_T * kk;                                  // contiguous floats of one co within the tile
    for (int i = threadIdx.x; i < PACK_T * row; i += 256) {
        const int col = i / row, j = i - col * row;               // j = ci_l * kk + t
        tile[col][j] = w[((size_t)(co0 + col) * cin + ci0) * kk + j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PACK_T * row; i += 256) {
        const int ci_l = i % PACK_T, co_l = (i / PACK_T) % PACK_T, t = i / (PACK_T * PACK_T);
        const size_t o = ((size_t)t * cout + co0 + co_l) * cin + ci0 + ci_l;      // forward pack: [t][co][ci]
        split_store(tile[co_l][ci_l * kk + t], E.fhi[o], E.flo[o]);
    }
    const int bk = E.bk > cout ? E.bk : cout;    // bk > cout: the contraction dim of the dgrad GEMM is zero-padded (never written: the pool is zero-filled)
    for (int i = threadIdx.x; i < PACK_T * row; i += 256) {
        const int co_l = i % PACK_T, ci_l = (i / PACK_T) % PACK_T, t = i / (PACK_T * PACK_T);
        const size_t o = ((size_t)(kk - 1 - t) * cin + ci0 + ci_l) * bk + co0 + co_l;  // dgrad pack: [t'][ci][co], taps rotated by 180 degrees
        split_store(tile[co_l][ci_l * kk + t], E.bhi[o], E.blo[o]);
    }
}

int k_pack_all(const float* params, const PackEntry* table_dev, int n, long long total, cudaStream_t s) {
    if (n > PACK_MAX) { simq_set_error("k_pack_all: %d entries > %d", n, PACK_MAX); return 1; }
    pack_all_kernel<<<(unsigned)total, 256, 0, s>>>(params, table_dev, n, total);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// backward: adjoint of the 48->96 bilinear upsample, gather form (no atomics)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float bilin_adj_weight(int dst, int src, int in_size, int out_size) {
    int i0, i1; float w0, w1;
    bilin_src(dst, in_size, out_size, i0, i1, w0, w1);
    float w = 0.f;
    if (i0 == src) w += w0;
    if (i1 == src) w += w1;
    return w;
}

// candidate destination range whose footprint can touch source index s
__device__ __forceinline__ void adj_range(int s, int in_size, int out_size, int& lo, int& hi) {
    float inv = (float)(out_size - 1) / (float)(in_size - 1);
    lo = (int)floorf((float)(s - 1) * inv) - 1;
    hi = (int)ceilf((float)(s + 1) * inv) + 1;
    if (lo < 0) lo = 0;
    if (hi > out_size - 1) hi = out_size - 1;
}

__global__ void __launch_bounds__(256) up2_adj_kernel(const float* __restrict__ dq, int B, int A, float* __restrict__ dt) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * 2304 * A;
    if (idx >= total) return;
    int a = (int)(idx % A);
    long long pos = idx / A;
    int n = (int)(pos / 2304), r = (int)(pos % 2304), sy = r / 48, sx = r % 48;
    int ylo, yhi, xlo, xhi;
    adj_range(sy, 48, 96, ylo, yhi);
    adj_range(sx, 48, 96, xlo, xhi);
    const float* base = dq + ((size_t)n * A + a) * 9216;
    float acc = 0.f;
    for (int oy = ylo; oy <= yhi; ++oy) {
        float wy = bilin_adj_weight(oy, sy, 48, 96);
        if (wy == 0.f) continue;
        for (int ox = xlo; ox <= xhi; ++ox) {
            float wx = bilin_adj_weight(ox, sx, 48, 96);
            if (wx != 0.f) acc += wy * wx * base[oy * 96 + ox];
        }
    }
    dt[idx] = acc;
}

int k_up2_adj(const float* dq, int B, int A, float* dt, cudaStream_t s) {
    long long n = (long long)B * 2304 * A;
    up2_adj_kernel<<<grid_for(n, 256), 256, 0, s>>>(dq, B, A, dt);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// adjoint of the 24->48 upsample: du1 dense [B,48,48,128] f32 -> g_h1 pitch-25 [.,128] f32
__global__ void __launch_bounds__(256) up1_adj_kernel(const float* __restrict__ du1, int B, float* __restrict__ g) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * IMG25 * 16;
    if (idx >= total) return;
    int p = (int)(idx >> 4), c = (int)(idx & 15) * 8;
    int n = p / IMG25, q = p % IMG25, sy = q / PITCH, sx = q % PITCH;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (sy < HW24 && sx < HW24) {
        int ylo, yhi, xlo, xhi;
        adj_range(sy, 24, 48, ylo, yhi);
        adj_range(sx, 24, 48, xlo, xhi);
        for (int oy = ylo; oy <= yhi; ++oy) {
            float wy = bilin_adj_weight(oy, sy, 24, 48);
            if (wy == 0.f) continue;
            for (int ox = xlo; ox <= xhi; ++ox) {
                float wx = bilin_adj_weight(ox, sx, 24, 48);
                if (wx == 0.f) continue;
                float v[8];
                load8(du1 + ((size_t)(n * 48 + oy) * 48 + ox) * 128 + c, v);
                float w = wy * wx;
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, v[i], acc[i]);
            }
        }
    }
    store8(g + (size_t)p * 128 + c, acc);
}

int k_up1_adj(const float* du1, int B, float* g_h1, cudaStream_t s) {
    long long n = (long long)B * IMG25 * 16;
    up1_adj_kernel<<<grid_for(n, 256), 256, 0, s>>>(du1, B, g_h1);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// head2 backward: lane = channel (32), warp strides over positions.
// partial rows: 0 = sum dz, 1 = sum dz*xhat, 2..2+A-1 = dW3[a][c], 2+A+a (column 0) = db3[a]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head2_reduce_kernel(const float* __restrict__ dt, const float* __restrict__ raw,
                                                           long long rows, int A, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, const float* __restrict__ w3,
                                                           float* __restrict__ partials) {
    __shared__ float sm[8][6][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float sc = scale[lane], sh = shift[lane], mu = mean[lane], is = invstd[lane];
    const float w0 = w3[lane], w1 = A > 1 ? w3[32 + lane] : 0.f;
    float s1 = 0, s2 = 0, dw0 = 0, dw1 = 0, db0 = 0, db1 = 0;
    for (long long p = (long long)blockIdx.x * 8 + warp; p < rows; p += (long long)gridDim.x * 8) {
        float r = raw[(size_t)p * 32 + lane];
        float h = fmaxf(fmaf(r, sc, sh), 0.f);
        float xh = (r - mu) * is;
        float d0 = dt[(size_t)p * A], d1 = A > 1 ? dt[(size_t)p * A + 1] : 0.f;
        float dz = h > 0.f ? (w0 * d0 + w1 * d1) : 0.f;
        s1 += dz; s2 += dz * xh; dw0 += d0 * h; dw1 += d1 * h; db0 += d0; db1 += d1;
    }
    sm[warp][0][lane] = s1; sm[warp][1][lane] = s2; sm[warp][2][lane] = dw0; sm[warp][3][lane] = dw1;
    sm[warp][4][lane] = db0; sm[warp][5][lane] = db1;
    __syncthreads();
    const int K = 2 + 2 * A;
    if (threadIdx.x < 32) {
        for (int k = 0; k < K; ++k) {
            int src;                        // map output row k to the fixed smem slot
            if (k < 2) src = k; else if (k < 2 + A) src = 2 + (k - 2); else src = 4 + (k - 2 - A);
            float t = 0;
            for (int w = 0; w < 8; ++w) t += sm[w][src][lane];
            partials[((size_t)blockIdx.x * K + k) * 32 + lane] = t;
        }
    }
}

int k_head2_reduce(const float* dt, const float* raw_h2, long long rows, int A, const float* scale,
                   const float* shift, const float* mean, const float* invstd, const float* w3, float* partials,
                   cudaStream_t s) {
    head2_reduce_kernel<<<STAT_BLOCKS, 256, 0, s>>>(dt, raw_h2, rows, A, scale, shift, mean, invstd, w3, partials);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int nblk, int K,
                                                              float* __restrict__ out, float mul) {
    __shared__ double sm[8][32];
    const int kl = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + kl;
    double t = 0;
    if (k < K)
        for (int b = r; b < nblk; b += 8) t += (double)partials[(size_t)b * K + k];
    sm[r][kl] = t;
    __syncthreads();
    if (r != 0 || k >= K) return;
    t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][kl];
    out[k] = (float)(t * (double)mul);
}

int k_reduce_partials(const float* partials, int nblk, int K, float* out, float mul, cudaStream_t s) {
    reduce_partials_kernel<<<ceil_div(K, 32), 256, 0, s>>>(partials, nblk, K, out, mul);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(256) head2_apply_kernel(const float* __restrict__ dt, const float* __restrict__ raw,
                                                          long long rows, int A, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, const float* __restrict__ mean,
                                                          const float* __restrict__ invstd, const float* __restrict__ w3,
                                                          const float* __restrict__ sums, float inv_count, Split dy) {
    const int lane = threadIdx.x & 31;
    long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= rows) return;
    const float sc = scale[lane], sh = shift[lane], mu = mean[lane], is = invstd[lane];
    const float c1 = sums[lane] * inv_count, c2 = sums[32 + lane] * inv_count;
    float r = raw[(size_t)p * 32 + lane];
    float h = fmaf(r, sc, sh);
    float xh = (r - mu) * is;
    float d0 = dt[(size_t)p * A], d1 = A > 1 ? dt[(size_t)p * A + 1] : 0.f;
    float dz = h > 0.f ? (w3[lane] * d0 + (A > 1 ? w3[32 + lane] * d1 : 0.f)) : 0.f;
    float v = sc * (dz - c1 - xh * c2);
    // dy rows are HEAD2_DY_STRIDE = 64 wide, channels 32..63 zero: the 32-channel conv2 then runs its dgrad / wgrad on the
    // tensor-core tiles (K resp. M of 64) instead of the FMA comparator kernels
    const size_t o = (size_t)p * HEAD2_DY_STRIDE + lane;
    split_store(v, dy.hi[o], dy.lo[o]);
    dy.hi[o + 32] = __float2bfloat16_rn(0.f); dy.lo[o + 32] = __float2bfloat16_rn(0.f);
}

int k_head2_apply(const float* dt, const float* raw_h2, long long rows, int A, const float* scale, const float* shift,
                  const float* mean, const float* invstd, const float* w3, const float* sums, double count,
                  Split dy, cudaStream_t s) {
    long long n = rows * 32;
    head2_apply_kernel<<<grid_for(n, 256), 256, 0, s>>>(dt, raw_h2, rows, A, scale, shift, mean, invstd, w3, sums,
                                                        (float)(1.0 / count), dy);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// generic BatchNorm backward over fp32 [rows][C]:  dz = G * mask
//   s1 = sum dz, s2 = sum dz*xhat (, s2d = sum dz*xhat_d for the downsample branch)
//   dy = gamma*invstd * (dz - s1/N - xhat*s2/N)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ G, long long rows, int C,
                                                            int mask_mode, const bf16* __restrict__ mask_hi,
                                                            const float* __restrict__ raw, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd, const float* __restrict__ rawd,
                                                            const float* __restrict__ meand, const float* __restrict__ invstdd,
                                                            float* __restrict__ partials) {
    __shared__ float sm[3][1024];
    const int C4 = C >> 2, rpi = 256 / C4;
    const int c4 = threadIdx.x % C4, rl = threadIdx.x / C4, c = c4 * 4;
    float mu[4], is[4], sc[4], sh[4], mud[4], isd[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        mu[i] = mean[c + i]; is[i] = invstd[c + i];
        sc[i] = mask_mode == 2 ? scale[c + i] : 0.f; sh[i] = mask_mode == 2 ? shift[c + i] : 0.f;
        mud[i] = rawd ? meand[c + i] : 0.f; isd[i] = rawd ? invstdd[c + i] : 0.f;
    }
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0}, s3[4] = {0, 0, 0, 0};
    // two rows per trip, all their loads issued before the first use (one row per trip keeps too few bytes in flight for
    // 296 x 256 threads: the kernel was latency-bound at ~60 % of the copy bandwidth); summation order unchanged
    const long long stride = (long long)gridDim.x * rpi;
    for (long long r = (long long)blockIdx.x * rpi + rl; r < rows; r += 2 * stride) {
        const long long rr[2] = {r, r + stride};
        float4 g4[2], r4[2], d4[2];
        uint2 mh[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const bool live = rr[u] < rows;
            const size_t off = (size_t)(live ? rr[u] : r) * C + c;
            g4[u] = *reinterpret_cast<const float4*>(G + off);
            r4[u] = *reinterpret_cast<const float4*>(raw + off);
            mh[u] = mask_mode == 1 ? *reinterpret_cast<const uint2*>(mask_hi + off) : make_uint2(0u, 0u);
            d4[u] = rawd ? *reinterpret_cast<const float4*>(rawd + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (rr[u] >= rows) continue;
            float g[4] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w}, rv[4] = {r4[u].x, r4[u].y, r4[u].z, r4[u].w};
            float rd[4] = {d4[u].x, d4[u].y, d4[u].z, d4[u].w};
            bool m[4] = {true, true, true, true};
            if (mask_mode == 1) {
                const bf16* mb = reinterpret_cast<const bf16*>(&mh[u]);
#pragma unroll
                for (int i = 0; i < 4; ++i) m[i] = bf2f(mb[i]) > 0.f;
            } else if (mask_mode == 2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) m[i] = fmaf(rv[i], sc[i], sh[i]) > 0.f;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float dz = m[i] ? g[i] : 0.f;
                s1[i] += dz;
                s2[i] += dz * (rv[i] - mu[i]) * is[i];
                s3[i] += dz * (rd[i] - mud[i]) * isd[i];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sm[0][rl * C + c + i] = s1[i]; sm[1][rl * C + c + i] = s2[i]; sm[2][rl * C + c + i] = s3[i];
    }
    __syncthreads();
    for (int cc = threadIdx.x; cc < C; cc += 256) {
        float t0 = 0, t1 = 0, t2 = 0;
        for (int r = 0; r < rpi; ++r) { t0 += sm[0][r * C + cc]; t1 += sm[1][r * C + cc]; t2 += sm[2][r * C + cc]; }
        partials[((size_t)blockIdx.x * 3 + 0) * C + cc] = t0;
        partials[((size_t)blockIdx.x * 3 + 1) * C + cc] = t1;
        partials[((size_t)blockIdx.x * 3 + 2) * C + cc] = t2;
    }
}

int k_bn_bwd_reduce(const float* G, long long rows, int C, int mask_mode, const bf16* mask_hi, const float* raw,
                    const float* scale, const float* shift, const float* mean, const float* invstd,
                    const float* rawd, const float* meand, const float* invstdd, float* partials, cudaStream_t s) {
    bn_bwd_reduce_kernel<<<STAT_BLOCKS, 256, 0, s>>>(G, rows, C, mask_mode, mask_hi, raw, scale, shift, mean, invstd,
                                                     rawd, meand, invstdd, partials);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// Same structure as bn_apply_kernel: a thread owns 8 channels (mean / invstd / the three per-channel coefficients of
//   dy = gamma*invstd * (dz - s1/N - xhat*s2/N)
// stay in registers), rows are walked with a grid stride, two per trip, every streaming load (G, raw, mask, rawd) issued before
// the first use.  The arithmetic per element is unchanged (same operations in the same order as the one-row-per-thread version).
template <int MASK_MODE, bool HAS_DS>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ G, long long rows, int C,
                                                           const bf16* __restrict__ mask_hi, const float* __restrict__ raw,
                                                           const float* __restrict__ scale, const float* __restrict__ shift,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ sums,
                                                           float inv_count, int pitch25, Split dy, float* __restrict__ dy_f32,
                                                           const float* __restrict__ rawd, const float* __restrict__ meand,
                                                           const float* __restrict__ invstdd, const float* __restrict__ gammad,
                                                           Split dyd, float* dgamma, float* dbeta, float* dgammad, float* dbetad, int hi_only) {
    const int C8 = C >> 3, rpi = 256 / C8;
    const int c = (threadIdx.x % C8) * 8, rl = threadIdx.x / C8;
    float s1[8], s2[8], mu[8], is[8], ga[8], sc[8], sh[8], s3[8], mud[8], isd[8], gad[8];
    load8(sums + c, s1); load8(sums + C + c, s2);
    load8(mean + c, mu); load8(invstd + c, is); load8(gamma + c, ga);
    if (MASK_MODE == 2) { load8(scale + c, sc); load8(shift + c, sh); }
    if (HAS_DS) { load8(sums + 2 * C + c, s3); load8(meand + c, mud); load8(invstdd + c, isd); load8(gammad + c, gad); }
    if (blockIdx.x == 0 && rl == 0) {                 // parameter gradients of the affine BN
        store8(dgamma + c, s2); store8(dbeta + c, s1);
        if (HAS_DS) { store8(dgammad + c, s3); store8(dbetad + c, s1); }
    }
    const long long stride = (long long)gridDim.x * rpi;
    for (long long r0 = (long long)blockIdx.x * rpi + rl; r0 < rows; r0 += 2 * stride) {
        const long long rr[2] = {r0, r0 + stride};
        float g[2][8], rv[2][8], rd[2][8];
        bf16x8 mh[2];
        bool live[2], valid[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            live[u] = rr[u] < rows;
            valid[u] = live[u] && !(pitch25 && !p25_valid((int)(rr[u] % IMG25)));
            const size_t off = (size_t)(live[u] ? rr[u] : r0) * C + c;
            if (valid[u]) {
                load8(G + off, g[u]); load8(raw + off, rv[u]);
                if (MASK_MODE == 1) mh[u] = *reinterpret_cast<const bf16x8*>(mask_hi + off);
                if (HAS_DS) load8(rawd + off, rd[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!live[u]) continue;
            const size_t off = (size_t)rr[u] * C + c;
            float o[8], od[8];
            if (!valid[u]) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { o[i] = 0.f; od[i] = 0.f; }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    bool m = true;
                    if (MASK_MODE == 1) m = bf2f(mh[u].v[i]) > 0.f;
                    if (MASK_MODE == 2) m = fmaf(rv[u][i], sc[i], sh[i]) > 0.f;
                    const float dz = m ? g[u][i] : 0.f;
                    const float xh = (rv[u][i] - mu[i]) * is[i];
                    o[i] = ga[i] * is[i] * (dz - s1[i] * inv_count - xh * s2[i] * inv_count);
                    if (HAS_DS) {
                        const float xd = (rd[u][i] - mud[i]) * isd[i];
                        od[i] = gad[i] * isd[i] * (dz - s1[i] * inv_count - xd * s3[i] * inv_count);
                    }
                }
            }
            if (dy_f32) store8(dy_f32 + off, o); else if (hi_only) store8_hi(dy, off, o); else store8_split(dy, off, o);
            if (HAS_DS) { if (hi_only) store8_hi(dyd, off, od); else store8_split(dyd, off, od); }
        }
    }
}

int k_bn_bwd_apply(const float* G, long long rows, int C, int mask_mode, const bf16* mask_hi, const float* raw,
                   const float* scale, const float* shift, const float* mean, const float* invstd, const float* gamma,
                   const float* sums, double count, int pitch25, Split dy, float* dy_f32, const float* rawd,
                   const float* meand, const float* invstdd, const float* gammad, Split dyd, float* dgamma,
                   float* dbeta, float* dgammad, float* dbetad, int hi_only, cudaStream_t s) {
    if (C % 8 != 0 || C > 2048 || 256 % (C / 8) != 0) { simq_set_error("k_bn_bwd_apply: C=%d", C); return 1; }
    const int rpi = 256 / (C / 8);
    long long want = (rows + 2 * rpi - 1) / (2 * rpi);
    const int grid = (int)(want < 1 ? 1 : want > 8 * 148 ? 8 * 148 : want);
    const float inv = (float)(1.0 / count);
#define BWD_APPLY(MM, DS) bn_bwd_apply_kernel<MM, DS><<<grid, 256, 0, s>>>(G, rows, C, mask_hi, raw, scale, shift, mean, invstd, gamma, sums, inv, \
        pitch25, dy, dy_f32, rawd, meand, invstdd, gammad, dyd, dgamma, dbeta, dgammad, dbetad, hi_only)
    if (rawd) {
        if (mask_mode == 1) BWD_APPLY(1, true); else if (mask_mode == 2) BWD_APPLY(2, true); else BWD_APPLY(0, true);
    } else {
        if (mask_mode == 1) BWD_APPLY(1, false); else if (mask_mode == 2) BWD_APPLY(2, false); else BWD_APPLY(0, false);
    }
#undef BWD_APPLY
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// stem: backward of 3x3/2 max-pool + ReLU: g_a0 pitch-25 [.,64] f32 -> dz0 dense [B,48,48,64] f32.
// Gather form (no atomics): an input pixel receives the gradient of every pooling window whose FIRST maximum (window scan order, as
// ATen's max_pool2d) it is -- the forward pass stored that position per output element (stem_pool_kernel's amax) -- and the ReLU
// gradient is zero where the activation is zero, so ties among zeros cannot matter.  Per thread: its own raw value, and for each
// of the <= 4 windows covering it 8 position bytes + 8 gradients (the first version recomputed every window: 36 loads per thread).
__global__ void __launch_bounds__(256) pool_bwd_kernel(const float* __restrict__ g_a0, const float* __restrict__ raw0,
                                                       const unsigned char* __restrict__ amax, int B,
                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                       float* __restrict__ dz0) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * 2304 * 8;
    if (idx >= total) return;
    long long pos = idx >> 3;
    int c = (int)(idx & 7) * 8;
    int n = (int)(pos / 2304), r = (int)(pos % 2304), y = r / 48, x = r % 48;
    float sc[8], sh[8], self[8], acc[8];
    load8(scale + c, sc); load8(shift + c, sh);
    load8(raw0 + (size_t)pos * 64 + c, self);
#pragma unroll
    for (int i = 0; i < 8; ++i) { self[i] = fmaxf(fmaf(self[i], sc[i], sh[i]), 0.f); acc[i] = 0.f; }
    int oy_lo = (y >= 1) ? (y - 1 + 1) / 2 : 0;      // ceil((y-1)/2)
    int oy_hi = (y + 1) / 2; if (oy_hi > 23) oy_hi = 23;
    int ox_lo = (x >= 1) ? (x - 1 + 1) / 2 : 0;
    int ox_hi = (x + 1) / 2; if (ox_hi > 23) ox_hi = 23;
    for (int oy = oy_lo; oy <= oy_hi; ++oy)
        for (int ox = ox_lo; ox <= ox_hi; ++ox) {
            const size_t o = ((size_t)n * IMG25 + oy * PITCH + ox) * 64 + c;
            const uint2 cw = *reinterpret_cast<const uint2*>(amax + o);
            const unsigned char* code = reinterpret_cast<const unsigned char*>(&cw);
            const unsigned char mine = (unsigned char)((y - (2 * oy - 1)) * 3 + (x - (2 * ox - 1)));      // this pixel's position in that window
            float g[8];
            load8(g_a0 + o, g);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (code[i] == mine && self[i] > 0.f) acc[i] += g[i];
        }
    store8(dz0 + (size_t)pos * 64 + c, acc);
}

int k_pool_bwd(const float* g_a0, const float* raw0, const unsigned char* amax, int B, const float* scale, const float* shift, float* dz0,
               cudaStream_t s) {
    long long n = (long long)B * 2304 * 8;
    pool_bwd_kernel<<<grid_for(n, 256), 256, 0, s>>>(g_a0, raw0, amax, B, scale, shift, dz0);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// per-channel column sums of a split tensor (conv bias gradients): partials [STAT_BLOCKS][C]
__global__ void __launch_bounds__(256) colsum_split_kernel(Split dy, long long rows, int C, float* __restrict__ partials) {
    __shared__ float sm[2048];
    const int C8 = C >> 3, rpi = 256 / C8;
    const int c8 = threadIdx.x % C8, rl = threadIdx.x / C8;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (rl < rpi)
        for (long long r = (long long)blockIdx.x * rpi + rl; r < rows; r += (long long)gridDim.x * rpi) {
            float v[8];
            load8_split(dy, (size_t)r * C + c8 * 8, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += v[i];
        }
    if (rl < rpi) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sm[rl * C + c8 * 8 + i] = acc[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float t = 0;
        for (int r = 0; r < rpi; ++r) t += sm[r * C + c];
        partials[(size_t)blockIdx.x * C + c] = t;
    }
}

int k_colsum_split(Split dy, long long rows, int C, float* partials, cudaStream_t s) {
    if (C % 8 != 0 || C > 2048 || 256 % (C / 8) != 0) { simq_set_error("k_colsum_split: C=%d", C); return 1; }
    colsum_split_kernel<<<STAT_BLOCKS, 256, 0, s>>>(dy, rows, C, partials);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

__global__ void head2_scatter_kernel(const float* __restrict__ sums, int A, float* dgamma2, float* dbeta2, float* dW3, float* db3) {
    int i = threadIdx.x;
    if (i < 32) { dbeta2[i] = sums[i]; dgamma2[i] = sums[32 + i]; }
    if (i < A * 32) dW3[i] = sums[64 + i];
    if (i < A) db3[i] = sums[(2 + A + i) * 32];
}

int k_head2_scatter(const float* sums, int A, float* dgamma2, float* dbeta2, float* dW3, float* db3, cudaStream_t s) {
    head2_scatter_kernel<<<1, 64, 0, s>>>(sums, A, dgamma2, dbeta2, dW3, db3);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// dst[j][:] = src[idx[j]][:]  (rows of row_floats fp32, multiple of 4): replay-batch assembly on the device
__global__ void __launch_bounds__(256) gather_rows_kernel(const float4* __restrict__ src, const long long* __restrict__ idx,
                                                          long long row_vec, float4* __restrict__ dst) {
    const long long j = blockIdx.y;
    const float4* s = src + idx[j] * row_vec;
    float4* d = dst + j * row_vec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < row_vec; i += (long long)gridDim.x * blockDim.x) d[i] = s[i];
}
int k_gather_rows(const float* src, const long long* idx, int n, long long row_floats, float* dst, cudaStream_t s) {
    if (n <= 0) return 0;
    if (row_floats % 4) { simq_set_error("k_gather_rows: row length %lld not a multiple of 4", row_floats); return 1; }
    dim3 grid(16, n);
    gather_rows_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(src), idx, row_floats / 4, reinterpret_cast<float4*>(dst));
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// debug / test import-export between dense NCHW f32 and the internal layouts
// ------------------------------------------------------------------------------------------
__global__ void export_p25_kernel(const float* raw, Split sp, int B, int C, float* out) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * C * 576;
    if (idx >= total) return;
    int x = (int)(idx % 24), y = (int)((idx / 24) % 24), c = (int)((idx / 576) % C), n = (int)(idx / (576LL * C));
    size_t off = ((size_t)n * IMG25 + y * PITCH + x) * C + c;
    out[idx] = raw ? raw[off] : bf2f(sp.hi[off]) + bf2f(sp.lo[off]);
}
int k_export_p25(const float* raw, Split sp, int B, int C, float* out, cudaStream_t s) {
    long long n = (long long)B * C * 576;
    export_p25_kernel<<<grid_for(n, 256), 256, 0, s>>>(raw, sp, B, C, out);
    SIMQ_LAUNCH_CHECK();
    return 0;
}
__global__ void export_dense_kernel(const float* raw, Split sp, int B, int C, int HW, float* out) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * C * HW * HW;
    if (idx >= total) return;
    int x = (int)(idx % HW), y = (int)((idx / HW) % HW), c = (int)((idx / ((long long)HW * HW)) % C);
    int n = (int)(idx / ((long long)HW * HW * C));
    size_t off = (((size_t)n * HW + y) * HW + x) * C + c;
    out[idx] = raw ? raw[off] : bf2f(sp.hi[off]) + bf2f(sp.lo[off]);
}
int k_export_dense(const float* raw, Split sp, int B, int C, int HW, float* out, cudaStream_t s) {
    long long n = (long long)B * C * HW * HW;
    export_dense_kernel<<<grid_for(n, 256), 256, 0, s>>>(raw, sp, B, C, HW, out);
    SIMQ_LAUNCH_CHECK();
    return 0;
}
__global__ void import_p25_kernel(const float* nchw, int B, int C, Split sp, float* raw) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * IMG25 * C;
    if (idx >= total) return;
    int c = (int)(idx % C);
    int p = (int)(idx / C);
    int n = p / IMG25, q = p % IMG25, y = q / PITCH, x = q % PITCH;
    float v = (y < HW24 && x < HW24) ? nchw[(((size_t)n * C + c) * 24 + y) * 24 + x] : 0.f;
    if (sp.hi) split_store(v, sp.hi[idx], sp.lo[idx]);
    if (raw) raw[idx] = v;
}
int k_import_p25(const float* nchw, int B, int C, Split sp, float* raw, cudaStream_t s) {
    long long n = (long long)B * IMG25 * C;
    import_p25_kernel<<<grid_for(n, 256), 256, 0, s>>>(nchw, B, C, sp, raw);
    SIMQ_LAUNCH_CHECK();
    return 0;
}
