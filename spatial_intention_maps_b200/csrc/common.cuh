// Shared helpers for the simq kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// "pitch-25" layout of the 24x24 feature maps: position p = n*625 + y*25 + x with y,x in [0,25);
// row 24 and column 24 are always ZERO.  Column 24 of row y is at once the right halo of row y and
// the left halo of row y+1, row 24 is the bottom halo of image n and the top halo of image n+1, so
// every 3x3 tap is the constant flat offset dy*25+dx and a conv is 9 shifted GEMMs over B*625 rows
// (8.5 % padding instead of 17 % for a 26x26 frame).  Reads before row 0 / after the last row are
// zero-filled (TMA out-of-bounds fill; bounds check in the FMA kernels).
// ---------------------------------------------------------------------------------------------
#define PITCH 25
#define IMG25 625
#define HW24 24

__host__ __device__ inline bool p25_valid(int p) {
    int q = p % IMG25;
    return (q < HW24 * PITCH) && (q % PITCH) < HW24;
}

struct Split {          // a tensor stored as two bf16 planes: value = hi + lo (16 mantissa bits)
    bf16* hi;
    bf16* lo;
};

__device__ __forceinline__ void split_store(float v, bf16& hi, bf16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__device__ __forceinline__ float bf2f(bf16 v) { return __bfloat162float(v); }

// 8 consecutive channels <-> two 16-byte vectors
struct __align__(16) bf16x8 { bf16 v[8]; };

__device__ __forceinline__ void load8(const float* p, float* o) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float* o) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
}
__device__ __forceinline__ void load8_split(const Split& s, size_t off, float* o) {
    bf16x8 h = *reinterpret_cast<const bf16x8*>(s.hi + off);
    bf16x8 l = *reinterpret_cast<const bf16x8*>(s.lo + off);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = bf2f(h.v[i]) + bf2f(l.v[i]);
}
__device__ __forceinline__ void store8_split(const Split& s, size_t off, const float* o) {
    bf16x8 h, l;
#pragma unroll
    for (int i = 0; i < 8; ++i) split_store(o[i], h.v[i], l.v[i]);
    *reinterpret_cast<bf16x8*>(s.hi + off) = h;
    *reinterpret_cast<bf16x8*>(s.lo + off) = l;
}

__device__ __forceinline__ void store8_hi(const Split& s, size_t off, const float* o) {      // hi plane only (two-term consumers)
    bf16x8 h;
#pragma unroll
    for (int i = 0; i < 8; ++i) h.v[i] = __float2bfloat16_rn(o[i]);
    *reinterpret_cast<bf16x8*>(s.hi + off) = h;
}

// bilinear x2, align_corners=True source index exactly as ATen's area_pixel_compute_source_index +
// upsample_bilinear2d (fp32): src = dst * (in-1)/(out-1); i0 = (int)src; i1 = i0 + (i0 < in-1);
// w1 = src - i0; w0 = 1 - w1.    (networks.py:21,25)
__host__ __device__ inline void bilin_src(int dst, int in_size, int out_size, int& i0, int& i1, float& w0, float& w1) {
    float scale = (float)(in_size - 1) / (float)(out_size - 1);
    float src = scale * (float)dst;
    i0 = (int)src;
    i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    w1 = src - (float)i0;
    w0 = 1.0f - w1;
}

// ---- nn.BatchNorm2d train-mode bookkeeping shared by bn_finalize_train_kernel and the conv epilogues' fused finalize ----
#define SIMQ_BN_EPS 1e-5
#define SIMQ_BN_MOM 0.1
// running = (1 - momentum) * running + momentum * batch_value, with the rounding of every operation pinned (no
// compiler-chosen FMA contraction): the immediate and the deferred update must agree bit for bit.
__device__ __forceinline__ float bn_running_mix(float running, double batch_value) {
    return (float)__dadd_rn(__dmul_rn(1.0 - SIMQ_BN_MOM, (double)running), __dmul_rn(SIMQ_BN_MOM, batch_value));
}
// One channel from its column sum S and sum of squares SS over `count` positions (eps 1e-5, momentum 0.1, unbiased variance into
// running_var).  `raw` excludes the conv bias: the batch mean of the true conv output is mean_raw + bias, and the bias cancels in
// the normalised value.  defer != NULL: a concurrent pass owns the running statistics, stash (mean + bias, unbiased var) instead.
__device__ __forceinline__ void bn_finalize_channel(double S, double SS, double count, int c, const float* gamma, const float* beta,
                                                    const float* conv_bias, float* rmean, float* rvar, double* defer, int defer_stride,
                                                    float* mean, float* invstd, float* scale, float* shift) {
    double m = S / count;
    double var = SS / count - m * m;
    if (var < 0) var = 0;
    float is = (float)(1.0 / sqrt(var + SIMQ_BN_EPS));
    float g = gamma[c], b = beta[c];
    float bias = conv_bias ? conv_bias[c] : 0.f;
    mean[c] = (float)m;
    invstd[c] = is;
    float sc = g * is;
    scale[c] = sc;
    shift[c] = b - (float)m * sc;
    double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    if (defer) {
        defer[c] = m + (double)bias;
        defer[defer_stride + c] = unbiased;
        return;
    }
    rmean[c] = bn_running_mix(rmean[c], m + (double)bias);
    rvar[c] = bn_running_mix(rvar[c], unbiased);
}

extern thread_local long long g_simq_launches;     // kernels launched by this library on the calling thread (host counter;
                                                   // api.cu credits the difference over an entry point to that call's context)
void simq_set_error(const char* fmt, ...);

#define SIMQ_CUDA(expr)                                                                          \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            simq_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

#define SIMQ_LAUNCH_CHECK()                                                                      \
    do {                                                                                         \
        ++g_simq_launches;                                                                       \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess) {                                                                 \
            simq_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e));    \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

// true the first time a call site runs on the current device (bit mask per device)
static inline bool first_use_on_device(unsigned long long& mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    if ((mask >> dev) & 1ull) return false;
    mask |= 1ull << dev;
    return true;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
