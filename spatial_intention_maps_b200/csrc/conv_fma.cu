// CUDA-core (fp32 FMA) convolutions.
//  * k_conv_fma / k_wgrad_fma: shifted-GEMM conv on the split-bf16 operands, fp32 accumulate.  This is
//    the on-device COMPARATOR for the tcgen05 kernels (identical inputs, independent arithmetic) and is
//    selectable as a debug back-end; it is not the product path for the 24x24 stages.
//  * k_stem_conv / k_stem_wgrad: the 7x7/2 stem (C <= 10 input channels, K = 49*C): too thin a
//    contraction for the tensor cores, HBM/FMA-bound, so FMA is the product path.
#include "kernels.h"

__device__ __forceinline__ int tap_offset(int t, int ntaps) {
    return ntaps == 9 ? (t / 3 - 1) * PITCH + (t % 3 - 1) : 0;
}

// ------------------------------------------------------------------------------------------
// out[m][n] = sum_t sum_k A[m + off_t][k] * W[t][n][k]
// ------------------------------------------------------------------------------------------
#define FB 64
#define FK 16
__global__ void __launch_bounds__(256) conv_fma_kernel(Split A, long long rows, int K, Split W, int N, int ntaps,
                                                       float* __restrict__ out, ConvEpilogue ep) {
    __shared__ float As[FK][FB + 4];
    __shared__ float Bs[FK][FB + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * FB;
    const int n0 = blockIdx.y * FB;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int t = 0; t < ntaps; ++t) {
        const long long arow = m0 + lrow + tap_offset(t, ntaps);
        const bool a_ok = arow >= 0 && arow < rows;
        const int wrow = n0 + lrow;
        const bool w_ok = wrow < N;
        for (int k0 = 0; k0 < K; k0 += FK) {
            float av[4] = {0, 0, 0, 0}, wv[4] = {0, 0, 0, 0};
            if (a_ok) {
                size_t off = (size_t)arow * K + k0 + lk;
                uint2 h = *reinterpret_cast<const uint2*>(A.hi + off);
                uint2 l = *reinterpret_cast<const uint2*>(A.lo + off);
                const bf16* hb = reinterpret_cast<const bf16*>(&h);
                const bf16* lb = reinterpret_cast<const bf16*>(&l);
#pragma unroll
                for (int i = 0; i < 4; ++i) av[i] = bf2f(hb[i]) + bf2f(lb[i]);
            }
            if (w_ok) {
                size_t off = ((size_t)t * N + wrow) * K + k0 + lk;
                uint2 h = *reinterpret_cast<const uint2*>(W.hi + off);
                uint2 l = *reinterpret_cast<const uint2*>(W.lo + off);
                const bf16* hb = reinterpret_cast<const bf16*>(&h);
                const bf16* lb = reinterpret_cast<const bf16*>(&l);
#pragma unroll
                for (int i = 0; i < 4; ++i) wv[i] = bf2f(hb[i]) + bf2f(lb[i]);
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i) { As[lk + i][lrow] = av[i]; Bs[lk + i][lrow] = wv[i]; }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < FK; ++kk) {
                float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long m = m0 + ty * 4 + i;
        if (m >= rows) continue;
        bool valid = !(ep.pitch25 && !p25_valid((int)(m % IMG25)));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            size_t o = (size_t)m * N + n;
            float v = 0.f;
            if (valid) {
                v = acc[i][j];
                if (ep.add_prev) v += ep.add_prev[o];
                if (ep.add_g && bf2f(ep.add_g_mask[o]) > 0.f) v += ep.add_g[o];
            }
            out[o] = v;
        }
    }
}

int k_conv_fma(Split A, long long rows, int K, Split W, int N, int ntaps, float* out, ConvEpilogue ep, cudaStream_t s) {
    if (K % FK != 0) { simq_set_error("k_conv_fma: K=%d not a multiple of %d", K, FK); return 1; }
    dim3 grid(ceil_div(rows, FB), ceil_div(N, FB));
    conv_fma_kernel<<<grid, 256, 0, s>>>(A, rows, K, W, N, ntaps, out, ep);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// dW[co][ci][t] = sum_p dY[p][co] * X[p + off_t][ci]       (OIHW output, atomics across row splits)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wgrad_fma_kernel(Split dY, Split X, long long rows, int Cout, int Cin, int ntaps,
                                                        int nsplit, float* __restrict__ partial) {
    __shared__ float Ys[FK][FB + 4];
    __shared__ float Xs[FK][FB + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int co0 = blockIdx.x * FB, ci0 = blockIdx.y * FB;
    const int t = blockIdx.z % ntaps, sp = blockIdx.z / ntaps;
    const int off = tap_offset(t, ntaps);
    long long chunk = ((rows + nsplit - 1) / nsplit + FK - 1) / FK * FK;
    long long p_begin = (long long)sp * chunk, p_end = p_begin + chunk;
    if (p_end > rows) p_end = rows;
    const int lp = tid >> 4, lc = (tid & 15) * 4;       // 16 rows x 64 channels per tile
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (long long p0 = p_begin; p0 < p_end; p0 += FK) {
        float yv[4] = {0, 0, 0, 0}, xv[4] = {0, 0, 0, 0};
        long long py = p0 + lp, px = py + off;
        if (py < p_end && co0 + lc < Cout) {
            size_t o = (size_t)py * Cout + co0 + lc;
            uint2 h = *reinterpret_cast<const uint2*>(dY.hi + o);
            uint2 l = *reinterpret_cast<const uint2*>(dY.lo + o);
            const bf16* hb = reinterpret_cast<const bf16*>(&h);
            const bf16* lb = reinterpret_cast<const bf16*>(&l);
#pragma unroll
            for (int i = 0; i < 4; ++i) yv[i] = bf2f(hb[i]) + bf2f(lb[i]);
        }
        if (py < p_end && px >= 0 && px < rows && ci0 + lc < Cin) {
            size_t o = (size_t)px * Cin + ci0 + lc;
            uint2 h = *reinterpret_cast<const uint2*>(X.hi + o);
            uint2 l = *reinterpret_cast<const uint2*>(X.lo + o);
            const bf16* hb = reinterpret_cast<const bf16*>(&h);
            const bf16* lb = reinterpret_cast<const bf16*>(&l);
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = bf2f(hb[i]) + bf2f(lb[i]);
        }
        __syncthreads();
        *reinterpret_cast<float4*>(&Ys[lp][lc]) = make_float4(yv[0], yv[1], yv[2], yv[3]);
        *reinterpret_cast<float4*>(&Xs[lp][lc]) = make_float4(xv[0], xv[1], xv[2], xv[3]);
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < FK; ++kk) {
            float4 a4 = *reinterpret_cast<const float4*>(&Ys[kk][ty * 4]);
            float4 b4 = *reinterpret_cast<const float4*>(&Xs[kk][tx * 4]);
            float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int co = co0 + ty * 4 + i;
        if (co >= Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int ci = ci0 + tx * 4 + j;
            if (ci >= Cin) continue;
            partial[((size_t)blockIdx.z * Cout + co) * Cin + ci] = acc[i][j];     // z = sp*ntaps + t
        }
    }
}

int k_wgrad_fma(Split dY, Split X, long long rows, int Cout, int Cin, int ntaps, float* dW, float* scratch, cudaStream_t s) {
    int base = ceil_div(Cout, FB) * ceil_div(Cin, FB) * ntaps;
    int nsplit = (1184 + base - 1) / base;
    long long max_split = rows / 512; if (max_split < 1) max_split = 1;
    if (nsplit > max_split) nsplit = (int)max_split;
    if (nsplit < 1) nsplit = 1;
    dim3 grid(ceil_div(Cout, FB), ceil_div(Cin, FB), ntaps * nsplit);
    while (nsplit > 1 && (size_t)nsplit * ntaps * Cout * Cin > umma_wgrad_scratch_floats()) --nsplit;
    grid.z = ntaps * nsplit;
    wgrad_fma_kernel<<<grid, 256, 0, s>>>(dY, X, rows, Cout, Cin, ntaps, nsplit, scratch);
    SIMQ_LAUNCH_CHECK();
    return k_wgrad_reduce(scratch, Cout, Cin, ntaps, nsplit, dW, s);      // fixed summation order: deterministic
}

// ------------------------------------------------------------------------------------------
// stem: 7x7 stride 2 pad 3, C -> 64, no bias (resnet.py:55-56).  x: NCHW or NHWC f32 [B,.,96,96]
// raw0: NHWC f32 [B,48,48,64].  One CTA = one 8x8 output tile of one image, all 64 output channels.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float stem_x(const float* __restrict__ x, int layout, int C, int n, int c, int y, int xx) {
    if (y < 0 || y >= 96 || xx < 0 || xx >= 96) return 0.f;
    if (layout == 0) return x[(((size_t)n * C + c) * 96 + y) * 96 + xx];
    const int pix = layout == 2 ? C + 1 : C;      // layout 2: NHWC with one trailing non-input channel per pixel
    return x[(((size_t)n * 96 + y) * 96 + xx) * pix + c];
}

__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ x, int layout, int C,
                                                        const float* __restrict__ w, float* __restrict__ raw0) {
    __shared__ float patch[21][22];
    __shared__ __align__(16) float ws[49][64];
    const int tid = threadIdx.x;
    const int n = blockIdx.z, ty0 = blockIdx.y * 8, tx0 = blockIdx.x * 8;
    const int pix = tid & 63, cog = tid >> 6, py = pix >> 3, px = pix & 7;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int c = 0; c < C; ++c) {
        __syncthreads();
        for (int i = tid; i < 441; i += 256) {
            int r = i / 21, q = i % 21;
            patch[r][q] = stem_x(x, layout, C, n, c, 2 * ty0 - 3 + r, 2 * tx0 - 3 + q);
        }
        for (int i = tid; i < 49 * 64; i += 256) {
            int co = i & 63, tap = i >> 6;
            ws[tap][co] = w[((size_t)co * C + c) * 49 + tap];
        }
        __syncthreads();
#pragma unroll 1
        for (int ky = 0; ky < 7; ++ky)
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                float v = patch[2 * py + ky][2 * px + kx];
                const float4* wp = reinterpret_cast<const float4*>(&ws[ky * 7 + kx][cog * 16]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 w4 = wp[q];
                    acc[q * 4 + 0] = fmaf(v, w4.x, acc[q * 4 + 0]);
                    acc[q * 4 + 1] = fmaf(v, w4.y, acc[q * 4 + 1]);
                    acc[q * 4 + 2] = fmaf(v, w4.z, acc[q * 4 + 2]);
                    acc[q * 4 + 3] = fmaf(v, w4.w, acc[q * 4 + 3]);
                }
            }
    }
    float* o = raw0 + (((size_t)n * 48 + ty0 + py) * 48 + tx0 + px) * 64 + cog * 16;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(o + q * 4) = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
}

int k_stem_conv(const float* x, int x_layout, int B, int C, const float* w, float* raw0, cudaStream_t s) {
    dim3 grid(6, 6, B);
    stem_conv_kernel<<<grid, 256, 0, s>>>(x, x_layout, C, w, raw0);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// stem wgrad: dW[co][c*49+tap] = sum_{n,oy,ox} dY0[n,oy,ox,co] * x[n,c,2oy+ky-3,2ox+kx-3]
// GEMM M=64 (co) x N=C*49 (<=490, padded to 512) x K=positions; K chunks of 16 output pixels of one row,
// im2col'ed into shared memory.  Persistent CTAs write partial dW; k_reduce_partials sums them in a fixed
// order (deterministic).
#define SW_BLOCKS 296
size_t stem_wgrad_partial_floats(int C) { return (size_t)SW_BLOCKS * 64 * C * 49; }

__global__ void __launch_bounds__(256, 1) stem_wgrad_kernel(const float* __restrict__ x, int layout, int B, int C,
                                                            const float* __restrict__ dy0, float* __restrict__ partials) {
    extern __shared__ float smem[];
    float* Ys = smem;                 // [16][64]
    float* Xc = smem + 16 * 64;       // [16][512]
    const int tid = threadIdx.x, cog = tid & 7, jg = tid >> 3;     // 8 co-groups x 32 j-groups
    const int NJ = C * 49;
    float acc[8][16];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
    const int nchunks = B * 48 * 3;
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        int n = ch / 144, r = ch % 144, oy = r / 3, ox0 = (r % 3) * 16;
        __syncthreads();
        for (int i = tid; i < 16 * 64; i += 256) {
            int pp = i >> 6, co = i & 63;
            Ys[i] = dy0[(((size_t)n * 48 + oy) * 48 + ox0 + pp) * 64 + co];
        }
        for (int i = tid; i < 16 * 512; i += 256) {
            int pp = i >> 9, j = i & 511;
            float v = 0.f;
            if (j < NJ) {
                int c = j / 49, tap = j % 49, ky = tap / 7, kx = tap % 7;
                v = stem_x(x, layout, C, n, c, 2 * oy + ky - 3, 2 * (ox0 + pp) + kx - 3);
            }
            Xc[i] = v;
        }
        __syncthreads();
#pragma unroll 2
        for (int pp = 0; pp < 16; ++pp) {
            float a[8], b[16];
            const float4* ap = reinterpret_cast<const float4*>(Ys + pp * 64 + cog * 8);
            float4 a0 = ap[0], a1 = ap[1];
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            const float4* bp = reinterpret_cast<const float4*>(Xc + pp * 512 + jg * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) { float4 b4 = bp[q]; b[q * 4] = b4.x; b[q * 4 + 1] = b4.y; b[q * 4 + 2] = b4.z; b[q * 4 + 3] = b4.w; }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    float* out = partials + (size_t)blockIdx.x * 64 * NJ;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            int jj = jg * 16 + j;
            if (jj < NJ) out[(size_t)(cog * 8 + i) * NJ + jj] = acc[i][j];
        }
}

int k_stem_wgrad(const float* x, int x_layout, int B, int C, const float* dy0, float* partials, float* dW,
                 cudaStream_t s) {
    size_t smem = sizeof(float) * (16 * 64 + 16 * 512);
    static unsigned long long attr_set = 0;
    if (first_use_on_device(attr_set)) SIMQ_CUDA(cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stem_wgrad_kernel<<<SW_BLOCKS, 256, smem, s>>>(x, x_layout, B, C, dy0, partials);
    SIMQ_LAUNCH_CHECK();
    return k_reduce_partials(partials, SW_BLOCKS, 64 * C * 49, dW, 1.0f, s);
}

// ------------------------------------------------------------------------------------------
// stem on the tensor cores: im2col of the 7x7/2 windows into a split-bf16 [B*2304][Kp] matrix
// (column j = (ky*7 + kx)*C + c -- tap-major, channel-minor, so that with NHWC input the 8 columns a thread
// writes are an almost contiguous run of the input window; Kp = 49*C rounded up to 64, zero padded) so
// that conv1 and its weight gradient are plain GEMMs for the tcgen05 kernels.  k_pack_stem / k_strip_stem
// translate between this column order and the OIHW flattening (c*49 + tap) of resnet18.conv1.weight.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, int layout, int B, int C, int Kp, Split acol) {
    const unsigned G = (unsigned)Kp >> 3;               // 32-bit index arithmetic (host checks B*2304*G < 2^32)
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)B * 2304u * G) return;
    const unsigned p = idx / G;
    const int g = (int)(idx - p * G);
    const int n = (int)(p / 2304u), r = (int)(p - (unsigned)n * 2304u), oy = r / 48, ox = r - oy * 48;
    const int NJ = C * 49;
    float v[8];
    int tap = (g * 8) / C, c = g * 8 - tap * C;          // one division per thread; (tap, c) advance incrementally
    int ky = tap / 7, kx = tap - ky * 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v[i] = (g * 8 + i < NJ) ? stem_x(x, layout, C, n, c, 2 * oy + ky - 3, 2 * ox + kx - 3) : 0.f;
        if (++c == C) { c = 0; if (++kx == 7) { kx = 0; ++ky; } }
    }
    store8_split(acol, (size_t)p * Kp + g * 8, v);
}

// Staged variant for C <= IM2COL_MAX_C: one block = one output row (48 pixels) of one image.  The 7 input rows it reads are staged in
// shared memory with coalesced loads ([ky][3 + ix][c], zero halo of 3 pixels either side), so the K elements of an output pixel for
// one kernel row ky -- 7 taps x C channels -- are ONE contiguous run of 7C floats starting at pixel 2*ox: building 8 consecutive K
// elements is 8 bank-conflict-free shared loads instead of 8 bounds-checked global gathers (137 -> ~55 us at batch 128, C = 5).
#define IM2COL_MAX_C 12
__global__ void __launch_bounds__(256) stem_im2col_rows_kernel(const float* __restrict__ x, int layout, int C, int Kp, Split acol) {
    __shared__ float tile[7 * 102 * IM2COL_MAX_C];
    const int n = blockIdx.y, oy = blockIdx.x;
    const int rowf = 102 * C;                            // floats per staged row
    if (layout == 0) {                                   // NCHW: generic gather
        for (int i = threadIdx.x; i < 7 * rowf; i += 256) {
            const int ky = i / rowf, r = i - ky * rowf, px = r / C, c = r - px * C;
            tile[i] = stem_x(x, layout, C, n, c, 2 * oy + ky - 3, px - 3);
        }
    } else {                                             // NHWC: one thread copies the C contiguous channels of a pixel (no per-element index math)
        const int pixf = layout == 2 ? C + 1 : C;
        for (int p = threadIdx.x; p < 7 * 102; p += 256) {
            const int ky = p / 102, px = p - ky * 102;
            const int iy = 2 * oy + ky - 3, ix = px - 3;
            float* dst = tile + ky * rowf + px * C;
            if (iy < 0 || iy >= 96 || ix < 0 || ix >= 96) {
                for (int c = 0; c < C; ++c) dst[c] = 0.f;
            } else {
                const float* src = x + (((size_t)n * 96 + iy) * 96 + ix) * pixf;
                for (int c = 0; c < C; ++c) dst[c] = src[c];
            }
        }
    }
    __syncthreads();
    const int G = Kp >> 3, NJ = C * 49, run = 7 * C;
    for (int item = threadIdx.x; item < 48 * G; item += 256) {
        const int ox = item / G, g = item - ox * G;
        int k = g * 8;
        int ky = k / run, rem = k - ky * run;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] = (k + i < NJ) ? tile[ky * rowf + 2 * ox * C + rem] : 0.f;
            if (++rem == run) { rem = 0; ++ky; }
        }
        store8_split(acol, ((size_t)n * 2304 + oy * 48 + ox) * Kp + g * 8, v);
    }
}

int k_stem_im2col(const float* x, int x_layout, int B, int C, int Kp, Split acol, cudaStream_t s) {
    long long n = (long long)B * 2304 * (Kp / 8);
    if (n >= (1LL << 32)) { simq_set_error("k_stem_im2col: batch too large for 32-bit indexing"); return 1; }
    if (C <= IM2COL_MAX_C) stem_im2col_rows_kernel<<<dim3(48, B), 256, 0, s>>>(x, x_layout, C, Kp, acol);
    else stem_im2col_kernel<<<ceil_div(n, 256), 256, 0, s>>>(x, x_layout, B, C, Kp, acol);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// w OIHW [64][C*49] f32 -> split [64][Kp], zero padded
__global__ void pack_stem_kernel(const float* __restrict__ w, int C, int Kp, Split out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 64 * Kp) return;
    const int NJ = C * 49;
    int co = idx / Kp, j = idx % Kp;
    int tap = j / C, c = j - tap * C;
    float v = j < NJ ? w[co * NJ + c * 49 + tap] : 0.f;
    split_store(v, out.hi[idx], out.lo[idx]);
}
int k_pack_stem(const float* w, int C, int Kp, Split out, cudaStream_t s) {
    pack_stem_kernel<<<ceil_div(64 * Kp, 256), 256, 0, s>>>(w, C, Kp, out);
    SIMQ_LAUNCH_CHECK();
    return 0;
}
// dW [64][C*49] <- tmp [64][Kp]
__global__ void strip_stem_kernel(const float* __restrict__ tmp, int C, int Kp, float* __restrict__ dW) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int NJ = C * 49;
    if (idx >= 64 * NJ) return;
    int co = idx / NJ, r = idx % NJ, c = r / 49, tap = r - c * 49;     // OIHW element (co, c, tap)
    dW[idx] = tmp[co * Kp + tap * C + c];
}
int k_strip_stem(const float* tmp, int C, int Kp, float* dW, cudaStream_t s) {
    strip_stem_kernel<<<ceil_div(64 * C * 49, 256), 256, 0, s>>>(tmp, C, Kp, dW);
    SIMQ_LAUNCH_CHECK();
    return 0;
}
