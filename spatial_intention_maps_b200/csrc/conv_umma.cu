// tcgen05 / TMA convolutions for the 24x24 stages (sm_100a only).
//
// Forward conv and dgrad (k_conv_umma): the pitch-25 layout (common.cuh) turns a 3x3 conv into 9
// row-shifted GEMMs  out[m][n] = sum_t sum_k A[m + off_t][k] * W[t][n][k]  over M = B*625 rows.
// One CTA owns a 128 x BN output tile; a TMA producer warp streams, per (tap, 64-wide K chunk), the
// hi and lo bf16 planes of the shifted A rows and of the W tile into a SWIZZLE_128B shared-memory
// ring (out-of-range rows are zero-filled by TMA = the conv padding before the first / after the
// last image); one thread issues three tcgen05.mma per K=16 step -- lo*hi, hi*lo, hi*hi -- into one
// fp32 accumulator in TMEM (value = hi + lo carries 16 mantissa bits); four
// epilogue warps read TMEM back with tcgen05.ld, zero the pitch-25 halo rows, optionally add a
// residual / previous gradient and write fp32.
//
// wgrad (k_wgrad_umma): dW[t][co][ci] = sum_p dY[p][co] * X[p + off_t][ci] is a GEMM whose
// contraction runs over the rows, i.e. both operands are MN-major in shared memory (rows of 64
// channels = 128 bytes, SWIZZLE_128B).  Grid = (co tiles, ci tiles, taps x row splits); every CTA
// writes its partial tile to a scratch buffer, a second kernel sums the splits in a fixed order
// (deterministic) into the OIHW gradient.
#include "kernels.h"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if (spin == 1024) t0 = clock64();
        if (spin > 1024 && (spin & 1023) == 0 && clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread l of the warp receives row (lane base + l)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// v = columns [taddr, +32) + columns [taddr + stride, +32): the two partial accumulators of an N-stacked MMA (see Conv1WCfg::STACK);
// both loads are in flight before the one wait
__device__ __forceinline__ void tmem_ld32_sum2(uint32_t taddr, uint32_t stride, float* v) {
    uint32_t r[32], q[32];
#define SIMQ_LD32(R, ADDR)                                                                                                      \
    asm volatile(                                                                                                               \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                               \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                               \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                               \
        : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]), "=r"(R[8]),           \
          "=r"(R[9]), "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15]), "=r"(R[16]),                \
          "=r"(R[17]), "=r"(R[18]), "=r"(R[19]), "=r"(R[20]), "=r"(R[21]), "=r"(R[22]), "=r"(R[23]), "=r"(R[24]),               \
          "=r"(R[25]), "=r"(R[26]), "=r"(R[27]), "=r"(R[28]), "=r"(R[29]), "=r"(R[30]), "=r"(R[31])                             \
        : "r"(ADDR)                                                                                                             \
        : "memory")
    SIMQ_LD32(r, taddr);
    SIMQ_LD32(q, taddr + stride);
#undef SIMQ_LD32
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + __uint_as_float(q[i]);
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
// layout_type: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10),
// a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__host__ __device__ constexpr int tmem_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

#define UM_BM 128
#define UM_BK 64
#define UM_THREADS 192          // warp 0: TMA producer, warp 1: TMEM alloc + MMA issue, warps 2-5: epilogue

// true in exactly one lane of a converged warp (elect.sync): the MMA warp runs its loop warp-wide -- every lane waits on the
// barriers -- and only the elected lane issues; ptxas then predicates the tcgen05 instructions instead of wrapping each one in a
// "for every active lane" loop, which is what a divergent `if (lane == 0)` region costs (5 extra instructions per MMA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps

#define EPI_LD 36               // floats per staged row (32 + 4 pad: conflict-free float4 row writes and column reads)

// BK = 64: 128-byte K rows, SWIZZLE_128B.  (A BK = 32 / SWIZZLE_64B variant and a single-CTA 128 x 256 tile were measured
// and dropped: no gain over these tiles; large-N layers use the CTA-pair kernel instead.)
// TERMS = 3: split-bf16 operands, lo*hi + hi*lo + hi*hi (parity mode).  TERMS = 1: hi planes only, one MMA per
// product (plain bf16 "fast" mode: ~3x less tensor work, fails the 1e-3 Q-map bar -- opt-in, see simq_set_precision).
// TERMS = 2 (backward GEMMs only, opt-in): the A operand -- the output gradient dy -- contributes its hi plane only,
// the other operand stays split: hi*lo + hi*hi, two MMAs per product and no dy.lo loads (gradients carry no parity bar in
// the north_star; the Q-map / arg-max path never uses it).
template <int BN, int TERMS = 3>
struct ConvCfg {
    static constexpr int BK = 64;
    static constexpr int A_BYTES = UM_BM * BK * 2;            // one plane of the A tile
    static constexpr int W_BYTES = BN * BK * 2;               // one plane of the W tile
    static constexpr uint32_t LAYOUT = BK == 64 ? 2u : 4u;    // SmemDescriptor layout_type
    static constexpr uint32_t SBO = BK == 64 ? 1024u : 512u;  // 8 rows of BK*2 bytes
    static constexpr int A_PLANES = TERMS == 3 ? 2 : 1;
    static constexpr int W_PLANES = TERMS >= 2 ? 2 : 1;
    static constexpr int W_OFF = A_PLANES * A_BYTES;          // stage layout: A hi [, A lo], W hi [, W lo]
    static constexpr int STAGE_BYTES = A_PLANES * A_BYTES + W_PLANES * W_BYTES;
    static constexpr int STAGING_BYTES = 4 * 32 * EPI_LD * 4; // per epilogue warp: 32 rows x 32 columns fp32
    static constexpr int CSUM_BYTES = 5 * 2 * BN * 4;         // per epilogue warp column sums / sums of squares + the CTA's running total
    static constexpr int FIXED = STAGING_BYTES + CSUM_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int STAGES = (227 * 1024 - FIXED) / STAGE_BYTES > 8 ? 8 : (227 * 1024 - FIXED) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FIXED;
    static constexpr bool STACK = TERMS >= 2;                 // a.hi * [w.hi ; w.lo] as ONE N = 2 BN instruction, see Conv1WCfg::STACK
    static constexpr int ACC_COLS = STACK ? 2 * BN : BN;
    static constexpr int TMEM_COLS = 2 * ACC_COLS;            // double-buffered accumulator
};

// ------------------------------------------------------------------------------------------------
// forward conv / dgrad: persistent CTAs (one per SM) walk the (m_tile, n_tile) list; the accumulator is
// double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
// ------------------------------------------------------------------------------------------------
// epilogue feature flags (compile-time: the epilogue is the bottleneck of the short-K layers)
enum { EF_STATS = 1, EF_AFFINE = 2, EF_PREV = 4, EF_RES = 8, EF_G = 16, EF_RELU = 32, EF_F32 = 64, EF_SPLIT = 128, EF_BNBWD = 256 };

// ------------------------------------------------------------------------------------------------
// epilogue of one 128 x BN tile (4 warps; warp `quad` owns TMEM lanes 32*quad..+31 = tile rows 32*quad..+31):
// tcgen05.ld 32 columns -> smem staging -> [column sums] -> transforms -> coalesced row stores.
// CLUSTER: the accumulator-free barrier lives in CTA 0 of the pair (remote arrive).
// ------------------------------------------------------------------------------------------------
// Column statistics (EF_STATS / EF_BNBWD): stat_per_cta != 0 -> the tile's sums are added to the CTA's running total (csum + 8*BN,
// fixed tile order = deterministic) and written ONCE per CTA by flush_cta_stats: the finalize / reduce kernel that follows reads
// <= 148 partial rows instead of one per m-tile (625 at batch 128); 0 -> one partial row per m-tile (CTAs whose tiles span several
// column blocks).
template <int BN, int FL, bool CLUSTER, bool STACK = false>
__device__ __forceinline__ void epilogue_tile(float* staging, float* csum, uint32_t tmem_acc, uint32_t tempty, int m_t, int n0,
                                              int rows, int N, float* __restrict__ out, const ConvEpilogue& ep, int quad, int lane,
                                              int stat_per_cta) {
    float* stg = staging + quad * 32 * EPI_LD;
    float* cs = csum + quad * 2 * BN;
    const int sr = lane >> 3, scol = (lane & 7) * 4;  // store phase: 4 rows per instruction, float4 per lane
    const int mw = m_t * UM_BM + quad * 32;           // first row of this warp
    const int m = mw + lane;
    const bool valid = m < rows && !(ep.pitch25 && !p25_valid(m % IMG25));
    const uint32_t vmask = __ballot_sync(0xffffffffu, valid);      // bit r: row mw + r carries data
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
        float v[32];
        if (STACK) tmem_ld32_sum2(tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, (uint32_t)BN, v);
        else tmem_ld32(tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, v);
        if (c + 32 >= BN) {                     // last read of this accumulator: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CLUSTER) {
                    asm volatile(
                        "{\n\t.reg .b32 ra;\n\t"
                        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
                        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(tempty)
                        : "memory");
                } else {
                    mbar_arrive(tempty);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            float4 w4 = valid ? make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(stg + lane * EPI_LD + i) = w4;
        }
        __syncwarp();
        if (FL & EF_STATS) {                    // lane = column: sums over this warp's 32 rows
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) { float x = stg[r * EPI_LD + lane]; s1 += x; s2 = fmaf(x, x, s2); }
            cs[c + lane] = s1; cs[BN + c + lane] = s2;
        }
        const int n = n0 + c + scol;
        float4 sc4 = make_float4(1.f, 1.f, 1.f, 1.f), sh4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (FL & EF_AFFINE) { sc4 = *reinterpret_cast<const float4*>(ep.scale + n); sh4 = *reinterpret_cast<const float4*>(ep.shift + n); }
        const size_t o0 = (size_t)(mw + sr) * N + n;
        // global operands of the transforms, all 8 row groups up front: issued back to back, their latencies overlap (inside
        // the store loop every load would wait behind the previous group's store -- add_prev may alias out -- and the
        // epilogue of a short-K launch becomes a chain of ~64 dependent DRAM round trips per tile)
        float4 pv[8], gv[8];
        uint2 rh[8], rl[8], gm[8];
        if (FL & (EF_PREV | EF_RES | EF_G)) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + sr;
                const size_t o = o0 + (size_t)(i * 4) * N;
                const bool live = mw + r < rows && ((vmask >> r) & 1u);
                if (FL & EF_PREV) pv[i] = live ? *reinterpret_cast<const float4*>(ep.add_prev + o) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (FL & EF_RES) {
                    rh[i] = live ? *reinterpret_cast<const uint2*>(ep.res.hi + o) : make_uint2(0u, 0u);
                    rl[i] = live ? *reinterpret_cast<const uint2*>(ep.res.lo + o) : make_uint2(0u, 0u);
                }
                if (FL & EF_G) {
                    gv[i] = live ? *reinterpret_cast<const float4*>(ep.add_g + o) : make_float4(0.f, 0.f, 0.f, 0.f);
                    gm[i] = live ? *reinterpret_cast<const uint2*>(ep.add_g_mask + o) : make_uint2(0u, 0u);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + sr;
            if (mw + r >= rows) continue;
            float4 x = *reinterpret_cast<const float4*>(stg + r * EPI_LD + scol);
            const size_t o = o0 + (size_t)(i * 4) * N;
            if ((vmask >> r) & 1u) {
                if (FL & EF_AFFINE) { x.x = fmaf(x.x, sc4.x, sh4.x); x.y = fmaf(x.y, sc4.y, sh4.y); x.z = fmaf(x.z, sc4.z, sh4.z); x.w = fmaf(x.w, sc4.w, sh4.w); }
                if (FL & EF_PREV) { const float4 p = pv[i]; x.x += p.x; x.y += p.y; x.z += p.z; x.w += p.w; }
                if (FL & EF_RES) {
                    const bf16* hb = reinterpret_cast<const bf16*>(&rh[i]); const bf16* lb = reinterpret_cast<const bf16*>(&rl[i]);
                    x.x += bf2f(hb[0]) + bf2f(lb[0]); x.y += bf2f(hb[1]) + bf2f(lb[1]);
                    x.z += bf2f(hb[2]) + bf2f(lb[2]); x.w += bf2f(hb[3]) + bf2f(lb[3]);
                }
                if (FL & EF_G) {
                    const float4 g = gv[i];
                    const bf16* mb = reinterpret_cast<const bf16*>(&gm[i]);
                    if (bf2f(mb[0]) > 0.f) x.x += g.x;
                    if (bf2f(mb[1]) > 0.f) x.y += g.y;
                    if (bf2f(mb[2]) > 0.f) x.z += g.z;
                    if (bf2f(mb[3]) > 0.f) x.w += g.w;
                }
                if (FL & EF_RELU) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
            } else {
                x = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (FL & EF_F32) *reinterpret_cast<float4*>(out + o) = x;
            if (FL & EF_BNBWD) *reinterpret_cast<float4*>(stg + r * EPI_LD + scol) = x;     // final gradient back to staging
            if (FL & EF_SPLIT) {
                bf16 h[4], l[4];
                split_store(x.x, h[0], l[0]); split_store(x.y, h[1], l[1]); split_store(x.z, h[2], l[2]); split_store(x.w, h[3], l[3]);
                *reinterpret_cast<uint2*>(ep.out_split.hi + o) = *reinterpret_cast<const uint2*>(h);
                *reinterpret_cast<uint2*>(ep.out_split.lo + o) = *reinterpret_cast<const uint2*>(l);
            }
        }
        __syncwarp();                           // staging is rewritten by the next chunk
        if (FL & EF_BNBWD) {                    // lane = column: BatchNorm-backward sums of this warp's 32 rows
            const int col = n0 + c + lane;
            const float mu = ep.bn_mean[col], is = ep.bn_invstd[col];
            float s1 = 0.f, s2 = 0.f;
            const int rmax = rows - mw < 32 ? rows - mw : 32;
#pragma unroll 8
            for (int r = 0; r < rmax; ++r) {
                const size_t o = (size_t)(mw + r) * N + col;
                const float g = stg[r * EPI_LD + lane];
                const float xr = ep.bn_raw[o];
                const float dz = bf2f(ep.bn_mask[o]) > 0.f ? g : 0.f;
                s1 += dz;
                s2 = fmaf(dz, (xr - mu) * is, s2);
            }
            cs[c + lane] = s1; cs[BN + c + lane] = s2;
            __syncwarp();
        }
    }
    if (FL & (EF_STATS | EF_BNBWD)) {           // combine the 4 warps in a fixed order -> one partial row per m_tile
        epi_bar_sync();
        const int t = threadIdx.x - 64;         // 0..127
        for (int j = t; j < 2 * BN; j += 128) {
            float tot = csum[j] + csum[2 * BN + j] + csum[4 * BN + j] + csum[6 * BN + j];
            if (stat_per_cta) {
                csum[8 * BN + j] += tot;            // always thread t for column j: no race
            } else {
                const int which = j / BN, col = j - which * BN;
                ep.stats[((size_t)m_t * 2 + which) * N + n0 + col] = tot;
            }
        }
        epi_bar_sync();
    }
}
template <int BN>
__device__ __forceinline__ void zero_cta_stats(float* csum) {
    for (int j = threadIdx.x - 64; j < 2 * BN; j += 128) csum[8 * BN + j] = 0.f;
}
template <int BN>
__device__ __forceinline__ void flush_cta_stats(const float* csum, float* __restrict__ stats, int row, int n0, int N) {
    for (int j = threadIdx.x - 64; j < 2 * BN; j += 128) {
        const int which = j / BN, col = j - which * BN;
        stats[((size_t)row * 2 + which) * N + n0 + col] = csum[8 * BN + j];
    }
}
// The 4 epilogue warps of the LAST CTA to finish reduce the launch's partial rows [nrows][2][N] (double accumulation) and do what
// the follow-up kernel used to do (ConvEpilogue::fin_mode).  Thread t owns float4 column group t % (N/4) and row group t / (N/4):
// narrow layers (N = 64: 16 column groups) split the rows over 8 row groups, so every thread has loads in flight; the row groups
// are combined through shared memory in group order.  Everything is a fixed order: the result does not depend on which CTA this is.
__device__ __noinline__ void stats_finalize(const ConvEpilogue& ep, int nrows, int N, double* red /* >= 128 * 8 doubles */) {
    const int t = threadIdx.x - 64;
    const float* st = ep.stats;
    const int cgs = N >> 2;                            // float4 column groups: 8 .. 128
    const int rgs = cgs >= 128 ? 1 : 128 / cgs;        // row groups
    const int cg = t % cgs, rg = t / cgs, c0 = cg * 4;
    double S[4] = {0, 0, 0, 0}, Q[4] = {0, 0, 0, 0};
    if (rg < rgs) {
        int r = rg;
        for (; r + 15 * rgs < nrows; r += 16 * rgs) {          // 32 float4 loads in flight per thread: the wide layers are latency-bound here
            float4 a[16], q[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                a[u] = __ldcg(reinterpret_cast<const float4*>(st + ((size_t)(r + u * rgs) * 2 + 0) * N + c0));
                q[u] = __ldcg(reinterpret_cast<const float4*>(st + ((size_t)(r + u * rgs) * 2 + 1) * N + c0));
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                S[0] += a[u].x; S[1] += a[u].y; S[2] += a[u].z; S[3] += a[u].w;
                Q[0] += q[u].x; Q[1] += q[u].y; Q[2] += q[u].z; Q[3] += q[u].w;
            }
        }
        for (; r + 7 * rgs < nrows; r += 8 * rgs) {
            float4 a[8], q[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a[u] = __ldcg(reinterpret_cast<const float4*>(st + ((size_t)(r + u * rgs) * 2 + 0) * N + c0));
                q[u] = __ldcg(reinterpret_cast<const float4*>(st + ((size_t)(r + u * rgs) * 2 + 1) * N + c0));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                S[0] += a[u].x; S[1] += a[u].y; S[2] += a[u].z; S[3] += a[u].w;
                Q[0] += q[u].x; Q[1] += q[u].y; Q[2] += q[u].z; Q[3] += q[u].w;
            }
        }
        for (; r < nrows; r += rgs) {
            const float4 a = __ldcg(reinterpret_cast<const float4*>(st + ((size_t)r * 2 + 0) * N + c0));
            const float4 q = __ldcg(reinterpret_cast<const float4*>(st + ((size_t)r * 2 + 1) * N + c0));
            S[0] += a.x; S[1] += a.y; S[2] += a.z; S[3] += a.w;
            Q[0] += q.x; Q[1] += q.y; Q[2] += q.z; Q[3] += q.w;
        }
    }
    if (rgs > 1) {                                     // combine the row groups in group order
        if (rg < rgs) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { red[(rg * cgs + cg) * 8 + i] = S[i]; red[(rg * cgs + cg) * 8 + 4 + i] = Q[i]; }
        }
        epi_bar_sync();
        if (rg == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { S[i] = 0; Q[i] = 0; }
            for (int g = 0; g < rgs; ++g) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { S[i] += red[(g * cgs + cg) * 8 + i]; Q[i] += red[(g * cgs + cg) * 8 + 4 + i]; }
            }
        }
    }
    if (rg == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (ep.fin_mode == 1) {
                bn_finalize_channel(S[i], Q[i], ep.fin_count, c0 + i, ep.fin_gamma, ep.fin_beta, ep.fin_bias, ep.fin_rmean, ep.fin_rvar,
                                    ep.fin_defer, MAX_CH, ep.fin_mean, ep.fin_invstd, ep.fin_scale, ep.fin_shift);
            } else {
                ep.fin_out[c0 + i] = (float)S[i];
                ep.fin_out[N + c0 + i] = (float)Q[i];
            }
        }
    }
    if (ep.fin_mode == 1 && t == 0 && ep.fin_nbt && !ep.fin_defer) ep.fin_nbt[0] += 1;
}
// after a CTA's last tile (epilogue warps only): publish its statistics, take a ticket, and finalize if it is the last one
template <int BN>
__device__ __forceinline__ void stats_tail(const ConvEpilogue& ep, float* staging, float* csum, int stat_per_cta, bool has_tiles, int row, int n0,
                                           int N, int nrows) {
    if (stat_per_cta && has_tiles) flush_cta_stats<BN>(csum, ep.stats, row, n0, N);
    if (!ep.fin_ticket) return;
    __threadfence();                                   // this thread's partial sums are visible device-wide ...
    epi_bar_sync();                                    // ... for all 128 epilogue threads, before the CTA takes its ticket
    int* flag = reinterpret_cast<int*>(csum);          // (the per-warp area of csum is idle now)
    if (threadIdx.x == 64) *flag = atomicAdd(ep.fin_ticket, 1u) == gridDim.x - 1 ? 1 : 0;
    epi_bar_sync();
    if (*flag) {
        __threadfence();                               // acquire: every other CTA's rows were published before its ticket
        stats_finalize(ep, nrows, N, reinterpret_cast<double*>(staging));      // (the staging tile is idle: 18 KB >= 8 KB)
        if (threadIdx.x == 64) *ep.fin_ticket = 0;     // ready for the next launch on this lane
    }
}

template <int BN, int FL, int TERMS>
__global__ void __launch_bounds__(UM_THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
                 const __grid_constant__ CUtensorMap mWhi, const __grid_constant__ CUtensorMap mWlo, int rows, int K,
                 int N, int ntaps, int m_tiles, int n_tiles, int nz, float* __restrict__ out, ConvEpilogue ep) {
    using Cfg = ConvCfg<BN, TERMS>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    float* staging = reinterpret_cast<float*>(base_ptr + Cfg::STAGES * Cfg::STAGE_BYTES);
    float* csum = reinterpret_cast<float*>(base_ptr + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES);
    const uint32_t bars = base + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES + Cfg::CSUM_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (Cfg::STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * Cfg::STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * Cfg::STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * Cfg::STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int BK = Cfg::BK;
    const int kchunks = K / BK;
    // nz > 1: split-K over the taps for small problems (batch-1 inference: 5 m-tiles) -- work item = (tile, z), z owns
    // taps [z*ntaps/nz, (z+1)*ntaps/nz) and writes a raw partial to out + z*rows*N; a second kernel sums the partials
    // and applies the epilogue (k_conv_umma).  nz == 1: one work item per tile, all taps.
    const int taps_per_z = ntaps / nz;
    const int iters = taps_per_z * kchunks;
    const int total_tiles = m_tiles * n_tiles * nz;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&mAhi); tma_prefetch_desc(&mAlo); tma_prefetch_desc(&mWhi); tma_prefetch_desc(&mWlo);
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int work = blockIdx.x; work < total_tiles; work += gridDim.x) {
                const int tile = work / nz, z = work - tile * nz;
                const int m_t = tile / n_tiles, n0 = (tile - m_t * n_tiles) * BN;
                const int m0 = m_t * UM_BM;
                for (int it = 0; it < iters; ++it) {
                    const int tl = it / kchunks, kc = it - tl * kchunks, t = z * taps_per_z + tl;
                    const int off = ntaps == 9 ? (t / 3 - 1) * PITCH + (t % 3 - 1) : 0;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
                    const uint32_t sa = base + s * Cfg::STAGE_BYTES;
                    const int arow = m0 + off;
                    tma_load_2d(sa, &mAhi, full_bar(s), kc * BK, arow);
                    tma_load_2d(sa + Cfg::W_OFF, &mWhi, full_bar(s), kc * BK, t * N + n0);
                    if (TERMS == 3) tma_load_2d(sa + Cfg::A_BYTES, &mAlo, full_bar(s), kc * BK, arow);
                    if (TERMS >= 2) tma_load_2d(sa + Cfg::W_OFF + Cfg::W_BYTES, &mWlo, full_bar(s), kc * BK, t * N + n0);
                    if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        {   // the whole warp walks the loop (every lane waits on the barriers), one elected lane issues: see elect_one()
            constexpr uint32_t idesc = umma_idesc(UM_BM, BN, 0, 0);
            constexpr uint32_t idesc2 = umma_idesc(UM_BM, 2 * BN, 0, 0);      // [W.hi ; W.lo] as one B operand
            int s = 0; uint32_t ph = 0;
            int acc = 0; uint32_t aph = 0;
            for (int work = blockIdx.x; work < total_tiles; work += gridDim.x) {
                mbar_wait(tempty_bar(acc), aph ^ 1u);           // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * Cfg::ACC_COLS);
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t sa = base + s * Cfg::STAGE_BYTES;
                    // one descriptor per operand plane; the K = 16 steps advance its start-address field ((addr >> 4), no carry: smem < 256 KB)
                    const uint64_t a_hi0 = umma_desc(sa, 16, Cfg::SBO, Cfg::LAYOUT), w_hi0 = umma_desc(sa + Cfg::W_OFF, 16, Cfg::SBO, Cfg::LAYOUT);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t a_hi = a_hi0 + (uint64_t)(k * 2), w_hi = w_hi0 + (uint64_t)(k * 2);
                            if (TERMS == 3) {
                                const uint64_t a_lo = a_hi + (uint64_t)(Cfg::A_BYTES >> 4);
                                tc_mma_bf16(d_tmem, a_hi, w_hi, idesc2, (it | k) != 0);    // a.hi * [w.hi ; w.lo]: the first one zeroes both halves
                                tc_mma_bf16(d_tmem, a_lo, w_hi, idesc, 1);                 // a.lo * w.hi into the first half
                            } else if (TERMS == 2) {
                                tc_mma_bf16(d_tmem, a_hi, w_hi, idesc2, (it | k) != 0);
                            } else {
                                tc_mma_bf16(d_tmem, a_hi, w_hi, idesc, (it | k) != 0);
                            }
                        }
                        tc_commit(empty_bar(s));        // frees the stage once these MMAs have read it
                        if (it == iters - 1) tc_commit(tfull_bar(acc));
                    }
                    __syncwarp();
                    if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
                }
                if (++acc == 2) { acc = 0; aph ^= 1u; }
            }
        }
    } else {
        const int quad = warp & 3;                      // TMEM lane quadrant this warp may read
        int acc = 0; uint32_t aph = 0;
        constexpr bool HAS_STATS = (FL & (EF_STATS | EF_BNBWD)) != 0;
        const int stat_per_cta = HAS_STATS && nz == 1 && gridDim.x % n_tiles == 0;      // this CTA's tiles share one column block
        if (stat_per_cta) zero_cta_stats<BN>(csum);
        for (int work = blockIdx.x; work < total_tiles; work += gridDim.x) {
            const int tile = work / nz, z = work - tile * nz;
            const int m_t = tile / n_tiles, n0 = (tile - m_t * n_tiles) * BN;
            mbar_wait(tfull_bar(acc), aph);
            tc_fence_after();
            epilogue_tile<BN, FL, false, Cfg::STACK>(staging, csum, tmem_base + (uint32_t)(acc * Cfg::ACC_COLS), tempty_bar(acc), m_t, n0, rows, N,
                                                     out + (size_t)z * rows * N, ep, quad, lane, stat_per_cta);
            if (++acc == 2) { acc = 0; aph ^= 1u; }
        }
        if (HAS_STATS)
            stats_tail<BN>(ep, staging, csum, stat_per_cta, (int)blockIdx.x < total_tiles, blockIdx.x / n_tiles, ((int)blockIdx.x % n_tiles) * BN, N,
                           stat_per_cta ? (int)gridDim.x / n_tiles : m_tiles);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) for N % 256 == 0: a cluster of two CTAs on one TPC owns a 256 x 256 tile.
// Each CTA stages its own 128 A rows and HALF of the W tile (128 of the 256 output channels); one
// tcgen05.mma.cta_group::2 (M = 256, N = 256) issued by the leader CTA reads both halves, so per SM the
// L2->smem traffic per MMA cycle is half that of the 128 x 128 single-CTA tile (64 KB per 1536 tensor cycles)
// and the smem operand reads drop by a quarter.  Barriers: full[s] lives in the leader (both CTAs' TMA loads
// complete_tx on it), empty[s] / tfull[a] exist in both CTAs and are signalled by multicast tcgen05.commit,
// tempty[a] lives in the leader and collects the 8 epilogue warps of the pair.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t addr) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_commit_mc2(uint32_t bar) {      // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int TERMS>
struct Conv2Cfg {
    static constexpr int BN = 256;                            // pair tile: 256 rows x 256 columns
    static constexpr int A_BYTES = UM_BM * UM_BK * 2;         // this CTA's 128 A rows, one plane (16 KB)
    static constexpr int W_BYTES = (BN / 2) * UM_BK * 2;      // this CTA's half of the W tile, one plane (16 KB)
    static constexpr int A_PLANES = TERMS == 3 ? 2 : 1;
    static constexpr int W_PLANES = TERMS >= 2 ? 2 : 1;
    static constexpr int W_OFF = A_PLANES * A_BYTES;
    static constexpr int STAGE_BYTES = A_PLANES * A_BYTES + W_PLANES * W_BYTES;
    static constexpr int STAGING_BYTES = 4 * 32 * EPI_LD * 4;
    static constexpr int CSUM_BYTES = 5 * 2 * BN * 4;
    static constexpr int FIXED = STAGING_BYTES + CSUM_BYTES + 1024 + 256;
    static constexpr int STAGES = TERMS == 3 ? 3 : TERMS == 2 ? 4 : 6;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FIXED;
    static constexpr int TMEM_COLS = 512;                     // two 256-column accumulators
};

template <int FL, int TERMS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(UM_THREADS, 1)
conv2_umma_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
                  const __grid_constant__ CUtensorMap mWhi, const __grid_constant__ CUtensorMap mWlo, int rows, int K,
                  int N, int ntaps, int m2_tiles, int n_tiles, float* __restrict__ out, ConvEpilogue ep) {
    using Cfg = Conv2Cfg<TERMS>;
    constexpr int BN = Cfg::BN;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    float* staging = reinterpret_cast<float*>(base_ptr + Cfg::STAGES * Cfg::STAGE_BYTES);
    float* csum = reinterpret_cast<float*>(base_ptr + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES);
    const uint32_t bars = base + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::STAGING_BYTES + Cfg::CSUM_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (Cfg::STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * Cfg::STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * Cfg::STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * Cfg::STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int kchunks = K / UM_BK;
    const int iters = ntaps * kchunks;
    const int total_tiles = m2_tiles * n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&mAhi); tma_prefetch_desc(&mAlo); tma_prefetch_desc(&mWhi); tma_prefetch_desc(&mWlo);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();             // barrier inits + TMEM allocation of both CTAs visible to the pair
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
                const int m2 = tile / n_tiles, n0 = (tile - m2 * n_tiles) * BN;
                const int m0 = m2 * 256 + (int)rank * UM_BM;                 // this CTA's 128 rows
                const int wn0 = n0 + (int)rank * (BN / 2);                   // this CTA's half of the output channels
                for (int it = 0; it < iters; ++it) {
                    const int t = it / kchunks, kc = it - t * kchunks;
                    const int off = ntaps == 9 ? (t / 3 - 1) * PITCH + (t % 3 - 1) : 0;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    if (rank == 0) mbar_expect_tx(full_bar(s), 2 * Cfg::STAGE_BYTES);   // bytes of BOTH CTAs land on the leader's barrier
                    const uint32_t fb = mapa_rank0(full_bar(s));
                    const uint32_t sa = base + s * Cfg::STAGE_BYTES;
                    tma_load_2d_cg2(sa, &mAhi, fb, kc * UM_BK, m0 + off);
                    tma_load_2d_cg2(sa + Cfg::W_OFF, &mWhi, fb, kc * UM_BK, t * N + wn0);
                    if (TERMS == 3) tma_load_2d_cg2(sa + Cfg::A_BYTES, &mAlo, fb, kc * UM_BK, m0 + off);
                    if (TERMS >= 2) tma_load_2d_cg2(sa + Cfg::W_OFF + Cfg::W_BYTES, &mWlo, fb, kc * UM_BK, t * N + wn0);
                    if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {            // warp-wide loop, one elected lane issues (elect_one)
            constexpr uint32_t idesc = umma_idesc(256, BN, 0, 0);
            int s = 0; uint32_t ph = 0;
            int acc = 0; uint32_t aph = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
                mbar_wait(tempty_bar(acc), aph ^ 1u);           // both CTAs' epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t sa = base + s * Cfg::STAGE_BYTES;
                    const uint64_t a_hi0 = umma_desc(sa, 16, 1024), w_hi0 = umma_desc(sa + Cfg::W_OFF, 16, 1024);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < UM_BK / 16; ++k) {
                            const uint64_t a_hi = a_hi0 + (uint64_t)(k * 2), w_hi = w_hi0 + (uint64_t)(k * 2);
                            if (TERMS == 3) {
                                const uint64_t a_lo = a_hi + (uint64_t)(Cfg::A_BYTES >> 4), w_lo = w_hi + (uint64_t)(Cfg::W_BYTES >> 4);
                                tc_mma_bf16_cg2(d_tmem, a_lo, w_hi, idesc, (it | k) != 0);
                                tc_mma_bf16_cg2(d_tmem, a_hi, w_lo, idesc, 1);
                                tc_mma_bf16_cg2(d_tmem, a_hi, w_hi, idesc, 1);
                            } else if (TERMS == 2) {
                                const uint64_t w_lo = w_hi + (uint64_t)(Cfg::W_BYTES >> 4);
                                tc_mma_bf16_cg2(d_tmem, a_hi, w_lo, idesc, (it | k) != 0);
                                tc_mma_bf16_cg2(d_tmem, a_hi, w_hi, idesc, 1);
                            } else {
                                tc_mma_bf16_cg2(d_tmem, a_hi, w_hi, idesc, (it | k) != 0);
                            }
                        }
                        tc_commit_mc2(empty_bar(s));
                        if (it == iters - 1) tc_commit_mc2(tfull_bar(acc));
                    }
                    __syncwarp();
                    if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
                }
                if (++acc == 2) { acc = 0; aph ^= 1u; }
            }
        }
    } else {
        const int quad = warp & 3;
        int acc = 0; uint32_t aph = 0;
        constexpr bool HAS_STATS = (FL & (EF_STATS | EF_BNBWD)) != 0;
        const int stat_per_cta = HAS_STATS && num_clusters % n_tiles == 0;              // this pair's tiles share one column block
        if (stat_per_cta) zero_cta_stats<BN>(csum);
        for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
            const int m2 = tile / n_tiles, n0 = (tile - m2 * n_tiles) * BN;
            mbar_wait(tfull_bar(acc), aph);
            tc_fence_after();
            epilogue_tile<BN, FL, true>(staging, csum, tmem_base + (uint32_t)(acc * BN), tempty_bar(acc), m2 * 2 + (int)rank, n0, rows, N,
                                        out, ep, quad, lane, stat_per_cta);
            if (++acc == 2) { acc = 0; aph ^= 1u; }
        }
        if (HAS_STATS)
            stats_tail<BN>(ep, staging, csum, stat_per_cta, cluster_id < total_tiles, (cluster_id / n_tiles) * 2 + (int)rank, (cluster_id % n_tiles) * BN, N,
                           stat_per_cta ? 2 * (num_clusters / n_tiles) : 2 * m2_tiles);
    }
    tc_fence_before();
    cluster_sync_all();             // nobody leaves while the peer may still touch its barriers / TMEM
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// Windowed CTA-pair variant for the 3x3 convs: the 9 taps of a pitch-25 conv read the SAME activation rows shifted by
// -26..+26, so instead of re-loading the 128 A rows of a tile for every tap (9 x 32 KB per 64-wide K chunk) each CTA
// stages ONE window of 184 rows per K chunk -- tile rows -26..+157 -- and the MMA of tap t reads it through a shared
// memory descriptor whose start address is advanced by (26 + off_t) rows of 128 bytes.  SWIZZLE_128B permutes the
// 16-byte chunks of a row by (row mod 8), both in what TMA writes and in what the MMA reads, as a function of the
// absolute shared-memory address bits [7,10): a descriptor whose start address is advanced by whole rows therefore stays
// consistent with what TMA wrote, with the base-offset field left at 0 (measured: setting it to (addr >> 7) & 7 breaks
// the result, leaving it 0 is exact for all 9 row shifts).  The loop order becomes K chunk outer, tap inner; W tiles stream
// through their own ring.  L2 -> shared traffic per tensor cycle drops from 64 KB to 32 + 47/9 = 37 KB per stage
// (-42 %): the step is power-bound, fewer bytes moved per MMA = more clock for the MMAs.
// (The single-CTA 128 x 64/128 tiles of stages 1-2 use the same scheme since round 2: conv1w_umma_kernel below.)
// ------------------------------------------------------------------------------------------------

#define WIN_LEAD 26                                           // PITCH + 1: window row of tile row 0 at tap offset 0
#define WIN_ROWS 184                                          // 128 + 2 * 26 = 180, rounded up to the 8-row swizzle atom

template <int TERMS, int BN_ = 256>
struct Conv2WCfg {
    static constexpr int BN = BN_;                            // pair tile: 256 rows x BN columns (256: stages 3-4; 128 / 64: stages 2 / 1)
    static constexpr int A_BYTES = WIN_ROWS * UM_BK * 2;      // one plane of this CTA's window (23 KB)
    static constexpr int W_BYTES = (BN / 2) * UM_BK * 2;      // this CTA's half of one W tile, one plane (16 KB at BN = 256)
    static constexpr int A_PLANES = TERMS == 3 ? 2 : 1;
    static constexpr int W_PLANES = TERMS >= 2 ? 2 : 1;
    static constexpr int A_BUFS = 2;
    static constexpr int A_BUF_BYTES = A_PLANES * A_BYTES;
    static constexpr int W_STAGE_BYTES = W_PLANES * W_BYTES;
    static constexpr int W_STAGES = BN < 256 ? 6 : TERMS == 3 ? 3 : TERMS == 2 ? 4 : 6;
    static constexpr int W_BASE = A_BUFS * A_BUF_BYTES;
    static constexpr int RING_BYTES = W_BASE + W_STAGES * W_STAGE_BYTES;
    static constexpr int STAGING_BYTES = 4 * 32 * EPI_LD * 4;
    static constexpr int CSUM_BYTES = 5 * 2 * BN * 4;
    static constexpr int FIXED = STAGING_BYTES + CSUM_BYTES + 1024 + 256;
    static constexpr int SMEM_BYTES = RING_BYTES + FIXED;
    static constexpr int TMEM_COLS = 2 * BN;                  // two accumulators (power of two for every BN used)
};

template <int FL, int TERMS, int BN_>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(UM_THREADS, 1)
conv2w_umma_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
                   const __grid_constant__ CUtensorMap mWhi, const __grid_constant__ CUtensorMap mWlo, int rows, int K,
                   int N, int m2_tiles, int n_tiles, float* __restrict__ out, ConvEpilogue ep) {
    using Cfg = Conv2WCfg<TERMS, BN_>;
    constexpr int BN = Cfg::BN;
    static_assert(Cfg::A_BYTES % 1024 == 0, "window planes must keep the 1024-byte swizzle alignment");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    float* staging = reinterpret_cast<float*>(base_ptr + Cfg::RING_BYTES);
    float* csum = reinterpret_cast<float*>(base_ptr + Cfg::RING_BYTES + Cfg::STAGING_BYTES);
    const uint32_t bars = base + Cfg::RING_BYTES + Cfg::STAGING_BYTES + Cfg::CSUM_BYTES;
    auto afull_bar = [&](int b) { return bars + 8u * b; };
    auto aempty_bar = [&](int b) { return bars + 8u * (2 + b); };
    auto wfull_bar = [&](int s) { return bars + 8u * (4 + s); };
    auto wempty_bar = [&](int s) { return bars + 8u * (4 + Cfg::W_STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (4 + 2 * Cfg::W_STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (6 + 2 * Cfg::W_STAGES + a); };
    const uint32_t tmem_slot = bars + 8u * (8 + 2 * Cfg::W_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int kchunks = K / UM_BK;
    const int total_tiles = m2_tiles * n_tiles;
    // this pair's tiles: cluster_id, cluster_id + num_clusters, ...; a "step" g = one (tile, K chunk) = one A window
    const int my_tiles = cluster_id < total_tiles ? (total_tiles - cluster_id + num_clusters - 1) / num_clusters : 0;
    const int steps = my_tiles * kchunks;

    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) { mbar_init(afull_bar(b), 1); mbar_init(aempty_bar(b), 1); }
        for (int s = 0; s < Cfg::W_STAGES; ++s) { mbar_init(wfull_bar(s), 1); mbar_init(wempty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&mAhi); tma_prefetch_desc(&mAlo); tma_prefetch_desc(&mWhi); tma_prefetch_desc(&mWlo);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {
            // the window of step g: rows m0 - 26 .. m0 + 157 of K chunk kc (rows outside the tensor are zero-filled)
            auto load_window = [&](int g) {
                const int tile = cluster_id + (g / kchunks) * num_clusters, kc = g % kchunks;
                const int m0 = (tile / n_tiles) * 256 + (int)rank * UM_BM;
                const int b = g & 1;
                mbar_wait(aempty_bar(b), ((uint32_t)(g >> 1) & 1u) ^ 1u);
                if (rank == 0) mbar_expect_tx(afull_bar(b), 2 * Cfg::A_BUF_BYTES);
                const uint32_t fb = mapa_rank0(afull_bar(b));
                const uint32_t sa = base + b * Cfg::A_BUF_BYTES;
                tma_load_2d_cg2(sa, &mAhi, fb, kc * UM_BK, m0 - WIN_LEAD);
                if (TERMS == 3) tma_load_2d_cg2(sa + Cfg::A_BYTES, &mAlo, fb, kc * UM_BK, m0 - WIN_LEAD);
            };
            int s = 0; uint32_t ph = 0;
            if (steps > 0) load_window(0);
            for (int g = 0; g < steps; ++g) {
                const int tile = cluster_id + (g / kchunks) * num_clusters, kc = g % kchunks;
                const int n0 = (tile % n_tiles) * BN;
                const int wn0 = n0 + (int)rank * (BN / 2);
                for (int t = 0; t < 9; ++t) {
                    if (t == 3 && g + 1 < steps) load_window(g + 1);        // prefetch: its buffer was freed a full step ago
                    mbar_wait(wempty_bar(s), ph ^ 1u);
                    if (rank == 0) mbar_expect_tx(wfull_bar(s), 2 * Cfg::W_STAGE_BYTES);
                    const uint32_t fb = mapa_rank0(wfull_bar(s));
                    const uint32_t sw = base + Cfg::W_BASE + s * Cfg::W_STAGE_BYTES;
                    tma_load_2d_cg2(sw, &mWhi, fb, kc * UM_BK, t * N + wn0);
                    if (TERMS >= 2) tma_load_2d_cg2(sw + Cfg::W_BYTES, &mWlo, fb, kc * UM_BK, t * N + wn0);
                    if (++s == Cfg::W_STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc(256, BN, 0, 0);
            int s = 0; uint32_t ph = 0;
            int acc = 0; uint32_t aph = 0;
            for (int g = 0; g < steps; ++g) {
                const int kc = g % kchunks;
                if (kc == 0) {
                    mbar_wait(tempty_bar(acc), aph ^ 1u);
                    tc_fence_after();
                }
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                const int b = g & 1;
                mbar_wait(afull_bar(b), (uint32_t)(g >> 1) & 1u);
                const uint32_t sa = base + b * Cfg::A_BUF_BYTES;
                for (int t = 0; t < 9; ++t) {
                    const int off = (t / 3 - 1) * PITCH + (t % 3 - 1);
                    const uint32_t arow = sa + (uint32_t)(WIN_LEAD + off) * 128u;
                    mbar_wait(wfull_bar(s), ph);
                    tc_fence_after();
                    const uint32_t sw = base + Cfg::W_BASE + s * Cfg::W_STAGE_BYTES;
                    // descriptors of the K = 16 steps / lo planes differ from the first one only in the start-address field
                    // ((addr >> 4) in the low bits; shared memory is < 256 KB, so the field cannot carry): one add per descriptor
                    // instead of rebuilding it -- the issuing thread is the bottleneck of the narrow tiles (N <= 128)
                    const uint64_t a_hi0 = umma_desc(arow, 16, 1024), w_hi0 = umma_desc(sw, 16, 1024);
                    if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < UM_BK / 16; ++k) {
                        const uint64_t a_hi = a_hi0 + (uint64_t)(k * 2), w_hi = w_hi0 + (uint64_t)(k * 2);
                        const uint32_t accumulate = (kc | t | k) != 0;
                        if (TERMS == 3) {
                            const uint64_t a_lo = a_hi + (uint64_t)(Cfg::A_BYTES >> 4), w_lo = w_hi + (uint64_t)(Cfg::W_BYTES >> 4);
                            tc_mma_bf16_cg2(d_tmem, a_lo, w_hi, idesc, accumulate);
                            tc_mma_bf16_cg2(d_tmem, a_hi, w_lo, idesc, 1);
                            tc_mma_bf16_cg2(d_tmem, a_hi, w_hi, idesc, 1);
                        } else if (TERMS == 2) {
                            const uint64_t w_lo = w_hi + (uint64_t)(Cfg::W_BYTES >> 4);
                            tc_mma_bf16_cg2(d_tmem, a_hi, w_lo, idesc, accumulate);
                            tc_mma_bf16_cg2(d_tmem, a_hi, w_hi, idesc, 1);
                        } else {
                            tc_mma_bf16_cg2(d_tmem, a_hi, w_hi, idesc, accumulate);
                        }
                    }
                    tc_commit_mc2(wempty_bar(s));
                    }
                    __syncwarp();
                    if (++s == Cfg::W_STAGES) { s = 0; ph ^= 1u; }
                }
                if (elect_one()) {
                    tc_commit_mc2(aempty_bar(b));                // the window is free once the 9 taps have read it
                    if (kc == kchunks - 1) tc_commit_mc2(tfull_bar(acc));
                }
                __syncwarp();
                if (kc == kchunks - 1) {
                    if (++acc == 2) { acc = 0; aph ^= 1u; }
                }
            }
        }
    } else {
        const int quad = warp & 3;
        int acc = 0; uint32_t aph = 0;
        constexpr bool HAS_STATS = (FL & (EF_STATS | EF_BNBWD)) != 0;
        const int stat_per_cta = HAS_STATS && num_clusters % n_tiles == 0;              // this pair's tiles share one column block
        if (stat_per_cta) zero_cta_stats<BN>(csum);
        for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
            const int m2 = tile / n_tiles, n0 = (tile - m2 * n_tiles) * BN;
            mbar_wait(tfull_bar(acc), aph);
            tc_fence_after();
            epilogue_tile<BN, FL, true>(staging, csum, tmem_base + (uint32_t)(acc * BN), tempty_bar(acc), m2 * 2 + (int)rank, n0, rows, N,
                                        out, ep, quad, lane, stat_per_cta);
            if (++acc == 2) { acc = 0; aph ^= 1u; }
        }
        if (HAS_STATS)
            stats_tail<BN>(ep, staging, csum, stat_per_cta, cluster_id < total_tiles, (cluster_id / n_tiles) * 2 + (int)rank, (cluster_id % n_tiles) * BN, N,
                           stat_per_cta ? 2 * (num_clusters / n_tiles) : 2 * m2_tiles);
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// Single-CTA windowed variant for the 3x3 convs of stages 1-2 (N = 64 / 128) on large batches: 128 x BN tiles like conv_umma_kernel,
// one 184-row activation window per K chunk like conv2w_umma_kernel.  conv_umma_kernel re-loads the 128 A rows for each of the 9 taps,
// which makes these launches L2-bandwidth-bound (390 / 520 MB of L2 traffic per launch at ~10 TB/s); a CTA PAIR does not help them
// either, because an SS-mode cta_group::2 MMA with N <= 128 re-reads its A slice from shared memory faster than the 128 B/clk port
// delivers (measured).  cta_group::1 with N = 128 needs exactly 128 B/clk.  (Round 1 measured "no gain" for this
// scheme: at that time the issuing thread spent ~54 cycles per MMA, more than a 32-cycle N = 64 MMA takes -- see elect_one.)
// ------------------------------------------------------------------------------------------------
template <int TERMS, int BN_>
struct Conv1WCfg {
    static constexpr int BN = BN_;
    static constexpr int A_BYTES = WIN_ROWS * UM_BK * 2;      // one plane of the window (23 KB)
    static constexpr int W_BYTES = BN * UM_BK * 2;            // one plane of one tap's W tile
    static constexpr int A_PLANES = TERMS == 3 ? 2 : 1;
    static constexpr int W_PLANES = TERMS >= 2 ? 2 : 1;
    static constexpr int A_BUF_BYTES = A_PLANES * A_BYTES;
    static constexpr int W_STAGE_BYTES = W_PLANES * W_BYTES;
    static constexpr int W_BASE = 2 * A_BUF_BYTES;
    static constexpr int STAGING_BYTES = 4 * 32 * EPI_LD * 4;
    static constexpr int CSUM_BYTES = 5 * 2 * BN * 4;
    static constexpr int FIXED = STAGING_BYTES + CSUM_BYTES + 1024 + 256;
    static constexpr int W_FIT = (227 * 1024 - FIXED - W_BASE) / W_STAGE_BYTES;
    static constexpr int W_STAGES = W_FIT > 6 ? 6 : W_FIT;
    static constexpr int RING_BYTES = W_BASE + W_STAGES * W_STAGE_BYTES;
    static constexpr int SMEM_BYTES = RING_BYTES + FIXED;
    // N-STACKED terms: the hi and lo planes of a W tile are adjacent in shared memory with the same 8-row group stride, so ONE
    // descriptor with N = 2 BN multiplies the A slice with [W.hi ; W.lo] -- a.hi*w.hi lands in accumulator columns [0, BN),
    // a.hi*w.lo in [BN, 2 BN) -- and the epilogue adds the two halves.  Same three products as before in 2 instructions instead
    // of 3 (1 instead of 2 at TERMS = 2), and the 128-row A slice -- the operand a narrow-N SS-mode MMA is starved by, see
    // profiles/r2_ncu_source_conv1w_n64.md -- is read from shared memory twice per product instead of three times.
    static constexpr bool STACK = TERMS >= 2;
    static constexpr int ACC_COLS = STACK ? 2 * BN : BN;
    static constexpr int TMEM_COLS = 2 * ACC_COLS;            // double-buffered
};

template <int FL, int TERMS, int BN_>
__global__ void __launch_bounds__(UM_THREADS, 1)
conv1w_umma_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
                   const __grid_constant__ CUtensorMap mWhi, const __grid_constant__ CUtensorMap mWlo, int rows, int K,
                   int N, int m_tiles, int n_tiles, float* __restrict__ out, ConvEpilogue ep) {
    using Cfg = Conv1WCfg<TERMS, BN_>;
    constexpr int BN = Cfg::BN;
    static_assert(Cfg::A_BYTES % 1024 == 0 && Cfg::W_BYTES % 1024 == 0 && Cfg::W_STAGES >= 2, "planes must keep the 1024-byte swizzle alignment");
    static_assert(Cfg::TMEM_COLS <= 512, "two accumulators must fit TMEM");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    float* staging = reinterpret_cast<float*>(base_ptr + Cfg::RING_BYTES);
    float* csum = reinterpret_cast<float*>(base_ptr + Cfg::RING_BYTES + Cfg::STAGING_BYTES);
    const uint32_t bars = base + Cfg::RING_BYTES + Cfg::STAGING_BYTES + Cfg::CSUM_BYTES;
    auto afull_bar = [&](int b) { return bars + 8u * b; };
    auto aempty_bar = [&](int b) { return bars + 8u * (2 + b); };
    auto wfull_bar = [&](int s) { return bars + 8u * (4 + s); };
    auto wempty_bar = [&](int s) { return bars + 8u * (4 + Cfg::W_STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (4 + 2 * Cfg::W_STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (6 + 2 * Cfg::W_STAGES + a); };
    const uint32_t tmem_slot = bars + 8u * (8 + 2 * Cfg::W_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kchunks = K / UM_BK;
    const int total_tiles = m_tiles * n_tiles;
    const int my_tiles = (int)blockIdx.x < total_tiles ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int steps = my_tiles * kchunks;                  // a "step" g = one (tile, K chunk) = one A window

    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) { mbar_init(afull_bar(b), 1); mbar_init(aempty_bar(b), 1); }
        for (int s = 0; s < Cfg::W_STAGES; ++s) { mbar_init(wfull_bar(s), 1); mbar_init(wempty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&mAhi); tma_prefetch_desc(&mAlo); tma_prefetch_desc(&mWhi); tma_prefetch_desc(&mWlo);
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {
            auto load_window = [&](int g) {        // rows m0 - 26 .. m0 + 157 of K chunk kc (rows outside the tensor are zero-filled)
                const int tile = (int)blockIdx.x + (g / kchunks) * (int)gridDim.x, kc = g % kchunks;
                const int m0 = (tile / n_tiles) * UM_BM;
                const int b = g & 1;
                mbar_wait(aempty_bar(b), ((uint32_t)(g >> 1) & 1u) ^ 1u);
                mbar_expect_tx(afull_bar(b), Cfg::A_BUF_BYTES);
                const uint32_t sa = base + b * Cfg::A_BUF_BYTES;
                tma_load_2d(sa, &mAhi, afull_bar(b), kc * UM_BK, m0 - WIN_LEAD);
                if (TERMS == 3) tma_load_2d(sa + Cfg::A_BYTES, &mAlo, afull_bar(b), kc * UM_BK, m0 - WIN_LEAD);
            };
            int s = 0; uint32_t ph = 0;
            if (steps > 0) load_window(0);
            for (int g = 0; g < steps; ++g) {
                const int tile = (int)blockIdx.x + (g / kchunks) * (int)gridDim.x, kc = g % kchunks;
                const int n0 = (tile % n_tiles) * BN;
                for (int t = 0; t < 9; ++t) {
                    if (t == 3 && g + 1 < steps) load_window(g + 1);        // prefetch: its buffer was freed a full step ago
                    mbar_wait(wempty_bar(s), ph ^ 1u);
                    mbar_expect_tx(wfull_bar(s), Cfg::W_STAGE_BYTES);
                    const uint32_t sw = base + Cfg::W_BASE + s * Cfg::W_STAGE_BYTES;
                    tma_load_2d(sw, &mWhi, wfull_bar(s), kc * UM_BK, t * N + n0);
                    if (TERMS >= 2) tma_load_2d(sw + Cfg::W_BYTES, &mWlo, wfull_bar(s), kc * UM_BK, t * N + n0);
                    if (++s == Cfg::W_STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        {   // warp-wide loop, one elected lane issues (elect_one)
            constexpr uint32_t idesc = umma_idesc(UM_BM, BN, 0, 0);
            constexpr uint32_t idesc2 = umma_idesc(UM_BM, 2 * BN, 0, 0);      // [W.hi ; W.lo] as one B operand
            int s = 0; uint32_t ph = 0;
            int acc = 0; uint32_t aph = 0;
            for (int g = 0; g < steps; ++g) {
                const int kc = g % kchunks;
                if (kc == 0) {
                    mbar_wait(tempty_bar(acc), aph ^ 1u);
                    tc_fence_after();
                }
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * Cfg::ACC_COLS);
                const int b = g & 1;
                mbar_wait(afull_bar(b), (uint32_t)(g >> 1) & 1u);
                const uint32_t sa = base + b * Cfg::A_BUF_BYTES;
                for (int t = 0; t < 9; ++t) {
                    const int off = (t / 3 - 1) * PITCH + (t % 3 - 1);
                    const uint32_t arow = sa + (uint32_t)(WIN_LEAD + off) * 128u;
                    mbar_wait(wfull_bar(s), ph);
                    tc_fence_after();
                    const uint32_t sw = base + Cfg::W_BASE + s * Cfg::W_STAGE_BYTES;
                    const uint64_t a_hi0 = umma_desc(arow, 16, 1024), w_hi0 = umma_desc(sw, 16, 1024);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < UM_BK / 16; ++k) {
                            const uint64_t a_hi = a_hi0 + (uint64_t)(k * 2), w_hi = w_hi0 + (uint64_t)(k * 2);
                            const uint32_t accumulate = (kc | t | k) != 0;
                            if (TERMS == 3) {
                                const uint64_t a_lo = a_hi + (uint64_t)(Cfg::A_BYTES >> 4);
                                tc_mma_bf16(d_tmem, a_hi, w_hi, idesc2, accumulate);       // a.hi * [w.hi ; w.lo]: zeroes both halves on the first
                                tc_mma_bf16(d_tmem, a_lo, w_hi, idesc, 1);                 // a.lo * w.hi into the first half
                            } else if (TERMS == 2) {
                                tc_mma_bf16(d_tmem, a_hi, w_hi, idesc2, accumulate);
                            } else {
                                tc_mma_bf16(d_tmem, a_hi, w_hi, idesc, accumulate);
                            }
                        }
                        tc_commit(wempty_bar(s));
                        if (t == 8) {
                            tc_commit(aempty_bar(b));            // the window is free once the 9 taps have read it
                            if (kc == kchunks - 1) tc_commit(tfull_bar(acc));
                        }
                    }
                    __syncwarp();
                    if (++s == Cfg::W_STAGES) { s = 0; ph ^= 1u; }
                }
                if (kc == kchunks - 1) {
                    if (++acc == 2) { acc = 0; aph ^= 1u; }
                }
            }
        }
    } else {
        const int quad = warp & 3;
        int acc = 0; uint32_t aph = 0;
        constexpr bool HAS_STATS = (FL & (EF_STATS | EF_BNBWD)) != 0;
        const int stat_per_cta = HAS_STATS && gridDim.x % n_tiles == 0;                 // this CTA's tiles share one column block
        if (stat_per_cta) zero_cta_stats<BN>(csum);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int m_t = tile / n_tiles, n0 = (tile - m_t * n_tiles) * BN;
            mbar_wait(tfull_bar(acc), aph);
            tc_fence_after();
            epilogue_tile<BN, FL, false, Cfg::STACK>(staging, csum, tmem_base + (uint32_t)(acc * Cfg::ACC_COLS), tempty_bar(acc), m_t, n0, rows, N, out,
                                                     ep, quad, lane, stat_per_cta);
            if (++acc == 2) { acc = 0; aph ^= 1u; }
        }
        if (HAS_STATS)
            stats_tail<BN>(ep, staging, csum, stat_per_cta, (int)blockIdx.x < total_tiles, blockIdx.x / n_tiles, ((int)blockIdx.x % n_tiles) * BN, N,
                           stat_per_cta ? (int)gridDim.x / n_tiles : m_tiles);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// wgrad: partial[z][co][ci] = sum_{p in split} dY[p][co] * X[p + off_t][ci],  z = split*ntaps + t
// ------------------------------------------------------------------------------------------------
template <int BN, int TERMS = 3>
struct WgradCfg {
    static constexpr int A_BYTES = UM_BM * UM_BK * 2;         // dY plane: 2 blocks of [64 rows][64 co]
    static constexpr int B_BYTES = BN * UM_BK * 2;            // X plane: BN/64 blocks of [64 rows][64 ci]
    static constexpr int A_PLANES = TERMS == 3 ? 2 : 1;       // TERMS = 2: dY contributes its hi plane only
    static constexpr int B_PLANES = TERMS >= 2 ? 2 : 1;
    static constexpr int B_OFF = A_PLANES * A_BYTES;          // stage layout: dY hi [, dY lo], X hi [, X lo]
    static constexpr int STAGE_BYTES = A_PLANES * A_BYTES + B_PLANES * B_BYTES;
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 6 ? 6 : (200 * 1024) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
    // single-CTA kernel only: dY.hi * [X.hi | X.lo] as ONE N = 2 BN instruction (the lo plane's 64-channel blocks follow the hi
    // plane's at the same 8 KB stride); the epilogue adds the two accumulator halves.  See Conv1WCfg::STACK.
    static constexpr bool STACK = TERMS >= 2 && BN <= 128;
    static constexpr int ACC_COLS = STACK ? 2 * BN : BN;
};

template <int BN, int TERMS>
__global__ void __launch_bounds__(UM_THREADS, 1)
wgrad_umma_kernel(const __grid_constant__ CUtensorMap mYhi, const __grid_constant__ CUtensorMap mYlo,
                  const __grid_constant__ CUtensorMap mXhi, const __grid_constant__ CUtensorMap mXlo, long long rows, int Cout,
                  int Cin, int ntaps, int nsplit, long long chunk, float* __restrict__ partial) {
    using Cfg = WgradCfg<BN, TERMS>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + Cfg::STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (Cfg::STAGES + s); };
    const uint32_t accum_bar = bars + 8u * (2 * Cfg::STAGES);
    const uint32_t tmem_slot = bars + 8u * (2 * Cfg::STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int co0 = blockIdx.x * UM_BM, ci0 = blockIdx.y * BN;
    const int t = blockIdx.z % ntaps, sp = blockIdx.z / ntaps;
    const int off = ntaps == 9 ? (t / 3 - 1) * PITCH + (t % 3 - 1) : 0;
    const long long p_begin = (long long)sp * chunk;
    long long p_end = p_begin + chunk;
    if (p_end > rows) p_end = rows;
    const int iters = p_end > p_begin ? (int)((p_end - p_begin + UM_BK - 1) / UM_BK) : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&mYhi); tma_prefetch_desc(&mYlo); tma_prefetch_desc(&mXhi); tma_prefetch_desc(&mXlo);
    }
    if (warp == 1) tmem_alloc<tmem_cols(Cfg::ACC_COLS)>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
                const int p0 = (int)(p_begin + (long long)it * UM_BK);
                mbar_wait(empty_bar(s), ph ^ 1u);
                mbar_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
                const uint32_t sa = base + s * Cfg::STAGE_BYTES;
#pragma unroll
                for (int j = 0; j < UM_BM / 64; ++j) {
                    tma_load_2d(sa + j * 8192, &mYhi, full_bar(s), co0 + j * 64, p0);
                    if (TERMS == 3) tma_load_2d(sa + Cfg::A_BYTES + j * 8192, &mYlo, full_bar(s), co0 + j * 64, p0);
                }
#pragma unroll
                for (int j = 0; j < BN / 64; ++j) {
                    tma_load_2d(sa + Cfg::B_OFF + j * 8192, &mXhi, full_bar(s), ci0 + j * 64, p0 + off);
                    if (TERMS >= 2) tma_load_2d(sa + Cfg::B_OFF + Cfg::B_BYTES + j * 8192, &mXlo, full_bar(s), ci0 + j * 64, p0 + off);
                }
                if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        {   // warp-wide loop, one elected lane issues (elect_one)
            constexpr uint32_t idesc = umma_idesc(UM_BM, BN, 1, 1);
            constexpr uint32_t idesc2 = umma_idesc(UM_BM, 2 * BN <= 256 ? 2 * BN : BN, 1, 1);      // [X.hi | X.lo] as one B operand
            int s = 0; uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t sa = base + s * Cfg::STAGE_BYTES;
                // MN-major SWIZZLE_128B: 8 rows of 128 B per atom (SBO = 1024 B between 8-row groups along
                // K), 64-channel blocks 8 KB apart (LBO); a K=16 step is two atoms = 2048 B (= +128 in the start-address field).
                const uint64_t y_hi0 = umma_desc(sa, 8192, 1024), x_hi0 = umma_desc(sa + Cfg::B_OFF, 8192, 1024);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < UM_BK / 16; ++k) {
                        const uint64_t y_hi = y_hi0 + (uint64_t)(k * 128), x_hi = x_hi0 + (uint64_t)(k * 128);
                        if (TERMS == 3 && Cfg::STACK) {
                            const uint64_t y_lo = y_hi + (uint64_t)(Cfg::A_BYTES >> 4);
                            tc_mma_bf16(tmem_base, y_hi, x_hi, idesc2, (it | k) != 0);     // the first one zeroes both halves
                            tc_mma_bf16(tmem_base, y_lo, x_hi, idesc, 1);
                        } else if (TERMS == 2 && Cfg::STACK) {
                            tc_mma_bf16(tmem_base, y_hi, x_hi, idesc2, (it | k) != 0);
                        } else if (TERMS == 3) {
                            const uint64_t y_lo = y_hi + (uint64_t)(Cfg::A_BYTES >> 4), x_lo = x_hi + (uint64_t)(Cfg::B_BYTES >> 4);
                            tc_mma_bf16(tmem_base, y_lo, x_hi, idesc, (it | k) != 0);
                            tc_mma_bf16(tmem_base, y_hi, x_lo, idesc, 1);
                            tc_mma_bf16(tmem_base, y_hi, x_hi, idesc, 1);
                        } else if (TERMS == 2) {
                            const uint64_t x_lo = x_hi + (uint64_t)(Cfg::B_BYTES >> 4);
                            tc_mma_bf16(tmem_base, y_hi, x_lo, idesc, (it | k) != 0);
                            tc_mma_bf16(tmem_base, y_hi, x_hi, idesc, 1);
                        } else {
                            tc_mma_bf16(tmem_base, y_hi, x_hi, idesc, (it | k) != 0);
                        }
                    }
                    tc_commit(empty_bar(s));
                    if (it == iters - 1) tc_commit(accum_bar);
                }
                __syncwarp();
                if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        const int quad = warp & 3;
        const int co = co0 + quad * 32 + lane;
        float* dst = partial + ((size_t)blockIdx.z * Cout + co) * Cin + ci0;
        if (iters > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
        }
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            float v[32];
            if (iters > 0) {
                if (Cfg::STACK) tmem_ld32_sum2(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, (uint32_t)BN, v);
                else tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, v);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            if (co < Cout && ci0 + c < Cin) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(dst + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<tmem_cols(Cfg::ACC_COLS)>(tmem_base);
    }
}

// CTA-pair wgrad (Cout % 256 == 0 and Cin % 256 == 0): the pair owns a 256 (co) x 256 (ci) tile of one tap / row split;
// each CTA stages its own 128 dY channels and half (128) of the X channels, the leader issues M = 256, N = 256 MMAs.
template <int TERMS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(UM_THREADS, 1)
wgrad2_umma_kernel(const __grid_constant__ CUtensorMap mYhi, const __grid_constant__ CUtensorMap mYlo,
                   const __grid_constant__ CUtensorMap mXhi, const __grid_constant__ CUtensorMap mXlo, long long rows, int Cout,
                   int Cin, int ntaps, int nsplit, long long chunk, float* __restrict__ partial) {
    using Cfg = WgradCfg<128, TERMS>;              // per-CTA stage: 128 co + 128 ci (hi and lo: 64 KB)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + Cfg::STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (Cfg::STAGES + s); };
    const uint32_t accum_bar = bars + 8u * (2 * Cfg::STAGES);
    const uint32_t tmem_slot = bars + 8u * (2 * Cfg::STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int co0 = (blockIdx.x >> 1) * 256 + (int)rank * 128;      // this CTA's dY channels (= its accumulator rows)
    const int ci_pair0 = blockIdx.y * 256;                          // the pair's X channels
    const int ci0 = ci_pair0 + (int)rank * 128;                     // the half this CTA stages
    const int t = blockIdx.z % ntaps, sp = blockIdx.z / ntaps;
    const int off = ntaps == 9 ? (t / 3 - 1) * PITCH + (t % 3 - 1) : 0;
    const long long p_begin = (long long)sp * chunk;
    long long p_end = p_begin + chunk;
    if (p_end > rows) p_end = rows;
    const int iters = p_end > p_begin ? (int)((p_end - p_begin + UM_BK - 1) / UM_BK) : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&mYhi); tma_prefetch_desc(&mYlo); tma_prefetch_desc(&mXhi); tma_prefetch_desc(&mXlo);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
                const int p0 = (int)(p_begin + (long long)it * UM_BK);
                mbar_wait(empty_bar(s), ph ^ 1u);
                if (rank == 0) mbar_expect_tx(full_bar(s), 2 * Cfg::STAGE_BYTES);
                const uint32_t fb = mapa_rank0(full_bar(s));
                const uint32_t sa = base + s * Cfg::STAGE_BYTES;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    tma_load_2d_cg2(sa + j * 8192, &mYhi, fb, co0 + j * 64, p0);
                    tma_load_2d_cg2(sa + Cfg::B_OFF + j * 8192, &mXhi, fb, ci0 + j * 64, p0 + off);
                    if (TERMS == 3) tma_load_2d_cg2(sa + Cfg::A_BYTES + j * 8192, &mYlo, fb, co0 + j * 64, p0);
                    if (TERMS >= 2) tma_load_2d_cg2(sa + Cfg::B_OFF + Cfg::B_BYTES + j * 8192, &mXlo, fb, ci0 + j * 64, p0 + off);
                }
                if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {            // warp-wide loop, one elected lane issues (elect_one)
            constexpr uint32_t idesc = umma_idesc(256, 256, 1, 1);
            int s = 0; uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t sa = base + s * Cfg::STAGE_BYTES;
                const uint64_t y_hi0 = umma_desc(sa, 8192, 1024), x_hi0 = umma_desc(sa + Cfg::B_OFF, 8192, 1024);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < UM_BK / 16; ++k) {
                        const uint64_t y_hi = y_hi0 + (uint64_t)(k * 128), x_hi = x_hi0 + (uint64_t)(k * 128);
                        if (TERMS == 3) {
                            const uint64_t y_lo = y_hi + (uint64_t)(Cfg::A_BYTES >> 4), x_lo = x_hi + (uint64_t)(Cfg::B_BYTES >> 4);
                            tc_mma_bf16_cg2(tmem_base, y_lo, x_hi, idesc, (it | k) != 0);
                            tc_mma_bf16_cg2(tmem_base, y_hi, x_lo, idesc, 1);
                            tc_mma_bf16_cg2(tmem_base, y_hi, x_hi, idesc, 1);
                        } else if (TERMS == 2) {
                            const uint64_t x_lo = x_hi + (uint64_t)(Cfg::B_BYTES >> 4);
                            tc_mma_bf16_cg2(tmem_base, y_hi, x_lo, idesc, (it | k) != 0);
                            tc_mma_bf16_cg2(tmem_base, y_hi, x_hi, idesc, 1);
                        } else {
                            tc_mma_bf16_cg2(tmem_base, y_hi, x_hi, idesc, (it | k) != 0);
                        }
                    }
                    tc_commit_mc2(empty_bar(s));
                    if (it == iters - 1) tc_commit_mc2(accum_bar);
                }
                __syncwarp();
                if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        const int quad = warp & 3;
        const int co = co0 + quad * 32 + lane;
        float* dst = partial + ((size_t)blockIdx.z * Cout + co) * Cin + ci_pair0;
        if (iters > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
        }
#pragma unroll 1
        for (int c = 0; c < 256; c += 32) {
            float v[32];
            if (iters > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, v);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(dst + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
        tc_fence_before();
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
    }
}

// dW[co][ci][t] (OIHW) = sum_sp partial[sp*ntaps + t][co][ci]
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int Cout, int Cin, int ntaps,
                                                           int nsplit, float* __restrict__ dW) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long per = (long long)Cout * Cin;
    if (idx >= per * ntaps) return;
    int t = (int)(idx / per);
    long long r = idx - (long long)t * per;          // co*Cin + ci
    float acc = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) acc += partial[((size_t)sp * ntaps + t) * per + r];
    dW[(size_t)r * ntaps + t] = acc;
}

int k_wgrad_reduce(const float* partial, int Cout, int Cin, int ntaps, int nsplit, float* dW, cudaStream_t s) {
    long long n = (long long)Cout * Cin * ntaps;
    wgrad_reduce_kernel<<<ceil_div(n, 256), 256, 0, s>>>(partial, Cout, Cin, ntaps, nsplit, dW);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// split-K tail: out = epilogue(sum_z partial[z]) -- the same transforms as epilogue_tile, elementwise
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, int nz, int rows, int N, ConvEpilogue ep,
                                                            float* __restrict__ out) {
    const int N4 = N >> 2;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * N4) return;
    const int m = (int)(idx / N4), n = (int)(idx - (long long)m * N4) * 4;
    const size_t o = (size_t)m * N + n, plane = (size_t)rows * N;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!(ep.pitch25 && !p25_valid(m % IMG25))) {
        for (int z = 0; z < nz; ++z) {
            const float4 p = *reinterpret_cast<const float4*>(partial + z * plane + o);
            x.x += p.x; x.y += p.y; x.z += p.z; x.w += p.w;
        }
        if (ep.scale) {
            const float4 sc = *reinterpret_cast<const float4*>(ep.scale + n), sh = *reinterpret_cast<const float4*>(ep.shift + n);
            x.x = fmaf(x.x, sc.x, sh.x); x.y = fmaf(x.y, sc.y, sh.y); x.z = fmaf(x.z, sc.z, sh.z); x.w = fmaf(x.w, sc.w, sh.w);
        }
        if (ep.add_prev) { const float4 p = *reinterpret_cast<const float4*>(ep.add_prev + o); x.x += p.x; x.y += p.y; x.z += p.z; x.w += p.w; }
        if (ep.res.hi) {
            float r[4];
            const uint2 h = *reinterpret_cast<const uint2*>(ep.res.hi + o), l = *reinterpret_cast<const uint2*>(ep.res.lo + o);
            const bf16* hb = reinterpret_cast<const bf16*>(&h); const bf16* lb = reinterpret_cast<const bf16*>(&l);
#pragma unroll
            for (int i = 0; i < 4; ++i) r[i] = bf2f(hb[i]) + bf2f(lb[i]);
            x.x += r[0]; x.y += r[1]; x.z += r[2]; x.w += r[3];
        }
        if (ep.relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
    }
    if (out) *reinterpret_cast<float4*>(out + o) = x;
    if (ep.out_split.hi) {
        bf16 h[4], l[4];
        split_store(x.x, h[0], l[0]); split_store(x.y, h[1], l[1]); split_store(x.z, h[2], l[2]); split_store(x.w, h[3], l[3]);
        *reinterpret_cast<uint2*>(ep.out_split.hi + o) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(ep.out_split.lo + o) = *reinterpret_cast<const uint2*>(l);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

int umma_init() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SIMQ_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) { simq_set_error("cuTensorMapEncodeTiled not available"); return 1; }
    g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    return 0;
}

// bf16 [rows][cols] row-major, box [box_rows][64 cols], SWIZZLE_128B, zero fill out of bounds
static int make_map(CUtensorMap* m, const bf16* ptr, long long rows, int cols, int box_rows, int box_cols = 64) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ptr, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { simq_set_error("cuTensorMapEncodeTiled failed: %d (rows=%lld cols=%d box=%d)", (int)r, rows, cols, box_rows); return 1; }
    return 0;
}

// SM count of the current device (cached per device; concurrent first calls write the same value)
static int g_sm_count[64] = {0};
static int num_sms() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = g_sm_count[dev];
    if (!n) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_sm_count[dev] = n;
    }
    return n;
}

template <int BN, int FL, int TERMS>
static int launch_conv(const UmmaTensor& A, const UmmaTensor& W, int N, int ntaps, float* out, ConvEpilogue ep, cudaStream_t s, int nz = 1) {
    using Cfg = ConvCfg<BN, TERMS>;
    static unsigned long long attr = 0;      // per-device: function attributes belong to the device context
    if (first_use_on_device(attr)) SIMQ_CUDA(cudaFuncSetAttribute(conv_umma_kernel<BN, FL, TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    const int g_num_sms = num_sms();
    if (A.rows >= (1LL << 31) - 256) { simq_set_error("k_conv_umma: too many rows"); return 1; }
    CUtensorMap mAhi, mAlo, mWhi, mWlo;
    if (make_map(&mAhi, A.t.hi, A.rows, A.cols, UM_BM, Cfg::BK) || make_map(&mAlo, A.t.lo, A.rows, A.cols, UM_BM, Cfg::BK) ||
        make_map(&mWhi, W.t.hi, W.rows, W.cols, BN, Cfg::BK) || make_map(&mWlo, W.t.lo, W.rows, W.cols, BN, Cfg::BK))
        return 1;
    const int m_tiles = ceil_div(A.rows, UM_BM), n_tiles = N / BN;
    const int grid = m_tiles * n_tiles * nz < g_num_sms ? m_tiles * n_tiles * nz : g_num_sms;
    if (ep.stat_rows_out) *ep.stat_rows_out = (nz == 1 && grid % n_tiles == 0) ? grid / n_tiles : m_tiles;     // see epilogue_tile
    // algorithmic FLOPs: 2 * valid output positions * N * K * taps (pitch-25 rows carry 576 of 625 valid)
    const double valid_rows = ep.pitch25 ? (double)A.rows * 576.0 / 625.0 : (double)A.rows;
    prof_mark(PROF_CONV, true, 2.0 * valid_rows * N * A.cols * ntaps, s, 2.0 * TERMS * (double)A.rows * N * A.cols * ntaps);
    conv_umma_kernel<BN, FL, TERMS><<<grid, UM_THREADS, Cfg::SMEM_BYTES, s>>>(mAhi, mAlo, mWhi, mWlo, (int)A.rows, A.cols, N, ntaps, m_tiles, n_tiles, nz, out, ep);
    prof_mark(PROF_CONV, false, 0, s);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

template <int FL, int TERMS>
static int launch_conv2(const UmmaTensor& A, const UmmaTensor& W, int N, int ntaps, float* out, ConvEpilogue ep, cudaStream_t s) {
    using Cfg = Conv2Cfg<TERMS>;
    static unsigned long long attr = 0;      // per-device: function attributes belong to the device context
    if (first_use_on_device(attr)) SIMQ_CUDA(cudaFuncSetAttribute(conv2_umma_kernel<FL, TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    const int g_num_sms = num_sms();
    CUtensorMap mAhi, mAlo, mWhi, mWlo;
    if (make_map(&mAhi, A.t.hi, A.rows, A.cols, UM_BM) || make_map(&mAlo, A.t.lo, A.rows, A.cols, UM_BM) ||
        make_map(&mWhi, W.t.hi, W.rows, W.cols, Cfg::BN / 2) || make_map(&mWlo, W.t.lo, W.rows, W.cols, Cfg::BN / 2))
        return 1;
    const int m2_tiles = ceil_div(A.rows, 256), n_tiles = N / Cfg::BN;
    int clusters = g_num_sms / 2;
    if (m2_tiles * n_tiles < clusters) clusters = m2_tiles * n_tiles;
    if (ep.stat_rows_out) *ep.stat_rows_out = clusters % n_tiles == 0 ? 2 * (clusters / n_tiles) : 2 * m2_tiles;    // see epilogue_tile
    const double valid_rows = ep.pitch25 ? (double)A.rows * 576.0 / 625.0 : (double)A.rows;
    prof_mark(PROF_CONV, true, 2.0 * valid_rows * N * A.cols * ntaps, s, 2.0 * TERMS * (double)A.rows * N * A.cols * ntaps);
    conv2_umma_kernel<FL, TERMS><<<2 * clusters, UM_THREADS, Cfg::SMEM_BYTES, s>>>(mAhi, mAlo, mWhi, mWlo, (int)A.rows, A.cols, N, ntaps, m2_tiles,
                                                                        n_tiles, out, ep);
    prof_mark(PROF_CONV, false, 0, s);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// windowed pair kernel (3x3 convs only): the A maps carry a 184-row box
static int conv_window_mode() {      // SIMQ_CONV_WINDOW=0 falls back to the per-tap A loads (A/B experiments)
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("SIMQ_CONV_WINDOW"); mode = e ? atoi(e) : 1; }
    return mode;
}
template <int FL, int TERMS, int BN = 256>
static int launch_conv2w(const UmmaTensor& A, const UmmaTensor& W, int N, float* out, ConvEpilogue ep, cudaStream_t s) {
    using Cfg = Conv2WCfg<TERMS, BN>;
    static unsigned long long attr = 0;      // per-device: function attributes belong to the device context
    if (first_use_on_device(attr)) SIMQ_CUDA(cudaFuncSetAttribute(conv2w_umma_kernel<FL, TERMS, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    const int g_num_sms = num_sms();
    CUtensorMap mAhi, mAlo, mWhi, mWlo;
    if (make_map(&mAhi, A.t.hi, A.rows, A.cols, WIN_ROWS) || make_map(&mAlo, A.t.lo, A.rows, A.cols, WIN_ROWS) ||
        make_map(&mWhi, W.t.hi, W.rows, W.cols, Cfg::BN / 2) || make_map(&mWlo, W.t.lo, W.rows, W.cols, Cfg::BN / 2))
        return 1;
    const int m2_tiles = ceil_div(A.rows, 256), n_tiles = N / Cfg::BN;
    int clusters = g_num_sms / 2;
    if (m2_tiles * n_tiles < clusters) clusters = m2_tiles * n_tiles;
    if (ep.stat_rows_out) *ep.stat_rows_out = clusters % n_tiles == 0 ? 2 * (clusters / n_tiles) : 2 * m2_tiles;    // see epilogue_tile
    const double valid_rows = ep.pitch25 ? (double)A.rows * 576.0 / 625.0 : (double)A.rows;
    prof_mark(PROF_CONV, true, 2.0 * valid_rows * N * A.cols * 9, s, 2.0 * TERMS * (double)A.rows * N * A.cols * 9);
    conv2w_umma_kernel<FL, TERMS, BN><<<2 * clusters, UM_THREADS, Cfg::SMEM_BYTES, s>>>(mAhi, mAlo, mWhi, mWlo, (int)A.rows, A.cols, N, m2_tiles, n_tiles,
                                                                         out, ep);
    prof_mark(PROF_CONV, false, 0, s);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

static int small_window_mode() {     // SIMQ_CONV_WINDOW_SMALL=0: stages 1-2 keep the per-tap kernel (A/B experiments)
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("SIMQ_CONV_WINDOW_SMALL"); mode = e ? atoi(e) : 1; }
    return mode;
}
template <int FL, int TERMS, int BN>
static int launch_conv1w(const UmmaTensor& A, const UmmaTensor& W, int N, float* out, ConvEpilogue ep, cudaStream_t s) {
    using Cfg = Conv1WCfg<TERMS, BN>;
    static unsigned long long attr = 0;      // per-device: function attributes belong to the device context
    if (first_use_on_device(attr)) SIMQ_CUDA(cudaFuncSetAttribute(conv1w_umma_kernel<FL, TERMS, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    const int g_num_sms = num_sms();
    if (A.rows >= (1LL << 31) - 256) { simq_set_error("k_conv_umma: too many rows"); return 1; }
    CUtensorMap mAhi, mAlo, mWhi, mWlo;
    if (make_map(&mAhi, A.t.hi, A.rows, A.cols, WIN_ROWS) || make_map(&mAlo, A.t.lo, A.rows, A.cols, WIN_ROWS) ||
        make_map(&mWhi, W.t.hi, W.rows, W.cols, BN) || make_map(&mWlo, W.t.lo, W.rows, W.cols, BN))
        return 1;
    const int m_tiles = ceil_div(A.rows, UM_BM), n_tiles = N / BN;
    const int grid = m_tiles * n_tiles < g_num_sms ? m_tiles * n_tiles : g_num_sms;
    if (ep.stat_rows_out) *ep.stat_rows_out = grid % n_tiles == 0 ? grid / n_tiles : m_tiles;      // see epilogue_tile
    const double valid_rows = ep.pitch25 ? (double)A.rows * 576.0 / 625.0 : (double)A.rows;
    prof_mark(PROF_CONV, true, 2.0 * valid_rows * N * A.cols * 9, s, 2.0 * TERMS * (double)A.rows * N * A.cols * 9);
    conv1w_umma_kernel<FL, TERMS, BN><<<grid, UM_THREADS, Cfg::SMEM_BYTES, s>>>(mAhi, mAlo, mWhi, mWlo, (int)A.rows, A.cols, N, m_tiles, n_tiles, out, ep);
    prof_mark(PROF_CONV, false, 0, s);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

// the epilogue variants the network uses
// PAIR: 0 = single-CTA 128 x BN tiles; 1 = CTA-pair 256 x 256 tiles (windowed for 3x3); 2 = single-CTA 128 x BN tiles with the
// activation window (3x3 only, large batches)
template <int BN, int PAIR, int F, int TERMS>
static int conv_launch_t(const UmmaTensor& A, const UmmaTensor& W, int N, int ntaps, float* out, ConvEpilogue ep, cudaStream_t s) {
    if constexpr (PAIR == 1) {
        if (ntaps == 9 && conv_window_mode() != 0) return launch_conv2w<F, TERMS>(A, W, N, out, ep, s);
        return launch_conv2<F, TERMS>(A, W, N, ntaps, out, ep, s);
    } else if constexpr (PAIR == 2) {
        return launch_conv1w<F, TERMS, BN>(A, W, N, out, ep, s);
    } else {
        return launch_conv<BN, F, TERMS>(A, W, N, ntaps, out, ep, s);
    }
}
template <int BN, int PAIR, int F>
static int conv_by_terms(const UmmaTensor& A, const UmmaTensor& W, int N, int ntaps, float* out, ConvEpilogue ep, cudaStream_t s) {
    // the two-term variant exists for the epilogues dgrad launches use (raw fp32 output, +previous gradient, +masked identity
    // gradient, +BatchNorm-backward sums): the forward never runs it
    constexpr bool DGRAD_EPILOGUE = (F & ~(EF_PREV | EF_G | EF_BNBWD)) == EF_F32;
    if (ep.terms == 1) return conv_launch_t<BN, PAIR, F, 1>(A, W, N, ntaps, out, ep, s);
    if (ep.terms == 2) {
        if constexpr (DGRAD_EPILOGUE) return conv_launch_t<BN, PAIR, F, 2>(A, W, N, ntaps, out, ep, s);
        simq_set_error("k_conv_umma: two-term operands are a backward-only mode (epilogue 0x%x)", F);
        return 1;
    }
    return conv_launch_t<BN, PAIR, F, 3>(A, W, N, ntaps, out, ep, s);
}
template <int BN, int PAIR = 0>
static int dispatch_conv(const UmmaTensor& A, const UmmaTensor& W, int N, int ntaps, float* out, ConvEpilogue ep, cudaStream_t s) {
    int fl = 0;
    if (ep.stats && !ep.bn_raw) fl |= EF_STATS;
    if (ep.scale) fl |= EF_AFFINE;
    if (ep.add_prev) fl |= EF_PREV;
    if (ep.res.hi) fl |= EF_RES;
    if (ep.add_g) fl |= EF_G;
    if (ep.relu) fl |= EF_RELU;
    if (out) fl |= EF_F32;
    if (ep.out_split.hi) fl |= EF_SPLIT;
    if (ep.bn_raw) fl |= EF_BNBWD;
    switch (fl) {
#define CASE(F) case (F): return conv_by_terms<BN, PAIR, (F)>(A, W, N, ntaps, out, ep, s);
        CASE(EF_F32);                                               // raw conv output (dgrad, eval head)
        CASE(EF_F32 | EF_STATS);                                    // train forward: raw + BN statistics
        CASE(EF_F32 | EF_PREV);                                     // dgrad accumulate (downsample branch)
        CASE(EF_F32 | EF_G);                                        // dgrad + masked identity gradient
        CASE(EF_F32 | EF_BNBWD);                                    // dgrad + BN-backward sums of the consumer BN
        CASE(EF_F32 | EF_PREV | EF_BNBWD);
        CASE(EF_F32 | EF_G | EF_BNBWD);
        CASE(EF_F32 | EF_AFFINE);                                   // eval downsample: BN'd identity
        CASE(EF_SPLIT | EF_AFFINE | EF_RELU);                       // eval conv1 -> BN -> ReLU
        CASE(EF_SPLIT | EF_AFFINE | EF_RELU | EF_RES);              // eval conv2 -> BN -> + identity -> ReLU
        CASE(EF_SPLIT | EF_AFFINE | EF_RELU | EF_PREV);             // eval conv2 -> BN -> + downsample -> ReLU
#undef CASE
        default: simq_set_error("k_conv_umma: epilogue combination 0x%x not instantiated", fl); return 1;
    }
}

int umma_conv_m_tiles(long long rows) { return ceil_div(rows, UM_BM); }
bool umma_conv_supported(int K, int N) { return K % UM_BK == 0 && (N == 32 || N % 64 == 0); }
bool umma_wgrad_supported(int Cout, int Cin) { return Cout % 64 == 0 && Cin % 64 == 0; }

// Split-K over the taps (small problems only: the 128 x BN tiles would leave most SMs idle, e.g. 5 m-tiles at batch 1):
// raw partials of nz tap groups into `scratch`, then one elementwise pass applies the epilogue.
template <int BN>
static int conv_splitk(const UmmaTensor& A, const UmmaTensor& W, int N, int ntaps, float* out, ConvEpilogue ep, int nz, float* scratch,
                       cudaStream_t s) {
    ConvEpilogue raw = conv_ep(0);
    raw.terms = ep.terms;
    int rc = ep.terms == 1 ? launch_conv<BN, EF_F32, 1>(A, W, N, ntaps, scratch, raw, s, nz)
           : ep.terms == 2 ? launch_conv<BN, EF_F32, 2>(A, W, N, ntaps, scratch, raw, s, nz)
                           : launch_conv<BN, EF_F32, 3>(A, W, N, ntaps, scratch, raw, s, nz);
    if (rc) return rc;
    const long long n = A.rows * (N / 4);
    splitk_reduce_kernel<<<ceil_div(n, 256), 256, 0, s>>>(scratch, nz, (int)A.rows, N, ep, out);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

int k_conv_umma(const UmmaTensor& A, const UmmaTensor& W, int N, int ntaps, float* out, ConvEpilogue ep, cudaStream_t s) {
    if (umma_init()) return 1;
    if (!umma_conv_supported(A.cols, N) || W.cols != A.cols || W.rows != (long long)ntaps * N) {
        simq_set_error("k_conv_umma: unsupported shape K=%d N=%d", A.cols, N);
        return 1;
    }
    if (N == 32) return dispatch_conv<32>(A, W, N, ntaps, out, ep, s);
    if (ntaps == 9 && !ep.stats && !ep.bn_raw && !ep.add_g && ep.splitk_scratch) {
        const int g_num_sms = num_sms();
        const int bn = N % 128 == 0 ? 128 : 64;
        const long long tiles = (long long)ceil_div(A.rows, 128) * (N / bn);
        const int nz = tiles * 9 <= 2 * g_num_sms ? 9 : tiles * 3 <= g_num_sms ? 3 : 1;
        if (nz > 1 && (size_t)nz * A.rows * N <= ep.splitk_floats)
            return bn == 128 ? conv_splitk<128>(A, W, N, ntaps, out, ep, nz, ep.splitk_scratch, s)
                             : conv_splitk<64>(A, W, N, ntaps, out, ep, nz, ep.splitk_scratch, s);
    }
    // tile policy for N % 256 == 0: the CTA-pair kernel (256 x 256 per pair) runs ~12 % more tensor work per cycle
    // than 128 x 128 single-CTA tiles but quantises worse on small problems; pick the cheaper estimate in units
    // of one 128 x 128 tile time.  SIMQ_CONV_TILE = 128 | pair overrides (experiments).
    static int policy = -1;
    if (policy < 0) {
        const char* e = getenv("SIMQ_CONV_TILE");
        policy = !e ? 0 : !strcmp(e, "128") ? 1 : !strcmp(e, "pair") ? 3 : 0;
    }
    if (N % 256 == 0 && policy != 1) {
        const int g_num_sms = num_sms();
        const long long t128 = (long long)ceil_div(A.rows, 128) * (N / 128), t256 = (long long)ceil_div(A.rows, 256) * (N / 256);
        const double cost_single = (double)((t128 + g_num_sms - 1) / g_num_sms);
        const double cost_pair = (double)((t256 + g_num_sms / 2 - 1) / (g_num_sms / 2)) * 2.0 * 0.88;
        if (policy == 3 || (policy == 0 && cost_pair < cost_single)) return dispatch_conv<128, 1>(A, W, N, ntaps, out, ep, s);
    }
    // (A windowed PAIR kernel with 256 x 128 / 256 x 64 tiles was built for these layers first and dropped: an SS-mode cta_group::2 MMA
    // with N <= 128 re-reads its A slice from shared memory faster than the port delivers.  DESIGN.md section 9.)
    // stages 1-2 (N = 64 / 128), 3x3, at least one tile per SM: the single-CTA kernel with the activation window
    if (ntaps == 9 && (N == 64 || N == 128) && policy != 1 && conv_window_mode() != 0 && small_window_mode() != 0 &&
        ceil_div(A.rows, UM_BM) >= num_sms())
        return N == 128 ? dispatch_conv<128, 2>(A, W, N, ntaps, out, ep, s) : dispatch_conv<64, 2>(A, W, N, ntaps, out, ep, s);
    if (N % 128 == 0) return dispatch_conv<128>(A, W, N, ntaps, out, ep, s);
    return dispatch_conv<64>(A, W, N, ntaps, out, ep, s);
}

size_t umma_wgrad_scratch_floats() { return (size_t)16 << 20; }     // 64 MB

template <int BN, int TERMS>
static int launch_wgrad(const UmmaTensor& dY, const UmmaTensor& X, int ntaps, float* dW, float* scratch, cudaStream_t s) {
    using Cfg = WgradCfg<BN, TERMS>;
    static unsigned long long attr = 0;      // per-device: function attributes belong to the device context
    if (first_use_on_device(attr)) SIMQ_CUDA(cudaFuncSetAttribute(wgrad_umma_kernel<BN, TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    const int Cout = dY.cols, Cin = X.cols;
    const long long rows = dY.rows;
    CUtensorMap mYhi, mYlo, mXhi, mXlo;
    if (make_map(&mYhi, dY.t.hi, rows, Cout, UM_BK) || make_map(&mYlo, dY.t.lo, rows, Cout, UM_BK) ||
        make_map(&mXhi, X.t.hi, rows, Cin, UM_BK) || make_map(&mXlo, X.t.lo, rows, Cin, UM_BK))
        return 1;
    const int tiles = ceil_div(Cout, UM_BM) * (Cin / BN) * ntaps;
    int nsplit = (444 + tiles - 1) / tiles;                               // ~3 waves of 148 SMs
    long long max_split = rows / (8 * UM_BK);
    if (max_split < 1) max_split = 1;
    if (nsplit > max_split) nsplit = (int)max_split;
    const size_t per_split = (size_t)ntaps * Cout * Cin;
    while (nsplit > 1 && per_split * nsplit > umma_wgrad_scratch_floats()) --nsplit;
    if (per_split * nsplit > umma_wgrad_scratch_floats()) { simq_set_error("wgrad scratch too small"); return 1; }
    long long chunk = ((rows + nsplit - 1) / nsplit + UM_BK - 1) / UM_BK * UM_BK;
    dim3 grid(ceil_div(Cout, UM_BM), Cin / BN, ntaps * nsplit);
    prof_mark(PROF_WGRAD, true, 2.0 * (double)rows * 576.0 / 625.0 * Cout * Cin * ntaps, s, 2.0 * TERMS * (double)rows * Cout * Cin * ntaps);
    wgrad_umma_kernel<BN, TERMS><<<grid, UM_THREADS, Cfg::SMEM_BYTES, s>>>(mYhi, mYlo, mXhi, mXlo, rows, Cout, Cin, ntaps, nsplit, chunk,
                                                                  scratch);
    SIMQ_LAUNCH_CHECK();
    long long n = (long long)per_split;
    wgrad_reduce_kernel<<<ceil_div(n, 256), 256, 0, s>>>(scratch, Cout, Cin, ntaps, nsplit, dW);
    prof_mark(PROF_WGRAD, false, 0, s);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

template <int TERMS>
static int launch_wgrad2(const UmmaTensor& dY, const UmmaTensor& X, int ntaps, float* dW, float* scratch, cudaStream_t s) {
    using Cfg = WgradCfg<128, TERMS>;
    static unsigned long long attr = 0;      // per-device: function attributes belong to the device context
    if (first_use_on_device(attr)) SIMQ_CUDA(cudaFuncSetAttribute(wgrad2_umma_kernel<TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    const int g_num_sms = num_sms();
    const int Cout = dY.cols, Cin = X.cols;
    const long long rows = dY.rows;
    CUtensorMap mYhi, mYlo, mXhi, mXlo;
    if (make_map(&mYhi, dY.t.hi, rows, Cout, UM_BK) || make_map(&mYlo, dY.t.lo, rows, Cout, UM_BK) ||
        make_map(&mXhi, X.t.hi, rows, Cin, UM_BK) || make_map(&mXlo, X.t.lo, rows, Cin, UM_BK))
        return 1;
    const int tiles = (Cout / 256) * (Cin / 256) * ntaps, clusters = g_num_sms / 2;
    // row splits: 2-4 rounds of the 74 CTA pairs, the split count chosen so that the last round is as full as possible
    long long max_split = rows / (8 * UM_BK);
    if (max_split < 1) max_split = 1;
    int nsplit = 1;
    double best = -1.0;
    for (int n = 1; n <= 64 && n <= max_split; ++n) {
        const int work = tiles * n, rounds = (work + clusters - 1) / clusters;
        if (rounds > 4 && best > 0) break;
        const double eff = (double)work / ((double)rounds * clusters) - (rounds < 2 ? 0.25 : 0.0);
        if (eff > best + 1e-9) { best = eff; nsplit = n; }
    }
    const size_t per_split = (size_t)ntaps * Cout * Cin;
    while (nsplit > 1 && per_split * nsplit > umma_wgrad_scratch_floats()) --nsplit;
    if (per_split * nsplit > umma_wgrad_scratch_floats()) { simq_set_error("wgrad scratch too small"); return 1; }
    long long chunk = ((rows + nsplit - 1) / nsplit + UM_BK - 1) / UM_BK * UM_BK;
    dim3 grid(2 * (Cout / 256), Cin / 256, ntaps * nsplit);
    prof_mark(PROF_WGRAD, true, 2.0 * (double)rows * 576.0 / 625.0 * Cout * Cin * ntaps, s, 2.0 * TERMS * (double)rows * Cout * Cin * ntaps);
    wgrad2_umma_kernel<TERMS><<<grid, UM_THREADS, Cfg::SMEM_BYTES, s>>>(mYhi, mYlo, mXhi, mXlo, rows, Cout, Cin, ntaps, nsplit, chunk, scratch);
    SIMQ_LAUNCH_CHECK();
    wgrad_reduce_kernel<<<ceil_div((long long)per_split, 256), 256, 0, s>>>(scratch, Cout, Cin, ntaps, nsplit, dW);
    prof_mark(PROF_WGRAD, false, 0, s);
    SIMQ_LAUNCH_CHECK();
    return 0;
}

int k_wgrad_umma(const UmmaTensor& dY, const UmmaTensor& X, int ntaps, float* dW, float* scratch, int terms, cudaStream_t s) {
    if (umma_init()) return 1;
    if (!umma_wgrad_supported(dY.cols, X.cols) || dY.rows != X.rows) {
        simq_set_error("k_wgrad_umma: unsupported shape Cout=%d Cin=%d", dY.cols, X.cols);
        return 1;
    }
    static int pair = -1;
    if (pair < 0) { const char* e = getenv("SIMQ_WGRAD_PAIR"); pair = e ? atoi(e) : 1; }
    if (pair && dY.cols % 256 == 0 && X.cols % 256 == 0)
        return terms == 1 ? launch_wgrad2<1>(dY, X, ntaps, dW, scratch, s) : terms == 2 ? launch_wgrad2<2>(dY, X, ntaps, dW, scratch, s)
                                                                                        : launch_wgrad2<3>(dY, X, ntaps, dW, scratch, s);
    if (X.cols % 128 == 0)
        return terms == 1 ? launch_wgrad<128, 1>(dY, X, ntaps, dW, scratch, s) : terms == 2 ? launch_wgrad<128, 2>(dY, X, ntaps, dW, scratch, s)
                                                                                            : launch_wgrad<128, 3>(dY, X, ntaps, dW, scratch, s);
    return terms == 1 ? launch_wgrad<64, 1>(dY, X, ntaps, dW, scratch, s) : terms == 2 ? launch_wgrad<64, 2>(dY, X, ntaps, dW, scratch, s)
                                                                                       : launch_wgrad<64, 3>(dY, X, ntaps, dW, scratch, s);
}
