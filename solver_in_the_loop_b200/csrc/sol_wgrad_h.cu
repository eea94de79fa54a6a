// Weight gradient of the 32->32 5x5 layers on the tensor cores (tcgen05, block-scaled 3xFP16), DEFERRED over the whole
// unrolled sweep (reference: TF autodiff of the ten hidden Conv2D layers, karman-2d/karman_train.py:107-133,449-457):
//     dW[tap][ci][co] = sum over ALL msteps x B x Y x X pixels of in[pix+tap][ci] * g[pix][co]
// one launch per layer (or per range of steps) instead of one per (layer, step).
//
// Same GEMM mapping as the 3xTF32 kernel (sol_wgrad_tc.cu) — K = pixels, both operands MN-major, M = 128 = 4 taps x 32 cin as
// four overlapping M-blocks of ONE staged halo tile, 7 tap groups x 64 TMEM columns — with fp16 operand pairs:
//   * kind::f16 covers K = 16 pixels (two tile rows) per instruction at 2 bytes per element: half the instructions and half
//     the shared-memory operand bytes of the tf32 form.
//   * fp16 needs a scale.  The accumulators run over ALL tiles of a launch, so the scale must be uniform per operand: the
//     producers of the activations / output gradients (k_conv5x5_c32_h, k_conv5x5_expand) keep a running max|x| per tensor
//     (one atomicMax per warp), and this kernel reads it at its start: S = 2^j with max|x|*S in [2^14, 2^15).  Elements
//     more than 2^16 below the maximum keep an absolute error of 2^-25 (2^-39 of the maximum) — irrelevant in a sum over
//     ~10^6 pixels.  x*S = hi + lo (+ 2^-24 relative), D += A_hi*[G_hi|G_lo] (one N = 64 MMA) + A_lo*G_hi (N = 32).
//   * layout: 16-bit MN-major operands with 64-byte rows (32 channels) use SWIZZLE_64B: an atom is 32 channels x 8 pixels
//     (512 B), the 16-byte chunk index is XORed with address bits [7,9).  The four splitter warps convert the TMA-staged fp32
//     tiles into that layout with generic stores (same address function), so overlapping, shifted atoms stay consistent.
//   * the tensor core accumulates with truncation: a chain of ~650 accumulations per CTA cost 5e-5 relative error in the
//     round-1 kernel.  The accumulators are therefore drained into the per-CTA fp32 partial sums every WGH_DRAIN tiles
//     (RN adds), which bounds the chain at 128 accumulations (K = 16 each).
#include <cuda_fp16.h>

#include <algorithm>

#include "sol_internal.cuh"
#include "sol_tc_common.cuh"

SOL_TRACE_TU()

namespace sol {

namespace {

using namespace tc;

constexpr int WGH_TX = 8, WGH_TY = 16, WGH_HW = 12, WGH_HH = 20;
constexpr int WGH_AF_BYTES = WGH_HH * WGH_HW * 128;     // 30720  fp32 activations halo tile (TMA staging)
constexpr int WGH_GF_BYTES = WGH_TY * WGH_TX * 128;     // 16384  fp32 output-gradient tile (TMA staging)
constexpr int WGH_A_BYTES = WGH_HH * WGH_HW * 64;       // 15360  fp16 hi (or lo) activations
constexpr int WGH_G_BYTES = WGH_TY * WGH_TX * 64;       //  8192  fp16 hi (or lo) gradients
constexpr int WGH_OFF_GF = WGH_AF_BYTES;
constexpr int WGH_OFF_AH = WGH_OFF_GF + WGH_GF_BYTES;   // 47104 = 92 x 512
constexpr int WGH_OFF_AL = WGH_OFF_AH + WGH_A_BYTES;    // 62464 = 122 x 512
constexpr int WGH_OFF_GH = WGH_OFF_AL + WGH_A_BYTES;    // 77824 = 152 x 512
constexpr int WGH_OFF_GL = WGH_OFF_GH + WGH_G_BYTES;    // 86016 = 168 x 512
constexpr int WGH_STAGE_BYTES = WGH_OFF_GL + WGH_G_BYTES;   // 94208 = 92 x 1024
constexpr int WGH_NSTAGE = 2;
constexpr int WGH_OFF_BAR = WGH_NSTAGE * WGH_STAGE_BYTES;
constexpr int WGH_SMEM = WGH_OFF_BAR + 256 + 1024;
constexpr int WGH_THREADS = 224;                        // warp 0 TMA, warps 1 and 6 MMA issuers, warps 2..5 splitters / drain
constexpr int WGH_TMEM_COLS = 512;                      // 7 tap groups x 64 columns used
constexpr int WGH_NGROUP = 7;
constexpr int WGH_DRAIN = 16;                           // tiles between two accumulator drains
static_assert(WGH_OFF_AH % 512 == 0 && WGH_OFF_AL % 512 == 0 && WGH_OFF_GH % 512 == 0 && WGH_OFF_GL % 512 == 0 && WGH_STAGE_BYTES % 1024 == 0,
              "operand tiles start on a swizzle-atom boundary");
// M = 128, D = f32, A = B = f16, both MN-major (bits 15, 16), N = 32 or 64
constexpr uint32_t WGH_IDESC32 = (1u << 4) | (1u << 15) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t WGH_IDESC64 = (1u << 4) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

struct WgHArgs {
    int issuers;                    // 1: warp 1 issues every MMA; 2: the tap groups are split between warps 1 and 6 (two instruction streams)
    float* part;                    // [gridDim.x][25*32*32 + 32]
    const uint32_t* amax_in;        // running max|activation| of the tensor (bit pattern), complete for the steps of this launch
    const uint32_t* amax_g;         // same for the output gradients
    int tiles_x, tiles_y, images;   // tiles per image row / column, number of images (steps*B)
    int B;                          // images per step (the 5th tensor dimension is the step)
    int accumulate;                 // CTAs with blockIdx.x < accumulate add to their partial slot, the others overwrite it
};

__device__ __forceinline__ void umma_f16_mn(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// MN-major 16-bit operand, SWIZZLE_64B (cute::UMMA::LayoutType 4): atom = 32 elements (64 B) x 8 K-rows;
//   lbo = byte distance between consecutive 32-element M/N blocks, sbo = byte distance between consecutive 8-row K groups
__device__ __forceinline__ uint64_t make_desc_mn64(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;       // LayoutType::SWIZZLE_64B
    return d;
}

__device__ __forceinline__ void pow2_scale_w(uint32_t amax_bits, float& s, float& inv_s) {
    const uint32_t e = (amax_bits >> 23) & 0xffu;
    if (e < 20u || e == 255u) { s = 1.0f; inv_s = 1.0f; return; }
    s = __uint_as_float((268u - e) << 23);
    inv_s = __uint_as_float((e - 14u) << 23);
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// 8 scaled channels -> hi chunk (8 halves) + lo chunk; also returns the plain sum contributions for the bias gradient
__device__ __forceinline__ void split8w(const float4& a, const float4& b, float s, uint4& hi, uint4& lo) {
    const float x[8] = {a.x * s, a.y * s, a.z * s, a.w * s, b.x * s, b.y * s, b.z * s, b.w * s};
    float r[8];
    uint32_t hp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __half2 h = __floats2half2_rn(x[2 * k], x[2 * k + 1]);
        const float2 hf = __half22float2(h);
        r[2 * k] = x[2 * k] - hf.x; r[2 * k + 1] = x[2 * k + 1] - hf.y;
        hp[k] = *reinterpret_cast<const uint32_t*>(&h);
    }
    hi = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    lo = make_uint4(pack_h2(r[0], r[1]), pack_h2(r[2], r[3]), pack_h2(r[4], r[5]), pack_h2(r[6], r[7]));
}

}  // namespace

__global__ void __launch_bounds__(WGH_THREADS, 1)
k_wgrad_c32_h(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_g, const WgHArgs a) {
    pdl_sync();
    extern __shared__ uint8_t wgh_smem_raw[];
    const uint32_t raw = smem_u32(wgh_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gbase = wgh_smem_raw + (base - raw);
    const uint32_t s_bar = base + WGH_OFF_BAR;
    // barriers: full[s] (TMA landed), split[s] (hi/lo ready), empty[s] (MMAs of the stage retired), acc (segment complete)
    const uint32_t bar_full = s_bar, bar_split = s_bar + 8 * WGH_NSTAGE, bar_empty = s_bar + 16 * WGH_NSTAGE, bar_acc = s_bar + 24 * WGH_NSTAGE;
    const uint32_t bar_drained = bar_acc + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + WGH_OFF_BAR + 24 * WGH_NSTAGE + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_image = a.tiles_x * a.tiles_y;
    const int ntiles = tiles_per_image * a.images;
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nseg = (my_tiles + WGH_DRAIN - 1) / WGH_DRAIN;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < WGH_NSTAGE; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_split + 8 * s, 128);
            mbar_init(bar_empty + 8 * s, (uint32_t)a.issuers);
        }
        mbar_init(bar_acc, (uint32_t)a.issuers);
        mbar_init(bar_drained, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)WGH_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        const bool leader = elect_one();
        int n = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++n) {
            const int s = n % WGH_NSTAGE;
            const uint32_t ph = (uint32_t)(n / WGH_NSTAGE) & 1u;
            mbar_wait(bar_empty + 8 * s, ph ^ 1u);
            if (leader) {
                const int img = t / tiles_per_image, rem = t - img * tiles_per_image;
                const int tyi = rem / a.tiles_x, txi = rem - tyi * a.tiles_x;
                const int step = img / a.B, bb = img - step * a.B;
                const uint32_t st = base + s * WGH_STAGE_BYTES;
                mbar_arrive_expect_tx(bar_full + 8 * s, WGH_AF_BYTES + WGH_GF_BYTES);
                tma_load_5d(st, &map_in, bar_full + 8 * s, 0, txi * WGH_TX - 2, tyi * WGH_TY - 2, bb, step);
                tma_load_5d(st + WGH_OFF_GF, &map_g, bar_full + 8 * s, 0, txi * WGH_TX, tyi * WGH_TY, bb, step);
            }
        }
    } else if (warp == 1 || warp == 6) {
        // ================= MMA issuer(s) =================
        // One UMMA of this shape costs ~35 clk of serialised issue / dependency overhead on top of its operand fetch when a single
        // thread issues them back to back; two independent instruction streams (disjoint accumulator groups) overlap that overhead.
        if (warp == 6 && a.issuers < 2) goto done;
        const int g_lo = (a.issuers == 2 && warp == 6) ? 4 : 0;
        const int g_hi = (a.issuers == 2 && warp == 1) ? 4 : WGH_NGROUP;
        const bool leader = elect_one();
        int n = 0;
        for (int seg = 0; seg < nseg; ++seg) {
            const int seg_tiles = min(WGH_DRAIN, my_tiles - seg * WGH_DRAIN);
            if (seg > 0) {
                mbar_wait(bar_drained, (uint32_t)(seg - 1) & 1u);       // the accumulators of the previous segment have been read out
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            for (int k = 0; k < seg_tiles; ++k, ++n) {
                const int s = n % WGH_NSTAGE;
                const uint32_t ph = (uint32_t)(n / WGH_NSTAGE) & 1u;
                mbar_wait(bar_split + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (leader) {
                    const uint32_t st = base + s * WGH_STAGE_BYTES;
                    // A: 4 M-blocks = 4 taps; block stride (LBO) = 64 B for horizontally adjacent taps, one halo row (12*64 B) for
                    // vertically adjacent ones; K = 16 = two tile rows: SBO = one halo row.  G: [G_hi | G_lo] as two N-blocks
                    // (LBO = distance between the two tiles), its two tile rows are contiguous: SBO = 512.
                    const uint64_t dAh_hi = make_desc_mn64(st + WGH_OFF_AH, 64, WGH_HW * 64), dAh_lo = make_desc_mn64(st + WGH_OFF_AL, 64, WGH_HW * 64);
                    const uint64_t dAv_hi = make_desc_mn64(st + WGH_OFF_AH, WGH_HW * 64, WGH_HW * 64), dAv_lo = make_desc_mn64(st + WGH_OFF_AL, WGH_HW * 64, WGH_HW * 64);
                    const uint64_t dG64 = make_desc_mn64(st + WGH_OFF_GH, WGH_G_BYTES, 512);      // [hi | lo]
                    const uint64_t dG32 = make_desc_mn64(st + WGH_OFF_GH, 64, 512);               // hi only
#pragma unroll 1
                    for (int y = 0; y < WGH_TY; y += 2) {
                        const uint64_t g_off = (uint64_t)(y * 32);                    // one tile row = 512 B
                        const uint32_t accum = (k == 0 && y == 0) ? 0u : 1u;
                        // pass 0: D[:, 0:64] (+)= A_hi x [G_hi | G_lo]   pass 1: D[:, 0:32] += A_lo x G_hi
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint64_t dG = (j == 0 ? dG64 : dG32) + g_off;
                            const uint32_t idesc = (j == 0) ? WGH_IDESC64 : WGH_IDESC32;
                            const uint32_t acc_flag = (j == 0) ? accum : 1u;
#pragma unroll
                            for (int grp = 0; grp < WGH_NGROUP; ++grp) {
                                if (grp < g_lo || grp >= g_hi) continue;
                                // groups 0..4: taps (dy=grp, dx=0..3); group 5: taps (dy=0..3, dx=4); group 6: tap (4,4) (+3 unused)
                                const int dy0 = (grp < 5) ? grp : (grp == 5 ? 0 : 4);
                                const int dx0 = (grp < 5) ? 0 : 4;
                                const uint64_t a_off = (uint64_t)(((y + dy0) * WGH_HW + dx0) * 4);      // 64-byte pixel rows in 16-byte units
                                const uint64_t dA = (grp == 5) ? ((j == 0) ? dAv_hi : dAv_lo) : ((j == 0) ? dAh_hi : dAh_lo);
                                umma_f16_mn(tmem_acc + 64u * (uint32_t)grp, dA + a_off, dG, idesc, acc_flag);
                            }
                        }
                    }
                    umma_commit(bar_empty + 8 * s);
                }
                __syncwarp();
            }
            if (leader) umma_commit(bar_acc);        // this segment's accumulators are complete when the commit arrives
            __syncwarp();
        }
    } else {
        // ================= splitter (warps 2..5), accumulator drain after every segment =================
        const int tt = threadIdx.x - 64;
        const int q = warp & 3;
        float* part = a.part + (size_t)blockIdx.x * (25 * 32 * 32 + 32);
        float sA, iA, sG, iG;
        pow2_scale_w(__ldg(a.amax_in), sA, iA);
        pow2_scale_w(__ldg(a.amax_g), sG, iG);
        // bias gradient for free: work items are (pixel, channel octet) with the octet fixed per thread (tt & 3)
        float bsum[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) bsum[c] = 0.0f;
        const int oct = tt & 3;
        int n = 0;
        for (int seg = 0; seg < nseg; ++seg) {
            const int seg_tiles = min(WGH_DRAIN, my_tiles - seg * WGH_DRAIN);
            for (int k = 0; k < seg_tiles; ++k, ++n) {
                const int s = n % WGH_NSTAGE;
                const uint32_t ph = (uint32_t)(n / WGH_NSTAGE) & 1u;
                mbar_wait(bar_full + 8 * s, ph);
                uint8_t* st = gbase + s * WGH_STAGE_BYTES;
                // activations: 240 pixels x 4 octets; fp32 chunks (2*oct, 2*oct+1) ^ (p & 7) -> fp16 chunk oct ^ ((p >> 1) & 3)
#pragma unroll 3
                for (int item = tt; item < WGH_HH * WGH_HW * 4; item += 128) {
                    const int p = item >> 2, r = p & 7;
                    const uint8_t* row = st + p * 128;
                    const float4 v0 = *reinterpret_cast<const float4*>(row + (((2 * oct) ^ r) << 4));
                    const float4 v1 = *reinterpret_cast<const float4*>(row + (((2 * oct + 1) ^ r) << 4));
                    uint4 hi, lo;
                    split8w(v0, v1, sA, hi, lo);
                    const int off = p * 64 + ((oct ^ ((p >> 1) & 3)) << 4);
                    *reinterpret_cast<uint4*>(st + WGH_OFF_AH + off) = hi;
                    *reinterpret_cast<uint4*>(st + WGH_OFF_AL + off) = lo;
                }
#pragma unroll 2
                for (int item = tt; item < WGH_TY * WGH_TX * 4; item += 128) {
                    const int p = item >> 2, r = p & 7;
                    const uint8_t* row = st + WGH_OFF_GF + p * 128;
                    const float4 v0 = *reinterpret_cast<const float4*>(row + (((2 * oct) ^ r) << 4));
                    const float4 v1 = *reinterpret_cast<const float4*>(row + (((2 * oct + 1) ^ r) << 4));
                    bsum[0] += v0.x; bsum[1] += v0.y; bsum[2] += v0.z; bsum[3] += v0.w;
                    bsum[4] += v1.x; bsum[5] += v1.y; bsum[6] += v1.z; bsum[7] += v1.w;
                    uint4 hi, lo;
                    split8w(v0, v1, sG, hi, lo);
                    const int off = p * 64 + ((oct ^ ((p >> 1) & 3)) << 4);
                    *reinterpret_cast<uint4*>(st + WGH_OFF_GH + off) = hi;
                    *reinterpret_cast<uint4*>(st + WGH_OFF_GL + off) = lo;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(bar_split + 8 * s);
            }
            // ---- drain: accumulator (group, M-block q), lane m -> tap, cin = lane, 32 columns = cout; RN fp32 adds into the partial slot
            mbar_wait(bar_acc, (uint32_t)seg & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const bool add = ((int)blockIdx.x < a.accumulate) || seg > 0;
#pragma unroll 1
            for (int grp = 0; grp < WGH_NGROUP; ++grp) {
                int dy, dx;
                if (grp < 5) { dy = grp; dx = q; }
                else if (grp == 5) { dy = q; dx = 4; }
                else { dy = 4; dx = 4 + q; }
                uint32_t v0[32], v1[32];
                const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + 64u * (uint32_t)grp;
                tmem_ld_32x32_nowait(taddr, v0);            // hh + lh
                tmem_ld_32x32_nowait(taddr + 32u, v1);      // hl
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (dx < 5) {
                    float4* dst = reinterpret_cast<float4*>(part + (((dy * 5 + dx) * 32 + lane) * 32));
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        // the two scales are applied one after the other: their product may leave the fp32 range
                        float4 f = make_float4((__uint_as_float(v0[4 * c]) + __uint_as_float(v1[4 * c])) * iA * iG,
                                               (__uint_as_float(v0[4 * c + 1]) + __uint_as_float(v1[4 * c + 1])) * iA * iG,
                                               (__uint_as_float(v0[4 * c + 2]) + __uint_as_float(v1[4 * c + 2])) * iA * iG,
                                               (__uint_as_float(v0[4 * c + 3]) + __uint_as_float(v1[4 * c + 3])) * iA * iG);
                        if (add) { const float4 o = dst[c]; f.x += o.x; f.y += o.y; f.z += o.z; f.w += o.w; }
                        dst[c] = f;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(bar_drained);
        }
        {   // bias slot: the per-thread channel sums go through shared memory and are added in thread order (deterministic)
            float* bsm = reinterpret_cast<float*>(gbase);         // [128 threads][8] (all stages are free now)
#pragma unroll
            for (int c = 0; c < 8; ++c) bsm[tt * 8 + c] = bsum[c];
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tt < 32) {
                float sb = 0.0f;
                const int o8 = tt >> 3, c = tt & 7;              // channel tt = octet o8, lane c: threads with (t & 3) == o8 own it
                for (int t = o8; t < 128; t += 4) sb += bsm[t * 8 + c];
                part[25 * 32 * 32 + tt] = ((int)blockIdx.x < a.accumulate) ? part[25 * 32 * 32 + tt] + sb : sb;
            }
        }
    }

done:
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)WGH_TMEM_COLS) : "memory");
    }
}

// running max|x| (bit pattern) of a tensor: the fallback for producers that do not track it themselves
__global__ void __launch_bounds__(256) k_amax(const float4* __restrict__ x, size_t n4, uint32_t* __restrict__ slot) {
    pdl_sync();
    uint32_t m = 0u;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
        const float4 v = __ldg(x + i);
        m = max(m, max(max(__float_as_uint(v.x) & 0x7fffffffu, __float_as_uint(v.y) & 0x7fffffffu),
                       max(__float_as_uint(v.z) & 0x7fffffffu, __float_as_uint(v.w) & 0x7fffffffu)));
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(slot, m);
}

// measured on B200 at the bench shape (profiles/r02/r02_g_wgrad_background.txt): every (CTAs, chunk) setting is SLOWER than no
// background launches (12.80 ms): with 48 SMs taken the conv chain fills every remaining SM two CTAs deep, the early CTAs of the
// next layer find no free slot (no prologue overlap), and the side stream's tail delays the end of the sweep.  Default off.
int g_wgrad_bg_ctas = 0;
int g_wgrad_bg_chunk = 4;
// Also measured and rejected (scripts/probes/wgrad_paired_ctas_experiment.diff, profiles/r02/r02_z_*): TWO co-resident CTAs per SM that
// split the tap groups 4 + 3 (256 TMEM columns and one 95 KB stage each).  Parity-green, but 187 us per layer against 166: unlike the
// K-major conv operands, the MN-major SWIZZLE_64B operand fetch of this kernel is already at its shared-memory rate with one CTA, so
// interleaving two instruction streams gains nothing and the duplicated tile split costs.
int g_wgrad_issuers = 1;     // MMA-issuing warps of k_wgrad_c32_h: measured 12.37 (2) vs 12.39 ms (1) per iteration — the UMMAs of one CTA execute in
                             // order whichever warp issues them (only co-resident CTAs overlap their per-instruction overhead), so one is the default

int launch_amax(cudaStream_t st, const float* x, size_t n, uint32_t* slot) {
    if (n % 4 || ((uintptr_t)x & 15)) return fail(SOL_ERR_INVALID, "amax: needs a 16-byte aligned tensor of 4k floats");
    const size_t n4 = n / 4;
    const int grid = (int)std::min<size_t>((n4 + 255) / 256, 592);
    SOL_CUDA(launch_kernel(k_amax, dim3(grid), dim3(256), 0, st, reinterpret_cast<const float4*>(x), n4, slot));
    SOL_LAUNCHED();
    return SOL_OK;
}

// in: activations of `steps` unrolled steps, image (step, b) at in + step*in_step_stride + b*Y*X*32 floats;
// g:  output gradients, image (step, b) at g + step*g_step_stride + b*Y*X*32.
// amax_in / amax_g: device slots holding max|in| / max|g| over (at least) these steps.
// part: sm_count x (25*32*32+32) floats; dW = sum over CTAs (k_wgrad_finalize).
int launch_wgrad_c32_h(cudaStream_t st, int sm_count, int steps, int B, int Y, int X, const float* in, size_t in_step_stride,
                       const float* g, size_t g_step_stride, const uint32_t* amax_in, const uint32_t* amax_g, float* part, int* nctas_out,
                       int accumulate) {
    tc::EncodeTiledFn enc = tc::get_encode_tiled();
    if (!enc) return fail(SOL_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (X % WGH_TX || Y % WGH_TY) return fail(SOL_ERR_UNSUPPORTED, "wgrad h: needs X % 8 == 0 and Y % 16 == 0");
    if ((in_step_stride * 4) % 16 || (g_step_stride * 4) % 16) return fail(SOL_ERR_INVALID, "wgrad h: step strides must be 16-byte multiples");
    if (!amax_in || !amax_g) return fail(SOL_ERR_INVALID, "wgrad h: operand maxima required");
    alignas(64) CUtensorMap map_in, map_g;
    const cuuint64_t dims[5] = {32, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)B, (cuuint64_t)steps};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    {
        const cuuint64_t strides[4] = {128, (cuuint64_t)X * 128, (cuuint64_t)Y * X * 128, (cuuint64_t)in_step_stride * 4};
        const cuuint32_t box[5] = {32, WGH_HW, WGH_HH, 1, 1};
        if (enc(&map_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return fail(SOL_ERR_CUDA, "cuTensorMapEncodeTiled(wgrad activations) failed");
    }
    {
        const cuuint64_t strides[4] = {128, (cuuint64_t)X * 128, (cuuint64_t)Y * X * 128, (cuuint64_t)g_step_stride * 4};
        const cuuint32_t box[5] = {32, WGH_TX, WGH_TY, 1, 1};
        if (enc(&map_g, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return fail(SOL_ERR_CUDA, "cuTensorMapEncodeTiled(wgrad gradients) failed");
    }
    WgHArgs a;
    a.issuers = g_wgrad_issuers == 1 ? 1 : 2;
    a.part = part; a.amax_in = amax_in; a.amax_g = amax_g;
    a.tiles_x = X / WGH_TX; a.tiles_y = Y / WGH_TY; a.images = steps * B; a.B = B; a.accumulate = accumulate;
    const int ntiles = a.tiles_x * a.tiles_y * a.images;
    int nctas = sm_count < 148 ? sm_count : 148;
    if (nctas > ntiles) nctas = ntiles;
    SOL_CUDA(cudaFuncSetAttribute(k_wgrad_c32_h, cudaFuncAttributeMaxDynamicSharedMemorySize, WGH_SMEM));
    SOL_CUDA(launch_kernel(k_wgrad_c32_h, dim3(nctas), dim3(WGH_THREADS), WGH_SMEM, st, map_in, map_g, a));
    SOL_LAUNCHED();
    if (nctas_out) *nctas_out = nctas;
    return SOL_OK;
}

}  // namespace sol
