// Weight gradient of the 32->32 5x5 layers on the tensor cores (tcgen05, 3xTF32), DEFERRED over the
// whole unrolled sweep: dW[tap][ci][co] = sum over ALL msteps x B x Y x X pixels of
// in[pix+tap][ci] * g[pix][co], one launch per layer instead of one per (layer, step).
//
// The 180 GB of HBM make this the natural B200 formulation: the adjoint sweep leaves every layer's
// output gradient in a stash (1 GB for SOL-32), and the reduction over ~800k pixels becomes ONE
// GEMM with K = pixels per layer — the per-step SIMT kernel with its 148 x 100 KB partial-sum
// round trips per (layer, step) disappears.
//
// GEMM mapping (per CTA, persistent over pixel tiles of 8 x 16):
//   K = pixels.  Both operands are "MN-major" tf32 (layout SWIZZLE_128B_BASE32B, the only legal one for
//   32-bit MN-major operands): a pixel row of the staged tiles is 128 B = 32 channels, 8 consecutive
//   pixels (one tile row) are two 4-row swizzle atoms = one K=8 UMMA step.
//   A = activations:  M = 128 = 4 taps x 32 cin.  The 4 M-blocks of a descriptor are 4 horizontally
//       adjacent taps of the SAME staged halo tile: leading-byte-offset = 128 B (one pixel), start =
//       halo + ((y+dy)*12 + x + dx0)*128.  (The swizzle phase is a function of the absolute
//       shared-memory address, so overlapping, non-1024-aligned atoms are consistent with what TMA wrote.)
//   B = output gradient: N = 32 cout, one atom per tile row.
//   D[(tap,cin)][cout] lives in TMEM: 7 tap groups (5 horizontal dx 0..3, one vertical dx=4/dy 0..3, one
//   for tap (4,4)) x 64 columns; consecutive MMAs go to different accumulators (MMA-latency chains).
//   3xTF32: A and G tiles are split once into hi/lo by the 4 helper warps; [G_hi|G_lo] are two N-blocks of
//   one descriptor, so D[:,0:64] += Ahi*[Ghi|Glo] is ONE N=64 MMA, then D[:,0:32] += Alo*Ghi.
//   The TMEM accumulators persist across all tiles of the CTA; at the end they are written to a
//   per-CTA partial slot, reduced by k_wgrad_finalize (sol_conv.cu).
#include "sol_internal.cuh"
#include "sol_tc_common.cuh"

SOL_TRACE_TU()

namespace sol {

namespace {

using namespace tc;

constexpr int WG_TX = 8, WG_TY = 16, WG_HW = 12, WG_HH = 20;
constexpr int WG_A_BYTES = WG_HH * WG_HW * 128;     // 30720
constexpr int WG_G_BYTES = WG_TY * WG_TX * 128;     // 16384
constexpr int WG_STAGE_BYTES = 2 * WG_A_BYTES + 2 * WG_G_BYTES;   // Ahi, Alo, Ghi, Glo = 94208
constexpr int WG_NSTAGE = 2;
constexpr int WG_OFF_BAR = WG_NSTAGE * WG_STAGE_BYTES;            // 188416
constexpr int WG_SMEM = WG_OFF_BAR + 256 + 1024;
constexpr int WG_THREADS = 192;
constexpr int WG_TMEM_COLS = 512;                                 // 7 tap groups x 64 columns used
constexpr int WG_NGROUP = 7;
// M=128, tf32, A and B MN-major (bits 15, 16), N = 32 or 64
constexpr uint32_t WG_IDESC32 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t WG_IDESC64 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

struct WgTcArgs {
    float* part;        // [gridDim.x][25*32*32 + 32]
    int tiles_x, tiles_y, images;   // tiles per image row / column, number of images (msteps*B)
    int B;              // images per step (the 5th tensor dimension is the step)
    int accumulate;                 // CTAs with blockIdx.x < accumulate add to their partial slot, the others overwrite it
};

}  // namespace

__global__ void __launch_bounds__(WG_THREADS, 1)
k_wgrad_c32_tc(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_g, const WgTcArgs a) {
    pdl_sync();
    extern __shared__ uint8_t wg_smem_raw[];
    const uint32_t raw = smem_u32(wg_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gbase = wg_smem_raw + (base - raw);
    const uint32_t s_bar = base + WG_OFF_BAR;
    // barriers: full[s] (TMA landed), split[s] (hi/lo ready), empty[s] (MMAs retired), acc (all done)
    const uint32_t bar_full = s_bar, bar_split = s_bar + 8 * WG_NSTAGE, bar_empty = s_bar + 16 * WG_NSTAGE, bar_acc = s_bar + 24 * WG_NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + WG_OFF_BAR + 24 * WG_NSTAGE + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_image = a.tiles_x * a.tiles_y;
    const int ntiles = tiles_per_image * a.images;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < WG_NSTAGE; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_split + 8 * s, 128);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)WG_TMEM_COLS) : "memory");
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
            const int s = n % WG_NSTAGE;
            const uint32_t ph = (uint32_t)(n / WG_NSTAGE) & 1u;
            mbar_wait(bar_empty + 8 * s, ph ^ 1u);
            if (leader) {
                const int img = t / tiles_per_image, rem = t - img * tiles_per_image;
                const int tyi = rem / a.tiles_x, txi = rem - tyi * a.tiles_x;
                const int step = img / a.B, bb = img - step * a.B;
                const uint32_t st = base + s * WG_STAGE_BYTES;
                mbar_arrive_expect_tx(bar_full + 8 * s, WG_A_BYTES + WG_G_BYTES);
                tma_load_5d(st, &map_in, bar_full + 8 * s, 0, txi * WG_TX - 2, tyi * WG_TY - 2, bb, step);
                tma_load_5d(st + 2 * WG_A_BYTES, &map_g, bar_full + 8 * s, 0, txi * WG_TX, tyi * WG_TY, bb, step);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const bool leader = elect_one();
        int n = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++n) {
            const int s = n % WG_NSTAGE;
            const uint32_t ph = (uint32_t)(n / WG_NSTAGE) & 1u;
            mbar_wait(bar_split + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (leader) {
                const uint32_t st = base + s * WG_STAGE_BYTES;
                // A: 4 M-blocks = 4 taps; block stride (LBO) = 128 B for horizontally adjacent taps, one
                // halo row (12*128 B) for vertically adjacent ones.  G: [G_hi | G_lo] as two N-blocks
                // (LBO = distance between the two tiles) -> Ahi*Ghi and Ahi*Glo in ONE N=64 MMA.
                const uint64_t dAh_hi = make_desc_mn(st, 128, 512), dAh_lo = make_desc_mn(st + WG_A_BYTES, 128, 512);
                const uint64_t dAv_hi = make_desc_mn(st, WG_HW * 128, 512), dAv_lo = make_desc_mn(st + WG_A_BYTES, WG_HW * 128, 512);
                const uint64_t dG64 = make_desc_mn(st + 2 * WG_A_BYTES, WG_G_BYTES, 512);      // [hi | lo]
                const uint64_t dG32 = make_desc_mn(st + 2 * WG_A_BYTES, 128, 512);             // hi only
#pragma unroll 1
                for (int y = 0; y < WG_TY; ++y) {
                    const uint64_t g_off = (uint64_t)(y * 64);                    // one tile row = 1024 B
                    const uint32_t accum = (n == 0 && y == 0) ? 0u : 1u;
                    // pass 0: D[:, 0:64] (+)= A_hi x [G_hi | G_lo]   pass 1: D[:, 0:32] += A_lo x G_hi
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint64_t dG = (j == 0 ? dG64 : dG32) + g_off;
                        const uint32_t idesc = (j == 0) ? WG_IDESC64 : WG_IDESC32;
                        const uint32_t acc_flag = (j == 0) ? accum : 1u;
#pragma unroll
                        for (int grp = 0; grp < WG_NGROUP; ++grp) {
                            // groups 0..4: taps (dy=grp, dx=0..3); group 5: taps (dy=0..3, dx=4); group 6: tap (4,4) (+3 unused)
                            const int dy0 = (grp < 5) ? grp : (grp == 5 ? 0 : 4);
                            const int dx0 = (grp < 5) ? 0 : 4;
                            const uint64_t a_off = (uint64_t)(((y + dy0) * WG_HW + dx0) * 8);
                            const uint64_t dA = (grp == 5) ? ((j == 0) ? dAv_hi : dAv_lo) : ((j == 0) ? dAh_hi : dAh_lo);
                            umma_tf32(tmem_acc + 64u * (uint32_t)grp, dA + a_off, dG, idesc, acc_flag);
                        }
                    }
                }
                umma_commit(bar_empty + 8 * s);
            }
            __syncwarp();
        }
        if (leader) umma_commit(bar_acc);
        __syncwarp();
    } else {
        // ================= splitter (warps 2..5), then accumulator dump =================
        const int tt = threadIdx.x - 64;
        // bias gradient for free: thread tt always touches the same 16-byte chunk position (tt & 7) of rows
        // with the same (row & 3), i.e. (32B-chunk swizzle) always the same 4 logical channels
        float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
        int n = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++n) {
            const int s = n % WG_NSTAGE;
            const uint32_t ph = (uint32_t)(n / WG_NSTAGE) & 1u;
            mbar_wait(bar_full + 8 * s, ph);
            float4* ahi = reinterpret_cast<float4*>(gbase + s * WG_STAGE_BYTES);
            float4* alo = reinterpret_cast<float4*>(gbase + s * WG_STAGE_BYTES + WG_A_BYTES);
            float4* ghi = reinterpret_cast<float4*>(gbase + s * WG_STAGE_BYTES + 2 * WG_A_BYTES);
            float4* glo = reinterpret_cast<float4*>(gbase + s * WG_STAGE_BYTES + 2 * WG_A_BYTES + WG_G_BYTES);
#pragma unroll 4
            for (int i = tt; i < WG_A_BYTES / 16; i += 128) {
                const float4 v = ahi[i];
                float4 h, l;
                h.x = tf32_rn(v.x); l.x = tf32_rn(v.x - h.x);
                h.y = tf32_rn(v.y); l.y = tf32_rn(v.y - h.y);
                h.z = tf32_rn(v.z); l.z = tf32_rn(v.z - h.z);
                h.w = tf32_rn(v.w); l.w = tf32_rn(v.w - h.w);
                ahi[i] = h; alo[i] = l;
            }
#pragma unroll 4
            for (int i = tt; i < WG_G_BYTES / 16; i += 128) {
                const float4 v = ghi[i];
                bsum.x += v.x; bsum.y += v.y; bsum.z += v.z; bsum.w += v.w;
                float4 h, l;
                h.x = tf32_rn(v.x); l.x = tf32_rn(v.x - h.x);
                h.y = tf32_rn(v.y); l.y = tf32_rn(v.y - h.y);
                h.z = tf32_rn(v.z); l.z = tf32_rn(v.z - h.z);
                h.w = tf32_rn(v.w); l.w = tf32_rn(v.w - h.w);
                ghi[i] = h; glo[i] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(bar_split + 8 * s);
        }
        // ---- dump: accumulator (dy, g), lane m -> tap dx = 4g + m/32, cin = m%32, 32 columns = cout
        mbar_wait(bar_acc, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        float* part = a.part + (size_t)blockIdx.x * (25 * 32 * 32 + 32);
        {   // bias slot: reduce the per-thread channel sums through shared memory (stage 0 is free now)
            float* bsm = reinterpret_cast<float*>(gbase);
            if (tt < 32) bsm[tt] = 0.0f;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int row = tt >> 3, cphys = tt & 7;                       // G tile: 8 chunks of 16 B per pixel row
            const int ch = ((((cphys >> 1) ^ (row & 3)) << 1) | (cphys & 1)) * 4;   // undo the 32-byte-chunk swizzle
            atomicAdd(bsm + ch + 0, bsum.x); atomicAdd(bsm + ch + 1, bsum.y);
            atomicAdd(bsm + ch + 2, bsum.z); atomicAdd(bsm + ch + 3, bsum.w);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tt < 32) part[25 * 32 * 32 + tt] = ((int)blockIdx.x < a.accumulate) ? part[25 * 32 * 32 + tt] + bsm[tt] : bsm[tt];
        }
#pragma unroll 1
        for (int grp = 0; grp < WG_NGROUP; ++grp) {
            // M-block q of group grp -> tap
            int dy, dx;
            if (grp < 5) { dy = grp; dx = q; }
            else if (grp == 5) { dy = q; dx = 4; }
            else { dy = 4; dx = 4 + q; }
            uint32_t v0[32], v1[32];
            const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + 64u * (uint32_t)grp;
            tmem_ld_32x32(taddr, v0);            // hh + lh
            tmem_ld_32x32(taddr + 32u, v1);      // hl
            if (dx < 5) {
                float4* dst = reinterpret_cast<float4*>(part + (((dy * 5 + dx) * 32 + lane) * 32));
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4 f = make_float4(__uint_as_float(v0[4 * c]) + __uint_as_float(v1[4 * c]),
                                           __uint_as_float(v0[4 * c + 1]) + __uint_as_float(v1[4 * c + 1]),
                                           __uint_as_float(v0[4 * c + 2]) + __uint_as_float(v1[4 * c + 2]),
                                           __uint_as_float(v0[4 * c + 3]) + __uint_as_float(v1[4 * c + 3]));
                    if ((int)blockIdx.x < a.accumulate) { const float4 o = dst[c]; f.x += o.x; f.y += o.y; f.z += o.z; f.w += o.w; }
                    dst[c] = f;
                }
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)WG_TMEM_COLS) : "memory");
    }
}

// db[co] (+)= sum over pixels of g[pix][co]
__global__ void __launch_bounds__(256) k_colsum32(const float* __restrict__ g, size_t npix, float* db) {
    pdl_sync();
    const int c4 = threadIdx.x & 7;          // 4 channels per thread
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t p = (size_t)blockIdx.x * 32 + (threadIdx.x >> 3); p < npix; p += (size_t)gridDim.x * 32) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(g + p * 32) + c4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    __shared__ float4 sm[256];
    sm[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < 8) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = threadIdx.x; k < 256; k += 8) { t.x += sm[k].x; t.y += sm[k].y; t.z += sm[k].z; t.w += sm[k].w; }
        atomicAdd(db + 4 * threadIdx.x + 0, t.x);
        atomicAdd(db + 4 * threadIdx.x + 1, t.y);
        atomicAdd(db + 4 * threadIdx.x + 2, t.z);
        atomicAdd(db + 4 * threadIdx.x + 3, t.w);
    }
}

int launch_colsum32(cudaStream_t st, const float* g, size_t npix, float* db) {
    SOL_CUDA(launch_kernel(k_colsum32, dim3(148), dim3(256), 0, st, g, npix, db));
    SOL_LAUNCHED();
    return SOL_OK;
}

// in: activations of `steps` unrolled steps, image (step, b) at in + step*in_step_stride + b*Y*X*32 floats;
// g:  output gradients, image (step, b) at g + step*g_step_stride + b*Y*X*32.
// part: sm_count x (25*32*32+32) floats; dW = sum over CTAs (k_wgrad_finalize).
int g_wgrad_window_us = 110;   // time budget of one solve window at 128x64 (scaled with the cell count)
int g_wgrad_overlap = 1;    // deferred weight-gradient GEMMs run beside the adjoint pressure solves (see sol_engine.cu)

int launch_wgrad_c32_tc(cudaStream_t st, int sm_count, int steps, int B, int Y, int X, const float* in, size_t in_step_stride,
                        const float* g, size_t g_step_stride, float* part, int* nctas_out, int accumulate) {
    tc::EncodeTiledFn enc = tc::get_encode_tiled();
    if (!enc) return fail(SOL_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (X % WG_TX || Y % WG_TY) return fail(SOL_ERR_UNSUPPORTED, "wgrad tc: needs X % 8 == 0 and Y % 16 == 0");
    if ((in_step_stride * 4) % 16 || (g_step_stride * 4) % 16) return fail(SOL_ERR_INVALID, "wgrad tc: step strides must be 16-byte multiples");
    alignas(64) CUtensorMap map_in, map_g;
    const cuuint64_t dims[5] = {32, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)B, (cuuint64_t)steps};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    {
        const cuuint64_t strides[4] = {128, (cuuint64_t)X * 128, (cuuint64_t)Y * X * 128, (cuuint64_t)in_step_stride * 4};
        const cuuint32_t box[5] = {32, WG_HW, WG_HH, 1, 1};
        if (enc(&map_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return fail(SOL_ERR_CUDA, "cuTensorMapEncodeTiled(wgrad activations) failed");
    }
    {
        const cuuint64_t strides[4] = {128, (cuuint64_t)X * 128, (cuuint64_t)Y * X * 128, (cuuint64_t)g_step_stride * 4};
        const cuuint32_t box[5] = {32, WG_TX, WG_TY, 1, 1};
        if (enc(&map_g, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return fail(SOL_ERR_CUDA, "cuTensorMapEncodeTiled(wgrad gradients) failed");
    }
    WgTcArgs a;
    a.part = part; a.tiles_x = X / WG_TX; a.tiles_y = Y / WG_TY; a.images = steps * B; a.B = B; a.accumulate = accumulate;
    const int ntiles = a.tiles_x * a.tiles_y * a.images;
    int nctas = sm_count < 148 ? sm_count : 148;
    if (nctas > ntiles) nctas = ntiles;
    static bool attr_done = false;
    if (!attr_done) {
        SOL_CUDA(cudaFuncSetAttribute(k_wgrad_c32_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
        attr_done = true;
    }
    SOL_CUDA(launch_kernel(k_wgrad_c32_tc, dim3(nctas), dim3(WG_THREADS), WG_SMEM, st, map_in, map_g, a));
    SOL_LAUNCHED();
    if (nctas_out) *nctas_out = nctas;
    return SOL_OK;
}

}  // namespace sol
