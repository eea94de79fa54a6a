// 5x5 32->32 convolution on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32-accurate through a
// block-scaled 3xFP16 operand split.  Default path of the engine for the ten hidden layers of model_mars_moon
// (reference: karman-2d/karman_train.py:107-133) and — with flipped weights — their data gradients.
//
// Why fp16 pairs instead of tf32 pairs (sol_conv_tc.cu): an N <= 64 UMMA is bound by the shared-memory operand read
// (128 B/clk), not by the tensor pipe.  A tf32 element costs 4 bytes of operand traffic for 11 mantissa bits, an fp16
// element 2 bytes for the same 11 bits, and one kind::f16 instruction covers K = 16 instead of 8: half the operand
// bytes and half the instructions per product.  What fp16 lacks is exponent range, which a power-of-two scale per
// staged tile (activations) and per layer (weights) restores exactly:
//     x*S = hi + lo + e,   hi = rn_fp16(x*S),  lo = rn_fp16(x*S - hi),  |e| <= 2^-24 |x*S|     (S: max|x*S| in [2^14, 2^15))
// elements more than 2^16 below the tile maximum keep an ABSOLUTE error of 2^-25 (2^-39 of the maximum).
//     D = A_hi*B_hi + A_hi*B_lo + A_lo*B_hi   (fp32 accumulation in TMEM; the dropped lo*lo term is 2^-24 relative)
// and the epilogue multiplies by 1/(S*T).
//
// Implicit GEMM per CTA:  D[128 pixels, 32 cout] += sum over 25 taps, 32 cin
//   * CTA tile = 8 (x) x 16 (y) output pixels -> UMMA M = 128, K = 16 per instruction (kind::f16)
//   * ONE TMA box load brings the fp32 halo tile [20 rows][12 px][32 ch] (30 KB, 128B-swizzled, out-of-image pixels
//     zero-filled = Keras 'same' padding).  All six warps convert it IN PLACE into the packed operand form: a pixel stays
//     one 128-byte swizzle row, now [32 x hi fp16 | 32 x lo fp16].  The A operand of tap (dy,dx), half h, k-step ks is
//     the same tile addressed with start = base + (dy*12+dx)*128 + h*64 + ks*32 and stride-byte-offset = one halo row.
//   * weights are pre-split once per optimiser step into [13 tap pairs][64 rows = hi n | lo n][tap a cin | tap b cin]
//     fp16 (one 128-byte row holds the K = 32 of TWO taps), streamed per pair (8 KB) through a TMA/mbarrier ring.
//     A_hi x [B_hi | B_lo] is ONE N = 64 instruction, A_lo x B_hi one N = 32 instruction: 4 MMAs per tap (tf32: 8).
//   * warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2..5 = epilogue
//     (tcgen05.ld -> scale, bias / residual / LeakyReLU(-derivative) -> swizzled staging -> one bulk tensor store).
#include <cuda_fp16.h>
#include <string.h>

#include "sol_internal.cuh"
#include "sol_tc_common.cuh"

SOL_TRACE_TU()

namespace sol {

namespace {

constexpr int H_TX = 8, H_TY = 16, H_HW = 12, H_HH = 20;
constexpr int H_A_BYTES = H_HH * H_HW * 128;          // 30720: fp32 halo tile, converted in place to packed fp16 hi|lo
constexpr int H_STAGE_BYTES = 8192;                   // one tap pair: 64 rows x 128 B
constexpr int H_NPAIR = 13;
constexpr int H_THREADS = 192;
constexpr int H_OFF_B = H_A_BYTES;                    // 1024-aligned (30 x 1024)
constexpr int h_smem_bytes(int stages) { return H_OFF_B + stages * H_STAGE_BYTES + 256 + 1024; }     // + barriers / scratch + alignment slack
constexpr int H_WPAIR_FLOATS = H_NPAIR * 64 * 32;     // split weights of one layer, in floats (two halves each)

struct HArgs {
    const float* bias;
    const float* addend;
    const float* ref;
    const float* w_inv_scale;   // 1/T of the pre-split weights (written by k_prep_h_weights)
    int B, Y, X;
    int act;
    float slope;
    int weights_ready;          // 1: the split weights were complete before the previous kernel of the stream started
    unsigned int* amax_out;     // optional: running max|out| (bit pattern) of the output tensor, for the fp16 weight-gradient GEMM
    long long* trace;           // diagnostics: 16 slots per CTA of phase stamps (null in production)
};

using namespace tc;

// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32 [4,6) = 1, A = B = f16 (format 0), both K-major,
// N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t H_IDESC32 = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t H_IDESC64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void h_stamp(long long* trace, int slot) {
    if (trace) {
        const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        trace[cta * 16 + slot] = clock64();
    }
}

// power-of-two scale that puts amax into [2^14, 2^15); 1 for zero / denormal / non-finite maxima
__device__ __forceinline__ void pow2_scale(uint32_t amax_bits, float& s, float& inv_s) {
    const uint32_t e = (amax_bits >> 23) & 0xffu;
    if (e < 20u || e == 255u) { s = 1.0f; inv_s = 1.0f; return; }
    s = __uint_as_float((268u - e) << 23);          // 2^(14 - (e - 127))
    inv_s = __uint_as_float((e - 14u) << 23);       // 2^((e - 127) - 14)
}

__device__ __forceinline__ uint32_t absbits(float f) { return __float_as_uint(f) & 0x7fffffffu; }

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// hi/lo fp16 split of 8 scaled channels -> one 16-byte chunk of hi halves, one of lo halves
__device__ __forceinline__ void split8(const float4& a, const float4& b, float s, uint4& hi, uint4& lo) {
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
    lo = make_uint4(pack_half2(r[0], r[1]), pack_half2(r[2], r[3]), pack_half2(r[4], r[5]), pack_half2(r[6], r[7]));
}

}  // namespace

// NSET: independent accumulator sets used alternately (chained UMMAs into one TMEM tile serialise on the MMA latency and
//       accumulate with truncation; the sets are summed with RN fp32 adds in the epilogue).
// MERGE: the A_lo x B_hi product accumulates into the columns of A_hi x B_lo (64 instead of 96 columns per set).
// H_STAGES: weight ring depth in tap pairs.  MINB: CTAs per SM the register allocation must allow.
template <int NSET, bool MERGE, int H_STAGES, int MINB>
__global__ void __launch_bounds__(H_THREADS, MINB)
k_conv5x5_c32_h(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_w,
                const __grid_constant__ CUtensorMap map_out, const HArgs a) {
    constexpr int SETW = MERGE ? 64 : 96;                   // TMEM columns per accumulator set
    constexpr int NBLK = NSET * SETW / 32;                  // 32-column blocks the epilogue sums
    constexpr uint32_t TMEM_COLS = NSET * SETW <= 64 ? 64u : (NSET * SETW <= 128 ? 128u : 256u);
    constexpr int H_OFF_BAR = H_OFF_B + H_STAGES * H_STAGE_BYTES;
    extern __shared__ uint8_t h_smem_raw[];
    const uint32_t raw = smem_u32(h_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gbase = h_smem_raw + (base - raw);
    const uint32_t s_a = base, s_b = base + H_OFF_B, s_bar = base + H_OFF_BAR;
    const uint32_t bar_afull = s_bar + 0, bar_asplit = s_bar + 8, bar_acc = s_bar + 16;
    const uint32_t bar_bfull = s_bar + 32, bar_bempty = s_bar + 32 + 8 * H_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + H_OFF_BAR + 32 + 16 * H_STAGES);
    uint32_t* amax_slot = tmem_slot + 2;                    // 6 per-warp maxima

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x0 = blockIdx.x * H_TX, y0 = blockIdx.y * H_TY, b = blockIdx.z;
    // every CTA streams the same 13 weight tiles: rotate the order per CTA so that co-resident CTAs do not hammer the
    // same L2 lines at the same time
    const int pair0 = (int)((blockIdx.x + blockIdx.y * gridDim.x + blockIdx.z * gridDim.x * gridDim.y) * 5u % (unsigned)H_NPAIR);

    if (threadIdx.x == 0 && a.trace) {
        const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        a.trace[cta * 16 + 0] = smid; a.trace[cta * 16 + 1] = (long long)gt;
        h_stamp(a.trace, 2);
    }
    if (warp == 0 && lane == 0) {
        mbar_init(bar_afull, 1);
        mbar_init(bar_asplit, H_THREADS);
        mbar_init(bar_acc, 1);
        for (int s = 0; s < H_STAGES; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_slot;
    if (threadIdx.x == 0) h_stamp(a.trace, 3);      // set-up done (barriers, TMEM)

    const int npre = a.weights_ready ? H_STAGES : 0;
    if (warp == 0) {
        // ================= TMA producer, part 1 (warp-uniform control flow, one elected lane issues) =================
        const bool leader = elect_one();
        // weights that were settled before the predecessor kernel started do not depend on it: fill the ring before
        // waiting for the predecessor (programmatic dependent launch), then fetch the halo tile
        for (int n = 0; n < npre; ++n) {
            int pair = pair0 + n; if (pair >= H_NPAIR) pair -= H_NPAIR;
            if (leader) {
                mbar_arrive_expect_tx(bar_bfull + 8 * n, H_STAGE_BYTES);
                tma_load_2d(s_b + n * H_STAGE_BYTES, &map_w, bar_bfull + 8 * n, 0, pair * 64);
            }
        }
        pdl_wait();
        if (leader) {
            mbar_arrive_expect_tx(bar_afull, H_A_BYTES);
            tma_load_4d(s_a, &map_in, bar_afull, 0, x0 - 2, y0 - 2, b);
        }
    } else {
        pdl_wait();
    }

    // ================= operand conversion by ALL warps, in place: fp32 [32 ch] -> [32 hi | 32 lo] fp16 per pixel =================
    mbar_wait(bar_afull, 0);
    // The next kernel of the stream may become resident from here on: this CTA no longer reads its input from global memory
    if (threadIdx.x == 64) pdl_trigger();
    if (threadIdx.x == 64) h_stamp(a.trace, 4);            // halo tile landed
    float inv_s;
    {
        // work item = (pixel row p, channel octet k): two fp32 16-byte chunks in, one hi and one lo fp16 chunk out.
        // 240 x 4 items, 5 per thread; a warp covers 8 rows x 4 octets: every load / store instruction is conflict-free.
        constexpr int ITEMS = H_HH * H_HW * 4 / H_THREADS;      // 5
        static_assert(H_HH * H_HW * 4 % H_THREADS == 0, "split geometry");
        float4 va[ITEMS], vb[ITEMS];
        uint32_t amax = 0u;
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const int item = (int)threadIdx.x + it * H_THREADS;
            const int p = item >> 2, k = item & 3, r = p & 7;
            const uint8_t* row = gbase + p * 128;
            va[it] = *reinterpret_cast<const float4*>(row + (((2 * k) ^ r) << 4));
            vb[it] = *reinterpret_cast<const float4*>(row + (((2 * k + 1) ^ r) << 4));
            // integer max of the magnitude bit patterns: NaN patterns sort above every finite value and survive (fail loudly)
            amax = max(amax, max(max(absbits(va[it].x), absbits(va[it].y)), max(absbits(va[it].z), absbits(va[it].w))));
            amax = max(amax, max(max(absbits(vb[it].x), absbits(vb[it].y)), max(absbits(vb[it].z), absbits(vb[it].w))));
        }
        amax = __reduce_max_sync(0xffffffffu, amax);
        if (lane == 0) amax_slot[warp] = amax;
        __syncthreads();                                   // every thread has read its items: in-place writes are safe
        uint32_t m = amax_slot[0];
#pragma unroll
        for (int w = 1; w < H_THREADS / 32; ++w) m = max(m, amax_slot[w]);
        float s;
        pow2_scale(m, s, inv_s);
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const int item = (int)threadIdx.x + it * H_THREADS;
            const int p = item >> 2, k = item & 3, r = p & 7;
            uint8_t* row = gbase + p * 128;
            uint4 hi, lo;
            split8(va[it], vb[it], s, hi, lo);
            *reinterpret_cast<uint4*>(row + ((k ^ r) << 4)) = hi;
            *reinterpret_cast<uint4*>(row + (((4 + k) ^ r) << 4)) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
        mbar_arrive(bar_asplit);
    }
    if (threadIdx.x == 64) h_stamp(a.trace, 5);            // conversion done (this thread)

    if (warp == 0) {
        // ================= TMA producer, part 2: keep the weight ring full =================
        const bool leader = elect_one();
#pragma unroll 1
        for (int n = npre; n < H_NPAIR; ++n) {
            const int s = n % H_STAGES;
            const uint32_t ph = (uint32_t)(n / H_STAGES) & 1u;
            int pair = pair0 + n; if (pair >= H_NPAIR) pair -= H_NPAIR;
            mbar_wait(bar_bempty + 8 * s, ph ^ 1u);     // first pass: fresh barrier, parity 1 passes
            if (leader) {
                mbar_arrive_expect_tx(bar_bfull + 8 * s, H_STAGE_BYTES);
                tma_load_2d(s_b + s * H_STAGE_BYTES, &map_w, bar_bfull + 8 * s, 0, pair * 64);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const bool leader = elect_one();
        const uint64_t dA = make_desc(s_a, H_HW * 128, 0);
        const uint64_t dB = make_desc(s_b, 1024, 0);
        mbar_wait(bar_asplit, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (leader) h_stamp(a.trace, 6);            // operands ready, MMA may start
        int cnt = 0;
#pragma unroll 1
        for (int n = 0; n < H_NPAIR; ++n) {
            const int s = n % H_STAGES;
            const uint32_t ph = (uint32_t)(n / H_STAGES) & 1u;
            int pair = pair0 + n; if (pair >= H_NPAIR) pair -= H_NPAIR;
            mbar_wait(bar_bfull + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (leader) {
                if (n == 0) h_stamp(a.trace, 7);    // first weight stage landed
                const int ntap = pair == H_NPAIR - 1 ? 1 : 2;
#pragma unroll 1
                for (int h = 0; h < ntap; ++h) {
                    const int tap = 2 * pair + h;
                    const int dy = tap / 5, dx = tap - dy * 5;
                    const uint64_t a_off = (uint64_t)((dy * H_HW + dx) * 8);                   // 128-byte rows in 16-byte units
                    const uint64_t b_off = (uint64_t)(s * (H_STAGE_BYTES / 16) + h * 4);       // second tap of the pair: +64 B
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks, ++cnt) {
                        const uint32_t set = (uint32_t)cnt & (uint32_t)(NSET - 1);
                        const uint32_t acc1 = tmem_acc + set * (uint32_t)SETW, acc2 = acc1 + (MERGE ? 32u : 64u);
                        const uint32_t accum = cnt < NSET ? 0u : 1u;
                        umma_f16(acc1, dA + a_off + 2 * ks, dB + b_off + 2 * ks, H_IDESC64, accum);                       // A_hi x [B_hi | B_lo]
                        umma_f16(acc2, dA + a_off + 4 + 2 * ks, dB + b_off + 2 * ks, H_IDESC32, MERGE ? 1u : accum);      // A_lo x B_hi
                    }
                }
                umma_commit(bar_bempty + 8 * s);     // frees the weight slot when these MMAs retire
            }
            __syncwarp();
        }
        if (leader) { umma_commit(bar_acc); h_stamp(a.trace, 8); }   // all MMAs issued
        __syncwarp();
    } else {
        // ================= epilogue (warps 2..5 = 128 threads) =================
        const int t = threadIdx.x - 64;
        const int q = warp & 3;                 // this warp may touch TMEM lanes [32q, 32q+32)
        const int r = q * 32 + lane;            // accumulator row = pixel
        const int gy = y0 + (r >> 3), gx = x0 + (r & 7);
        const bool inside = gy < a.Y && gx < a.X;
        const size_t o = (((size_t)b * a.Y + gy) * a.X + gx) * 32;
        // the residual row and the sign mask of the activation reference are fetched while the tensor core is busy
        float4 ad[8];
        uint32_t pos = 0xffffffffu;             // bit c: reference channel c > 0
#pragma unroll
        for (int c = 0; c < 8; ++c) ad[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (inside) {
            if (a.addend) {
#pragma unroll
                for (int c = 0; c < 8; ++c) ad[c] = __ldg(reinterpret_cast<const float4*>(a.addend + o) + c);
            }
            if (a.act == SOL_ACT_DLRELU) {
                pos = 0u;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 rf = __ldg(reinterpret_cast<const float4*>(a.ref + o) + c);
                    pos |= (rf.x > 0.f ? 1u : 0u) << (4 * c) | (rf.y > 0.f ? 1u : 0u) << (4 * c + 1) |
                           (rf.z > 0.f ? 1u : 0u) << (4 * c + 2) | (rf.w > 0.f ? 1u : 0u) << (4 * c + 3);
                }
            }
        }
        if (a.bias) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias) + c);
                ad[c].x += bv.x; ad[c].y += bv.y; ad[c].z += bv.z; ad[c].w += bv.w;
            }
        }
        const float inv_t = __ldg(a.w_inv_scale);      // applied one after the other: the product of the two scales may leave the fp32 range
        mbar_wait(bar_acc, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (t == 0) h_stamp(a.trace, 9);            // accumulators complete
        float acc[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = 0.f;
        const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int j = 0; j + 1 < NBLK; j += 2) {         // two loads in flight per wait
            uint32_t v0[32], v1[32];
            tmem_ld_32x32_nowait(taddr + 32u * (uint32_t)j, v0);
            tmem_ld_32x32_nowait(taddr + 32u * (uint32_t)(j + 1), v1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] += __uint_as_float(v0[c]) + __uint_as_float(v1[c]);
        }
        if (NBLK & 1) {
            uint32_t v0[32];
            tmem_ld_32x32_nowait(taddr + 32u * (uint32_t)(NBLK - 1), v0);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] += __uint_as_float(v0[c]);
        }
        // Output tile -> shared memory (the operand tile is free once the accumulators are complete) in the 128B-swizzled
        // layout of the output tensor map, then ONE bulk tensor store per CTA (clipped at the image border by the TMA unit).
        uint32_t omax = 0u;
        {
            uint8_t* stage = gbase + r * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 f = make_float4(fmaf(acc[4 * c] * inv_s, inv_t, ad[c].x), fmaf(acc[4 * c + 1] * inv_s, inv_t, ad[c].y),
                                       fmaf(acc[4 * c + 2] * inv_s, inv_t, ad[c].z), fmaf(acc[4 * c + 3] * inv_s, inv_t, ad[c].w));
                if (a.act == SOL_ACT_LRELU) {
                    f.x = f.x > 0.f ? f.x : a.slope * f.x; f.y = f.y > 0.f ? f.y : a.slope * f.y;
                    f.z = f.z > 0.f ? f.z : a.slope * f.z; f.w = f.w > 0.f ? f.w : a.slope * f.w;
                } else if (a.act == SOL_ACT_DLRELU) {
                    f.x = (pos >> (4 * c)) & 1u ? f.x : a.slope * f.x; f.y = (pos >> (4 * c + 1)) & 1u ? f.y : a.slope * f.y;
                    f.z = (pos >> (4 * c + 2)) & 1u ? f.z : a.slope * f.z; f.w = (pos >> (4 * c + 3)) & 1u ? f.w : a.slope * f.w;
                }
                *reinterpret_cast<float4*>(stage + ((c ^ (r & 7)) << 4)) = f;
                omax = max(omax, max(max(absbits(f.x), absbits(f.y)), max(absbits(f.z), absbits(f.w))));
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> bulk store reads
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (t == 0) {
            tma_store_4d(&map_out, s_a, 0, x0, y0, b);
            tma_store_commit();
        }
        if (a.amax_out) {                // rows outside the image hold garbage the bulk store clips: they do not count
            omax = __reduce_max_sync(0xffffffffu, inside ? omax : 0u);
            if (lane == 0 && omax) atomicMax(a.amax_out, omax);
        }
        if (t == 0) tma_store_wait_read();       // shared memory must stay valid until the bulk store has read it
    }

    if (threadIdx.x == 64 && a.trace) {              // the thread that issued the bulk store: end of this CTA's useful work
        const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        a.trace[cta * 16 + 12] = (long long)gt;
        h_stamp(a.trace, 10);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TMEM_COLS) : "memory");
    }
}

// wsplit[pair][row][64 halves]: row < 32: hi of Bt[n = row], row >= 32: lo of Bt[n = row - 32];  halves 0..31 = cin of tap
// 2*pair, 32..63 = cin of tap 2*pair+1 (zero for the missing 26th tap);  Bt[tap][n][k] = w[tap][k][n] * T  (w: Keras
// [5,5,K=Cin,N=Cout]).  T = 2^j puts max|w| into [2^14, 2^15); 1/T goes to the float slot behind the 13 pairs.
// One launch splits `gridDim.y` layers: layer l reads w + l*w_stride and writes wsplit + l*split_stride (floats).
__global__ void __launch_bounds__(1024) k_prep_h_weights(const float* __restrict__ w_base, size_t w_stride, float* __restrict__ split_base,
                                                         size_t split_stride) {
    pdl_sync();
    const float* w = w_base + (size_t)blockIdx.y * w_stride;
    __half* wsplit = reinterpret_cast<__half*>(split_base + (size_t)blockIdx.y * split_stride);
    float* inv_scale = split_base + (size_t)blockIdx.y * split_stride + H_WPAIR_FLOATS;
    __shared__ uint32_t red[32];
    uint32_t amax = 0u;
    for (int i = threadIdx.x; i < 25 * 32 * 32; i += 1024) amax = max(amax, __float_as_uint(__ldg(w + i)) & 0x7fffffffu);
    amax = __reduce_max_sync(0xffffffffu, amax);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
    __syncthreads();
    uint32_t m = red[0];
#pragma unroll
    for (int k = 1; k < 32; ++k) m = max(m, red[k]);
    float T, invT;
    pow2_scale(m, T, invT);
    const int pair = blockIdx.x;
    for (int idx = threadIdx.x; idx < 64 * 64; idx += 1024) {
        // consecutive threads read consecutive n (the contiguous index of w) and write 128-byte-strided halves: the reads coalesce
        const int n = idx & 31, lo = (idx >> 5) & 1, col = idx >> 6;
        const int tap = 2 * pair + (col >> 5), k = col & 31, row = lo * 32 + n;
        const float v = tap < 25 ? __ldg(w + (tap * 32 + k) * 32 + n) * T : 0.0f;
        const __half h = __float2half_rn(v);
        wsplit[((size_t)pair * 64 + row) * 64 + col] = lo ? __float2half_rn(v - __half2float(h)) : h;
    }
    if (pair == 0 && threadIdx.x == 0) *inv_scale = invT;
}

int g_conv_variant = 0;     // accumulator layout of k_conv5x5_c32_h: 0 = two merged sets of 64 columns (default), 1 = two sets of 96, 2 = one set of 96

size_t h_weights_floats() { return (size_t)H_WPAIR_FLOATS + 64; }      // 13 pairs + the scale slot, 256-byte multiple

int launch_prep_h_weights(cudaStream_t st, const float* w, float* wsplit, int nlayers, size_t w_stride, size_t split_stride) {
    SOL_CUDA(launch_kernel(k_prep_h_weights, dim3(H_NPAIR, nlayers), dim3(1024), 0, st, w, w_stride, wsplit, split_stride));
    SOL_LAUNCHED();
    return SOL_OK;
}

static long long* g_h_trace = nullptr;
static int g_h_trace_cap = 0, g_h_trace_seq = 0;     // capacity in launches, launches traced so far

int launch_conv5x5_h(cudaStream_t st, int B, int Y, int X, const float* in, const float* wsplit, const float* bias,
                     const float* addend, const float* ref, int act, float slope, float* out, bool weights_ready, unsigned int* amax_out) {
    tc::EncodeTiledFn enc = tc::get_encode_tiled();
    if (!enc) return fail(SOL_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (((uintptr_t)in & 15) || ((uintptr_t)wsplit & 15) || ((uintptr_t)out & 15)) return fail(SOL_ERR_INVALID, "conv h: operands must be 16-byte aligned");
    if (addend && ((uintptr_t)addend & 15)) return fail(SOL_ERR_INVALID, "conv h: addend must be 16-byte aligned");
    if (ref && ((uintptr_t)ref & 15)) return fail(SOL_ERR_INVALID, "conv h: ref must be 16-byte aligned");
    alignas(64) CUtensorMap map_in, map_w, map_out;
    {
        cuuint64_t dims[4] = {32, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)B};
        cuuint64_t strides[3] = {128, (cuuint64_t)X * 128, (cuuint64_t)Y * X * 128};
        cuuint32_t box[4] = {32, H_HW, H_HH, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SOL_ERR_CUDA, "cuTensorMapEncodeTiled(activations) failed");
        cuuint32_t obox[4] = {32, H_TX, H_TY, 1};
        r = enc(&map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)out, dims, strides, obox, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SOL_ERR_CUDA, "cuTensorMapEncodeTiled(output) failed");
    }
    {
        // the packed fp16 weights are moved as 128-byte rows of 32 "floats": the bulk copy and its swizzle are byte-wise
        cuuint64_t dims[2] = {32, (cuuint64_t)H_NPAIR * 64};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {32, 64};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wsplit, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SOL_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed");
    }
    HArgs a;
    a.trace = nullptr;
    if (g_h_trace && g_h_trace_seq < g_h_trace_cap) {
        a.trace = g_h_trace + (size_t)g_h_trace_seq * cdiv(X, H_TX) * cdiv(Y, H_TY) * B * 16;
        ++g_h_trace_seq;
    }
    a.weights_ready = weights_ready ? 1 : 0;
    a.amax_out = amax_out;
    a.w_inv_scale = wsplit + H_WPAIR_FLOATS;
    a.bias = bias; a.addend = addend; a.ref = ref; a.B = B; a.Y = Y; a.X = X; a.act = act; a.slope = slope;
    dim3 grid(cdiv(X, H_TX), cdiv(Y, H_TY), B);
    auto go = [&](auto kern, int stages) -> cudaError_t {
        const int smem = h_smem_bytes(stages);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);      // cheap; also safe per device
        if (e != cudaSuccess) return e;
        return launch_kernel(kern, grid, dim3(H_THREADS), (size_t)smem, st, map_in, map_w, map_out, a);
    };
    // measured on B200 at the bench shape (profiles/r02/r02_c_conv_variants.txt): the three accumulator layouts are within 1 %
    // (9.7-9.8 us per layer); deeper rings do not help; 3-4 CTAs per SM are SLOWER (12-14 us: programmatic dependent launch
    // packs the early CTAs of the next layer three deep on the same SMs, which then share one shared-memory port)
    switch (g_conv_variant) {
        case 1: SOL_CUDA(go(k_conv5x5_c32_h<2, false, 5, 2>, 5)); break;
        case 2: SOL_CUDA(go(k_conv5x5_c32_h<1, false, 5, 2>, 5)); break;
        default: SOL_CUDA(go(k_conv5x5_c32_h<2, true, 5, 2>, 5)); break;
    }
    SOL_LAUNCHED();
    return SOL_OK;
}

}  // namespace sol

// Diagnostics hook (not part of the public ABI): device buffer of `launches` x gridsize x 16 int64 that the next `launches`
// fp16-split convolution launches fill with clock64 / globaltimer phase stamps; pass null to switch tracing off.
extern "C" void sol_debug_conv_h_trace(long long* buf, int launches) {
    sol::g_h_trace = buf; sol::g_h_trace_cap = launches; sol::g_h_trace_seq = 0;
}
