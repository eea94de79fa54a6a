// 5x5 32->32 convolution on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32-accurate
// through 3xTF32 operand splitting.  Drop-in for k_conv5x5_c32 (same NHWC fp32 inputs/outputs and
// epilogue), used for the forward layers and — with flipped weights — the data gradient.
//
// Implicit GEMM per CTA:  D[128 pixels, 32 cout] += sum over 25 taps, 32 cin
//   * CTA tile = 8 (x) x 16 (y) output pixels -> UMMA M = 128, N = 32, K = 8 per instruction (tf32)
//   * ONE TMA box load brings the input halo tile [20 rows][12 px][32 ch] fp32 (30 KB, 128B-swizzled,
//     out-of-image pixels zero-filled by TMA = Keras 'same' padding).  A pixel is one 128-byte
//     swizzle row, the 8 pixels of a tile row are one 8-row core group, so the A operand of tap
//     (dy,dx) is the SAME shared-memory tile addressed with start = base + (dy*12+dx)*128 B and
//     stride-byte-offset = one halo row (1536 B): no im2col copy, 25x reuse of the staged tile.
//   * 3xTF32: the epilogue warps split the staged fp32 tile once into hi = rn_tf32(x) and
//     lo = rn_tf32(x - hi) (in place + a second 30 KB tile); weights are pre-split on the device
//     once per optimiser step.  D += Ahi*Bhi + Ahi*Blo + Alo*Bhi with fp32 accumulation in TMEM.
//   * weights stream per tap (8 KB, K-major, 128B-swizzled) through a 6-stage TMA/mbarrier ring.
//   * warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2..5 =
//     operand splitter, then epilogue (tcgen05.ld 32 lanes x 32 columns -> bias / residual /
//     LeakyReLU(-derivative) -> 128-byte row stores).
#include <string.h>

#include "sol_internal.cuh"
#include "sol_tc_common.cuh"

SOL_TRACE_TU()

namespace sol {

namespace {

constexpr int TC_TX = 8;        // tile width  (pixels)
constexpr int TC_TY = 16;       // tile height (pixels)
constexpr int TC_HW = 12;       // halo box width  (8 + 4); the swizzle phase is address-based, so any row pitch works
constexpr int TC_HH = 20;       // halo box height (16 + 4)
constexpr int TC_A_BYTES = TC_HH * TC_HW * 128;   // 30720
constexpr int TC_B_BYTES = 32 * 128;              // one tap, one half (hi or lo): 4096; a stage = hi + lo = 8192
constexpr int TC_STAGES = 6;      // 6 taps of weights in flight; 2 x 109 KB CTAs per SM
constexpr int TC_THREADS = 192;
constexpr int TC_NSET = 2;        // independent accumulator sets {[hh|hl] 64 cols, lh 32 cols}, used alternately
constexpr int TC_TMEM_COLS = 256; // 2 x 96 columns, power of two
// dynamic shared memory layout (offsets from a 1024-aligned base)
constexpr int TC_OFF_AHI = 0;
constexpr int TC_OFF_ALO = TC_A_BYTES;
constexpr int TC_OFF_B = 2 * TC_A_BYTES;                            // [stage][hi|lo][4096]
constexpr int TC_OFF_BAR = TC_OFF_B + TC_STAGES * 2 * TC_B_BYTES;   // mbarriers
constexpr int TC_SMEM = TC_OFF_BAR + 256 + 1024;                    // + alignment slack

struct TcArgs {
    int base_offset_mode;   // 1: descriptor base_offset = (start >> 7) & 7 (PTX rule); 0: leave it zero
    const float* bias;
    const float* addend;
    const float* ref;
    float* out;
    int B, Y, X;
    int act;
    float slope;
    int weights_ready;      // 1: the split weights were complete before the previous kernel of the stream started
    long long* trace;       // diagnostics: 16 slots per CTA of clock64 phase stamps (null in production), already offset
                            // to this launch's block of gridsize x 16 slots
};

using namespace tc;

__device__ __forceinline__ void tc_stamp(long long* trace, int slot) {
    if (trace) {
        const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        trace[cta * 16 + slot] = clock64();
    }
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2,
// K-major A and B (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t TC_IDESC32 = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t TC_IDESC64 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

}  // namespace

__global__ void __launch_bounds__(TC_THREADS, 2)
k_conv5x5_c32_tc(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_out, const TcArgs a) {
    
    extern __shared__ uint8_t tc_smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    const uint32_t raw = smem_u32(tc_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gbase = tc_smem_raw + (base - raw);
    const uint32_t s_ahi = base + TC_OFF_AHI, s_alo = base + TC_OFF_ALO, s_b = base + TC_OFF_B;
    const uint32_t s_bar = base + TC_OFF_BAR;
    const uint32_t bar_afull = s_bar + 0, bar_asplit = s_bar + 8, bar_acc = s_bar + 16;
    const uint32_t bar_bfull = s_bar + 32, bar_bempty = s_bar + 32 + 8 * TC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + TC_OFF_BAR + 32 + 16 * TC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x0 = blockIdx.x * TC_TX, y0 = blockIdx.y * TC_TY, b = blockIdx.z;
    // Every CTA streams the same 25 weight tiles; if all CTAs walked them in the same order the 200
    // co-resident CTAs would hammer the same few L2 lines at the same time (measured: ~1500 cycles per
    // tap).  Each CTA therefore starts at a different tap and wraps around.
    const int tap0 = (int)((blockIdx.x + blockIdx.y * gridDim.x + blockIdx.z * gridDim.x * gridDim.y) * 7u % 25u);

    if (threadIdx.x == 0 && a.trace) {
        const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        a.trace[cta * 16 + 0] = smid; a.trace[cta * 16 + 1] = (long long)gt;
        tc_stamp(a.trace, 2);
    }
    if (warp == 0 && lane == 0) {
        mbar_init(bar_afull, 1);
        mbar_init(bar_asplit, TC_THREADS);
        mbar_init(bar_acc, 1);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: TC_NACC accumulators of 32 fp32 columns x 128 lanes
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_slot;
    if (threadIdx.x == 0) tc_stamp(a.trace, 3);      // set-up done (barriers, TMEM)

    const int npre = a.weights_ready ? TC_STAGES : 0;
    if (warp == 0) {
        // ================= TMA producer, part 1 (warp-uniform control flow, one elected lane issues) =================
        const bool leader = elect_one();
        // Weights that were ready before the predecessor kernel started do not depend on it: fill the ring
        // before waiting for the predecessor (programmatic dependent launch), then fetch the halo tile.
        for (int n = 0; n < npre; ++n) {
            int tap = tap0 + n; if (tap >= 25) tap -= 25;
            if (leader) {
                mbar_arrive_expect_tx(bar_bfull + 8 * n, 2 * TC_B_BYTES);
                tma_load_2d(s_b + 2 * n * TC_B_BYTES, &map_w, bar_bfull + 8 * n, 0, tap * 64);
            }
        }
        pdl_wait();
        if (leader && a.trace) {
            const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
            unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            a.trace[cta * 16 + 13] = (long long)gt;      // dependency on the previous kernel resolved
        }
        if (leader) {
            mbar_arrive_expect_tx(bar_afull, TC_A_BYTES);
            tma_load_4d(s_ahi, &map_in, bar_afull, 0, x0 - 2, y0 - 2, b);
        }
    } else {
        pdl_wait();
    }

    // ================= operand split by ALL warps: hi = rn_tf32(x) in place, lo = rn_tf32(x - hi) =================
    mbar_wait(bar_afull, 0);
    // The next kernel of the stream may become resident from here on: this CTA no longer reads its input from
    // global memory, and every producer tile under its halo is complete (the first launch_dependents of a CTA counts).
    if (threadIdx.x == 64) pdl_trigger();
    if (threadIdx.x == 64) tc_stamp(a.trace, 4);            // halo tile landed
    {
        float4* hi4 = reinterpret_cast<float4*>(gbase + TC_OFF_AHI);
        float4* lo4 = reinterpret_cast<float4*>(gbase + TC_OFF_ALO);
        static_assert(TC_A_BYTES / 16 % TC_THREADS == 0, "split geometry");
#pragma unroll 5
        for (int it = 0; it < TC_A_BYTES / 16 / TC_THREADS; ++it) {      // 10 float4 per thread, 5 loads in flight
            const int i = (int)threadIdx.x + it * TC_THREADS;
            const float4 v = hi4[i];
            float4 h, l;
            h.x = tf32_rn(v.x); l.x = tf32_rn(v.x - h.x);
            h.y = tf32_rn(v.y); l.y = tf32_rn(v.y - h.y);
            h.z = tf32_rn(v.z); l.z = tf32_rn(v.z - h.z);
            h.w = tf32_rn(v.w); l.w = tf32_rn(v.w - h.w);
            hi4[i] = h;
            lo4[i] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
        mbar_arrive(bar_asplit);
    }
    if (threadIdx.x == 64) tc_stamp(a.trace, 5);            // split done (this thread)

    if (warp == 0) {
        // ================= TMA producer, part 2: keep the weight ring full =================
        const bool leader = elect_one();
#pragma unroll 1
        for (int n = npre; n < 25; ++n) {
            const int s = n % TC_STAGES;
            const uint32_t ph = (uint32_t)(n / TC_STAGES) & 1u;
            int tap = tap0 + n; if (tap >= 25) tap -= 25;
            mbar_wait(bar_bempty + 8 * s, ph ^ 1u);     // first pass: fresh barrier, parity 1 passes
            if (leader) {
                mbar_arrive_expect_tx(bar_bfull + 8 * s, 2 * TC_B_BYTES);
                tma_load_2d(s_b + 2 * s * TC_B_BYTES, &map_w, bar_bfull + 8 * s, 0, tap * 64);   // 32 hi rows + 32 lo rows
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // An N<=64 tf32 UMMA is only 16-32 cycles of tensor-pipe work, so the issue path matters: the
        // whole warp runs the (uniform) control flow and one elected lane issues; descriptors are a
        // per-tile constant plus a 16-byte-unit offset; [Bhi;Blo] are adjacent in the weight stage so
        // Ahi*Bhi and Ahi*Blo are ONE N=64 instruction ([hh|hl] accumulator) next to the N=32 Alo*Bhi.
        // Accumulations into the same TMEM tile serialise on the MMA latency and the tensor core
        // accumulates with truncation, so two independent accumulator sets are used alternately and
        // summed with RN fp32 adds in the epilogue.
        const bool leader = elect_one();
        const uint64_t dA_hi = make_desc(s_ahi, TC_HW * 128, 0);
        const uint64_t dA_lo = make_desc(s_alo, TC_HW * 128, 0);
        const uint64_t dB = make_desc(s_b, 1024, 0);
        mbar_wait(bar_asplit, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (leader) tc_stamp(a.trace, 6);            // operands split, MMA may start
#pragma unroll 1
        for (int n = 0; n < 25; ++n) {
            const int s = n % TC_STAGES;
            const uint32_t ph = (uint32_t)(n / TC_STAGES) & 1u;
            int tap = tap0 + n; if (tap >= 25) tap -= 25;
            mbar_wait(bar_bfull + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (leader) {
                if (n == 0) tc_stamp(a.trace, 7);    // first weight stage landed
                const int dy = tap / 5, dx = tap - dy * 5;
                const uint64_t a_off = (uint64_t)((dy * TC_HW + dx) * 8);            // 128 B rows in 16 B units
                const uint64_t b_off = (uint64_t)(s * (2 * TC_B_BYTES / 16));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t set = (uint32_t)(ks & 1);
                    const uint32_t acc1 = tmem_acc + set * 96u, acc2 = acc1 + 64u;
                    const uint32_t first = (n == 0 && ks < 2) ? 0u : 1u;
                    umma_tf32(acc1, dA_hi + a_off + 2 * ks, dB + b_off + 2 * ks, TC_IDESC64, first);   // [hh | hl]
                    umma_tf32(acc2, dA_lo + a_off + 2 * ks, dB + b_off + 2 * ks, TC_IDESC32, first);   // lh
                }
                umma_commit(bar_bempty + 8 * s);     // frees the weight slot when these MMAs retire
            }
            __syncwarp();
        }
        if (leader) { umma_commit(bar_acc); tc_stamp(a.trace, 8); }   // all MMAs issued
        __syncwarp();
    } else {
        // ================= splitter, then epilogue (warps 2..5 = 128 threads) =================
        const int t = threadIdx.x - 64;
        // ---- epilogue: TMEM lane = pixel row of the tile, 32 columns = cout ----
        const int q = warp & 3;                 // this warp may touch TMEM lanes [32q, 32q+32)
        const int r = q * 32 + lane;            // accumulator row = pixel
        const int gy = y0 + (r >> 3), gx = x0 + (r & 7);
        const bool inside = gy < a.Y && gx < a.X;
        const size_t o = (((size_t)b * a.Y + gy) * a.X + gx) * 32;
        // the residual / activation-reference rows are fetched while the tensor core is still busy
        float4 ad[8], rf[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { ad[c] = make_float4(0.f, 0.f, 0.f, 0.f); rf[c] = make_float4(1.f, 1.f, 1.f, 1.f); }
        if (inside) {
            if (a.addend) {
#pragma unroll
                for (int c = 0; c < 8; ++c) ad[c] = __ldg(reinterpret_cast<const float4*>(a.addend + o) + c);
            }
            if (a.act == SOL_ACT_DLRELU) {
#pragma unroll
                for (int c = 0; c < 8; ++c) rf[c] = __ldg(reinterpret_cast<const float4*>(a.ref + o) + c);
            }
        }
        if (a.bias) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias) + c);
                ad[c].x += bv.x; ad[c].y += bv.y; ad[c].z += bv.z; ad[c].w += bv.w;
            }
        }
        mbar_wait(bar_acc, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (t == 0) tc_stamp(a.trace, 9);            // accumulators complete
        float acc[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) { acc[4 * c] = ad[c].x; acc[4 * c + 1] = ad[c].y; acc[4 * c + 2] = ad[c].z; acc[4 * c + 3] = ad[c].w; }
        // six 32-column blocks (per set: hh, hl, lh), two loads in flight per wait
        auto tmem_ld32 = [&](uint32_t (&v)[32], int j) {
            const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + 32u * (uint32_t)j;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
        };
#pragma unroll 1
        for (int j = 0; j < 3 * TC_NSET; j += 2) {
            uint32_t v0[32], v1[32];
            tmem_ld32(v0, j);
            tmem_ld32(v1, j + 1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] += __uint_as_float(v0[c]) + __uint_as_float(v1[c]);
        }
        // Output tile -> shared memory (the A_lo tile is free once the accumulators are complete) in the 128B-swizzled
        // layout of the output tensor map, then ONE bulk tensor store per CTA: full 128-byte lines instead of 16-byte
        // pieces per thread, clipped at the image border by the TMA unit.
        {
            uint8_t* stage = gbase + TC_OFF_ALO + r * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 f = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
                if (a.act == SOL_ACT_LRELU) {
                    f.x = f.x > 0.f ? f.x : a.slope * f.x; f.y = f.y > 0.f ? f.y : a.slope * f.y;
                    f.z = f.z > 0.f ? f.z : a.slope * f.z; f.w = f.w > 0.f ? f.w : a.slope * f.w;
                } else if (a.act == SOL_ACT_DLRELU) {
                    f.x = rf[c].x > 0.f ? f.x : a.slope * f.x; f.y = rf[c].y > 0.f ? f.y : a.slope * f.y;
                    f.z = rf[c].z > 0.f ? f.z : a.slope * f.z; f.w = rf[c].w > 0.f ? f.w : a.slope * f.w;
                }
                *reinterpret_cast<float4*>(stage + ((c ^ (r & 7)) << 4)) = f;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> bulk store reads
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (t == 0) {
            tma_store_4d(&map_out, s_alo, 0, x0, y0, b);
            tma_store_commit();
            tma_store_wait_read();       // shared memory must stay valid until the bulk store has read it
        }
    }

    if (threadIdx.x == 64 && a.trace) {              // the thread that issued the bulk store: end of this CTA's useful work
        const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        a.trace[cta * 16 + 12] = (long long)gt;
        tc_stamp(a.trace, 10);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)TC_TMEM_COLS) : "memory");
    }
}

// wprep[tap][hi|lo][n][k]: hi / lo halves of  Bt[tap][n][k] = w[tap][k][n]   (w: Keras [5,5,K=Cin,N=Cout])
__global__ void __launch_bounds__(256) k_prep_tc_weights(const float* __restrict__ w, float* __restrict__ wprep) {
    pdl_sync();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 25 * 32 * 32) return;
    const int k = idx & 31, n = (idx >> 5) & 31, tap = idx >> 10;
    const float v = w[(tap * 32 + k) * 32 + n];
    const float h = tf32_rn(v);
    wprep[((tap * 2 + 0) * 32 + n) * 32 + k] = h;
    wprep[((tap * 2 + 1) * 32 + n) * 32 + k] = tf32_rn(v - h);
}

// ------------------------------------------------------------------------------------------------
namespace tc {
EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
}  // namespace tc

static long long* g_tc_trace = nullptr;
static int g_tc_trace_cap = 0, g_tc_trace_seq = 0;     // capacity in launches, launches traced so far
int g_tc_base_offset_mode = 0;   // measured on B200: the swizzle phase comes from the absolute address bits; base_offset stays 0
int g_conv_path = 2;    // default: tcgen05 3xFP16 convolutions (1 = fp32 SIMT validation kernels, 3 = tcgen05 3xTF32)
int g_wgrad_path = 2;   // default: deferred tcgen05 3xFP16 weight-gradient GEMM (1 = per-step fp32 SIMT, 3 = deferred tcgen05 3xTF32)

int tc_tiles_per_launch(int B, int Y, int X) { return cdiv(X, TC_TX) * cdiv(Y, TC_TY) * B; }

size_t tc_weights_floats() { return (size_t)2 * 25 * 32 * 32; }      // the larger of the two split layouts (3xTF32); 256-byte multiple

int launch_prep_tc_weights(cudaStream_t st, const float* w, float* wprep) {
    SOL_CUDA(launch_kernel(k_prep_tc_weights, dim3(cdiv(25 * 32 * 32, 256)), dim3(256), 0, st, w, wprep));
    SOL_LAUNCHED();
    return SOL_OK;
}

int launch_conv5x5_tc(cudaStream_t st, int B, int Y, int X, const float* in, const float* wprep, const float* bias,
                      const float* addend, const float* ref, int act, float slope, float* out, bool weights_ready) {
    tc::EncodeTiledFn enc = tc::get_encode_tiled();
    if (!enc) return fail(SOL_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (((uintptr_t)in & 15) || ((uintptr_t)wprep & 15)) return fail(SOL_ERR_INVALID, "conv tc: operands must be 16-byte aligned");
    if ((uintptr_t)out & 15) return fail(SOL_ERR_INVALID, "conv tc: output must be 16-byte aligned");
    alignas(64) CUtensorMap map_in, map_w, map_out;
    {
        cuuint64_t dims[4] = {32, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)B};
        cuuint64_t strides[3] = {128, (cuuint64_t)X * 128, (cuuint64_t)Y * X * 128};
        cuuint32_t box[4] = {32, TC_HW, TC_HH, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SOL_ERR_CUDA, "cuTensorMapEncodeTiled(activations) failed");
        cuuint32_t obox[4] = {32, TC_TX, TC_TY, 1};
        r = enc(&map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)out, dims, strides, obox, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SOL_ERR_CUDA, "cuTensorMapEncodeTiled(output) failed");
    }
    {
        cuuint64_t dims[2] = {32, 1600};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {32, 64};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wprep, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SOL_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed");
    }
    TcArgs a;
    a.base_offset_mode = g_tc_base_offset_mode;
    a.trace = nullptr;
    if (g_tc_trace && g_tc_trace_seq < g_tc_trace_cap) {
        a.trace = g_tc_trace + (size_t)g_tc_trace_seq * cdiv(X, TC_TX) * cdiv(Y, TC_TY) * B * 16;
        ++g_tc_trace_seq;
    }
    a.weights_ready = weights_ready ? 1 : 0;
    a.bias = bias; a.addend = addend; a.ref = ref; a.out = out; a.B = B; a.Y = Y; a.X = X; a.act = act; a.slope = slope;
    static bool attr_done = false;
    if (!attr_done) {
        SOL_CUDA(cudaFuncSetAttribute(k_conv5x5_c32_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
        attr_done = true;
    }
    dim3 grid(cdiv(X, TC_TX), cdiv(Y, TC_TY), B);
    SOL_CUDA(launch_kernel(k_conv5x5_c32_tc, grid, dim3(TC_THREADS), TC_SMEM, st, map_in, map_w, map_out, a));
    SOL_LAUNCHED();
    return SOL_OK;
}

// ---- path dispatch for the 32->32 layers: 1 = fp32 SIMT, 2 = tcgen05 3xFP16 (sol_conv_h.cu, default), 3 = tcgen05 3xTF32 ----
bool conv_path_is_tc() { return g_conv_path == 2 || g_conv_path == 3; }

int launch_split_weights(cudaStream_t st, const float* w, float* wsplit) {
    if (g_conv_path == 3) return launch_prep_tc_weights(st, w, wsplit);
    return launch_prep_h_weights(st, w, wsplit);
}

int launch_split_weights_multi(cudaStream_t st, const float* w, size_t w_stride, float* wsplit, int nlayers) {
    if (g_conv_path == 3) {
        for (int l = 0; l < nlayers; ++l) SOL_TRY(launch_prep_tc_weights(st, w + (size_t)l * w_stride, wsplit + (size_t)l * tc_weights_floats()));
        return SOL_OK;
    }
    return launch_prep_h_weights(st, w, wsplit, nlayers, w_stride, tc_weights_floats());
}

int launch_conv5x5_c32_presplit(cudaStream_t st, int B, int Y, int X, const float* in, const float* wsplit, const float* bias,
                                const float* addend, const float* ref, int act, float slope, float* out, bool weights_ready,
                                unsigned int* amax_out) {
    if (act == SOL_ACT_DLRELU && !ref) return fail(SOL_ERR_INVALID, "conv5x5: SOL_ACT_DLRELU needs ref");
    if (g_conv_path == 3) {
        if (amax_out) return fail(SOL_ERR_UNSUPPORTED, "conv5x5: the 3xTF32 kernel does not track max|out|");
        return launch_conv5x5_tc(st, B, Y, X, in, wsplit, bias, addend, ref, act, slope, out, weights_ready);
    }
    return launch_conv5x5_h(st, B, Y, X, in, wsplit, bias, addend, ref, act, slope, out, weights_ready, amax_out);
}

int launch_conv5x5_c32_auto(cudaStream_t st, int B, int Y, int X, const float* in, const float* w, const float* wprep,
                            const float* bias, const float* addend, const float* ref, int act, float slope, float* out,
                            unsigned int* amax_out) {
    if (!conv_path_is_tc()) {
        if (amax_out) return fail(SOL_ERR_UNSUPPORTED, "conv5x5: the SIMT 32->32 kernel does not track max|out|");
        return launch_conv5x5(st, B, Y, X, 32, 32, in, w, bias, addend, ref, act, slope, out);
    }
    const bool engine_weights = wprep != nullptr;   // the unrolled sweep splits all weights before its first step
    if (!wprep) {
        // stand-alone call: split the weights into a process-wide scratch buffer (stream-ordered reuse)
        static float* scratch = nullptr;
        if (!scratch) SOL_CUDA(cudaMalloc((void**)&scratch, tc_weights_floats() * sizeof(float)));
        SOL_TRY(launch_split_weights(st, w, scratch));
        wprep = scratch;
    }
    return launch_conv5x5_c32_presplit(st, B, Y, X, in, wprep, bias, addend, ref, act, slope, out, engine_weights, amax_out);
}

}  // namespace sol

// Diagnostics hook (not part of the public ABI): device buffer of `launches` x gridsize x 16 int64 that the next
// `launches` tensor-core convolution launches (also when captured into a CUDA graph) fill with clock64 /
// globaltimer phase stamps; pass null to switch tracing off.
extern "C" void sol_debug_conv_trace(long long* buf, int launches) {
    sol::g_tc_trace = buf; sol::g_tc_trace_cap = launches; sol::g_tc_trace_seq = 0;
}
