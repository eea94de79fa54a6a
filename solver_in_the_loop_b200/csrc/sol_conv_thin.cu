// First / last layers of the correction CNN and their data gradients: 5x5 'same' convolutions Cin <= 4 -> 32 and
// 32 -> Cout <= 4 (reference: karman-2d/karman_train.py:103 and :135, burgers/burgers_train.py:115-150), exact fp32.
//
// At the bench shape (3 x 128 x 64 = 24,576 pixels on 148 SMs) these layers are pure latency: 59 / 39 MFMA per launch
// are 1.6 / 1.1 us of FP32 issue for the whole GPU, but the first-generation kernels (sol_conv.cu: one pixel x 32 couts
// per thread, 5 warps per SM, 9 shared loads per 32 FMAs) measured 12 - 30 us inside the replayed iteration
// (profiles/r02/r02_n_chain_trace_sol32.txt).  These kernels trade registers for parallelism instead:
//   * a CTA is P pairs of image rows x 32 columns; the four warps of a row pair split the 32 output channels (expand) or the
//     32 input channels (reduce), so an SM holds 10 - 12 warps of independent work instead of 5;
//   * P is chosen so that the grid fits ONE wave with ONE CTA per SM (132 CTAs of 12 warps at the bench shape), and that is
//     enforced by a shared-memory request of more than half an SM: with programmatic dependent launch the CTAs become
//     resident while the PREVIOUS kernel is still running, wherever there is room at that moment - behind k_direct_apply
//     (129 CTAs x 1024 threads) the 19 idle SMs absorbed most of the 384 small CTAs of the first version and the layer took
//     15 us instead of 7 (scripts/chain_trace.py with fuse_solver_io = 0 / 1).  The other half of the SM stays free for the
//     early CTAs of the next kernel;
//   * lane = x (conflict-free shared reads, coalesced global traffic), a thread owns the two vertically adjacent pixels:
//     one 6-row column of inputs and one set of warp-uniform 128-bit weight loads feed 5 taps x 2 pixels;
//   * all global loads of a thread are issued before the first shared store; one barrier; the reduce kernel sums its four
//     channel groups through shared memory in a fixed order (deterministic).
#include "sol_internal.cuh"

SOL_TRACE_TU()

namespace sol {

namespace {

struct ThinArgs {
    const float* in;
    const float* w;
    const float* bias;
    const float* addend;
    const float* ref;
    float* out;
    int B, Y, X;
    int act;
    float slope;
    unsigned int* amax_out;
    int weights_ready;          // 1: the weights were complete before the previous kernel of the stream started (fetch them before the wait)
};

__device__ __forceinline__ float thin_act(float v, int act, float slope, float ref) {
    if (act == SOL_ACT_LRELU) return v > 0.0f ? v : slope * v;
    if (act == SOL_ACT_DLRELU) return ref > 0.0f ? v : slope * v;
    return v;
}

constexpr int TW = 32, PW = TW + 4;

// Asynchronous global -> shared copies (LDGSTS): no staging registers, every copy of a thread is in flight at once (ptxas
// interleaves register-staged loads and stores in groups of four, which serialises the L2 latency); `ok` = false zero-fills.
__device__ __forceinline__ uint32_t thin_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async_4(void* dst, const float* src, bool ok) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(thin_smem_u32(dst)), "l"(src), "r"(ok ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* dst, const void* src, bool ok) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(thin_smem_u32(dst)), "l"(src), "r"(ok ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

// ---- Cin <= 4 -> 32.  warp = (row pair, 8 output channels), lane = x, thread = pixels (y, x) and (y + 1, x) ----
template <int CIN, int P>
__global__ void __launch_bounds__(128 * P) k_conv5x5_expand2(const ThinArgs a) {
    constexpr int NT = 128 * P, TR = 2 * P, PH = TR + 4;
    __shared__ float4 ws4[25 * CIN * 8];            // [tap][ci][32 couts]
    __shared__ float tin[CIN * PH * PW];            // planar [ci][row][px]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cg = warp & 3, rp = warp >> 2;        // channel group, row pair
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TR, b = blockIdx.z;
    constexpr int TOTALW = 25 * CIN * 8, ITERW = (TOTALW + NT - 1) / NT;
    constexpr int TOTAL = PH * PW * CIN, ITER = (TOTAL + NT - 1) / NT;
    const float4* wg = reinterpret_cast<const float4*>(a.w);
    auto fetch_weights = [&]() {
#pragma unroll
        for (int it = 0; it < ITERW; ++it) { const int idx = tid + it * NT; if (idx < TOTALW) cp_async_16(ws4 + idx, wg + idx, true); }
    };
    if (a.weights_ready) fetch_weights();      // settled weights do not depend on the predecessor kernel: fetch them before waiting for it
    pdl_sync();      // (triggering the dependent kernel only after the FMA loops measured slower: 11.75 vs 11.44 ms per iteration)
    if (!a.weights_ready) fetch_weights();
    {
        const float* inb = a.in + (size_t)b * a.Y * a.X * CIN;
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int idx = tid + it * NT;
            const int row = idx / (PW * CIN), rem = idx - row * (PW * CIN);     // [row][px][ci] order = global order
            const int px = rem / CIN, ci = rem - px * CIN;
            const int gy = y0 + row - 2, gx = x0 + px - 2;
            const bool ok = gy >= 0 && gy < a.Y && gx >= 0 && gx < a.X;
            if (idx < TOTAL) cp_async_4(tin + (ci * PH + row) * PW + px, ok ? inb + ((size_t)gy * a.X + gx) * CIN + ci : inb, ok);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    float acc[2][8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { acc[0][c] = 0.0f; acc[1][c] = 0.0f; }
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
            float v[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) v[r] = tin[(ci * PH + 2 * rp + r) * PW + lane + dx];
#pragma unroll
            for (int dy = 0; dy < 5; ++dy) {
                const float4* wq = ws4 + ((dy * 5 + dx) * CIN + ci) * 8 + cg * 2;
                const float4 w0 = wq[0], w1 = wq[1];
#pragma unroll
                for (int pz = 0; pz < 2; ++pz) {
                    const float u = v[dy + pz];
                    acc[pz][0] = fmaf(u, w0.x, acc[pz][0]); acc[pz][1] = fmaf(u, w0.y, acc[pz][1]);
                    acc[pz][2] = fmaf(u, w0.z, acc[pz][2]); acc[pz][3] = fmaf(u, w0.w, acc[pz][3]);
                    acc[pz][4] = fmaf(u, w1.x, acc[pz][4]); acc[pz][5] = fmaf(u, w1.y, acc[pz][5]);
                    acc[pz][6] = fmaf(u, w1.z, acc[pz][6]); acc[pz][7] = fmaf(u, w1.w, acc[pz][7]);
                }
            }
        }
    }
    const int gx = x0 + lane;
    unsigned int amax = 0u;
    float4 bv0 = make_float4(0.f, 0.f, 0.f, 0.f), bv1 = bv0;
    if (a.bias) { bv0 = __ldg(reinterpret_cast<const float4*>(a.bias) + cg * 2); bv1 = __ldg(reinterpret_cast<const float4*>(a.bias) + cg * 2 + 1); }
#pragma unroll
    for (int pz = 0; pz < 2; ++pz) {
        const int gy = y0 + 2 * rp + pz;
        if (gy < a.Y && gx < a.X) {
            const size_t o4 = (((size_t)b * a.Y + gy) * a.X + gx) * 8 + cg * 2;
            float4 f0 = make_float4(acc[pz][0] + bv0.x, acc[pz][1] + bv0.y, acc[pz][2] + bv0.z, acc[pz][3] + bv0.w);
            float4 f1 = make_float4(acc[pz][4] + bv1.x, acc[pz][5] + bv1.y, acc[pz][6] + bv1.z, acc[pz][7] + bv1.w);
            if (a.addend) {
                const float4 d0 = __ldg(reinterpret_cast<const float4*>(a.addend) + o4), d1 = __ldg(reinterpret_cast<const float4*>(a.addend) + o4 + 1);
                f0.x += d0.x; f0.y += d0.y; f0.z += d0.z; f0.w += d0.w; f1.x += d1.x; f1.y += d1.y; f1.z += d1.z; f1.w += d1.w;
            }
            float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
            if (a.act == SOL_ACT_DLRELU) { r0 = __ldg(reinterpret_cast<const float4*>(a.ref) + o4); r1 = __ldg(reinterpret_cast<const float4*>(a.ref) + o4 + 1); }
            f0.x = thin_act(f0.x, a.act, a.slope, r0.x); f0.y = thin_act(f0.y, a.act, a.slope, r0.y);
            f0.z = thin_act(f0.z, a.act, a.slope, r0.z); f0.w = thin_act(f0.w, a.act, a.slope, r0.w);
            f1.x = thin_act(f1.x, a.act, a.slope, r1.x); f1.y = thin_act(f1.y, a.act, a.slope, r1.y);
            f1.z = thin_act(f1.z, a.act, a.slope, r1.z); f1.w = thin_act(f1.w, a.act, a.slope, r1.w);
            float4* out4 = reinterpret_cast<float4*>(a.out) + o4;
            out4[0] = f0; out4[1] = f1;
            amax = max(amax, max(max(__float_as_uint(f0.x) & 0x7fffffffu, __float_as_uint(f0.y) & 0x7fffffffu),
                                 max(__float_as_uint(f0.z) & 0x7fffffffu, __float_as_uint(f0.w) & 0x7fffffffu)));
            amax = max(amax, max(max(__float_as_uint(f1.x) & 0x7fffffffu, __float_as_uint(f1.y) & 0x7fffffffu),
                                 max(__float_as_uint(f1.z) & 0x7fffffffu, __float_as_uint(f1.w) & 0x7fffffffu)));
        }
    }
    if (a.amax_out) {       // one atomic per warp (warp-uniform branch: every lane reaches the reduction)
        amax = __reduce_max_sync(0xffffffffu, amax);
        if (lane == 0 && amax) atomicMax(a.amax_out, amax);
    }
}

// ---- 32 -> Cout <= 4.  warp = (row pair, 8 input channels), lane = x, thread = pixels (y, x) and (y + 1, x); the four channel
//      groups of a row pair are summed through shared memory by its first warp in group order ----
constexpr int PS = 36;                              // padded pixel stride (floats): 128-bit reads of consecutive pixels are conflict-free
template <int COUT, int P>
constexpr int reduce2_smem_floats() { return 25 * 32 * COUT + (2 * P + 4) * PW * PS; }

template <int COUT, int P>
__global__ void __launch_bounds__(128 * P) k_conv5x5_reduce2(const ThinArgs a) {
    constexpr int NT = 128 * P, TR = 2 * P, PH = TR + 4;
    extern __shared__ float4 thin_smem4[];
    float* ws = reinterpret_cast<float*>(thin_smem4);           // [25][32][COUT] (the Keras order)
    float* tin = ws + 25 * 32 * COUT;                           // [PH * PW][PS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cg = warp & 3, rp = warp >> 2;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TR, b = blockIdx.z;
    constexpr int TOTALW = 25 * 32 * COUT / 4, ITERW = (TOTALW + NT - 1) / NT;
    constexpr int TOTAL = PH * PW * 8, ITER = (TOTAL + NT - 1) / NT;
    const float4* wg = reinterpret_cast<const float4*>(a.w);
    auto fetch_weights = [&]() {
#pragma unroll
        for (int it = 0; it < ITERW; ++it) { const int idx = tid + it * NT; if (idx < TOTALW) cp_async_16(thin_smem4 + idx, wg + idx, true); }
    };
    if (a.weights_ready) fetch_weights();
    pdl_sync();
    if (!a.weights_ready) fetch_weights();
    {
        const float* inb = a.in + (size_t)b * a.Y * a.X * 32;
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int idx = tid + it * NT;
            const int c4 = idx & 7, pix = idx >> 3;
            const int row = pix / PW, px = pix - row * PW;
            const int gy = y0 + row - 2, gx = x0 + px - 2;
            const bool ok = gy >= 0 && gy < a.Y && gx >= 0 && gx < a.X;
            if (idx < TOTAL) cp_async_16(tin + pix * PS + c4 * 4, ok ? inb + ((size_t)gy * a.X + gx) * 32 + c4 * 4 : inb, ok);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    float acc[2][COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) { acc[0][c] = 0.0f; acc[1][c] = 0.0f; }
#pragma unroll
    for (int cq = 0; cq < 2; ++cq) {
        const int c4 = cg * 2 + cq;                             // float4 chunk of input channels
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
            float4 iv[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) iv[r] = *reinterpret_cast<const float4*>(tin + ((2 * rp + r) * PW + lane + dx) * PS + c4 * 4);
#pragma unroll
            for (int dy = 0; dy < 5; ++dy) {
                float wq[4 * COUT];                             // [ci 0..3][co]: 4*COUT consecutive floats, warp-uniform
                const float4* wp = reinterpret_cast<const float4*>(ws + ((dy * 5 + dx) * 32 + c4 * 4) * COUT);
#pragma unroll
                for (int q = 0; q < COUT; ++q) {
                    const float4 w = wp[q];
                    wq[4 * q] = w.x; wq[4 * q + 1] = w.y; wq[4 * q + 2] = w.z; wq[4 * q + 3] = w.w;
                }
#pragma unroll
                for (int pz = 0; pz < 2; ++pz) {
                    const float4 u = iv[dy + pz];
#pragma unroll
                    for (int co = 0; co < COUT; ++co)
                        acc[pz][co] = fmaf(u.x, wq[co], fmaf(u.y, wq[COUT + co], fmaf(u.z, wq[2 * COUT + co], fmaf(u.w, wq[3 * COUT + co], acc[pz][co]))));
                }
            }
        }
    }
    __syncthreads();                        // everyone is done with the input tile: reuse it for the group sums
    float* red = tin + rp * (3 * 2 * COUT * 32);      // [3 groups][2 * COUT][32 lanes] per row pair
    if (cg > 0) {
#pragma unroll
        for (int pz = 0; pz < 2; ++pz)
#pragma unroll
            for (int co = 0; co < COUT; ++co) red[((cg - 1) * 2 * COUT + pz * COUT + co) * 32 + lane] = acc[pz][co];
    }
    __syncthreads();
    if (cg > 0) return;
    const int gx = x0 + lane;
#pragma unroll
    for (int pz = 0; pz < 2; ++pz) {
        const int gy = y0 + 2 * rp + pz;
        if (gy >= a.Y || gx >= a.X) continue;
        const size_t o = (((size_t)b * a.Y + gy) * a.X + gx) * COUT;
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            float v = acc[pz][co];
#pragma unroll
            for (int g = 0; g < 3; ++g) v += red[(g * 2 * COUT + pz * COUT + co) * 32 + lane];
            if (a.bias) v += __ldg(a.bias + co);
            if (a.addend) v += __ldg(a.addend + o + co);
            const float rf = (a.act == SOL_ACT_DLRELU) ? __ldg(a.ref + o + co) : 0.0f;
            a.out[o + co] = thin_act(v, a.act, a.slope, rf);
        }
    }
}

// P = row pairs per CTA: the smallest that lets the grid fit one wave (one CTA per SM); then the shared-memory request is raised
// above half an SM so that the early (programmatic) placement cannot put two CTAs on one SM.  Larger grids: P = 4, no padding.
struct ThinGeom { int P; bool one_wave; };
int g_thin_cap_mask = 7;        // tuning: bit 0 expand layers, bit 1 reduce layers with Cout = 2 (forward), bit 2 other reduce layers
ThinGeom thin_geom(int B, int Y, int X) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    for (int P = 1; P <= 4; ++P)
        if ((long)cdiv(X, TW) * cdiv(Y, 2 * P) * B <= sms) return ThinGeom{P, true};
    return ThinGeom{4, false};
}
// Shared-memory split.  An SM keeps the L1 / shared-memory carve-out of the kernel whose CTAs arrived first, and the early CTAs of the
// NEXT kernel can only join if their shared memory still fits that split.  The reduce kernels are followed (adjoint sweep: directly)
// by the cluster kernel of the direct solver, 112 KB per CTA: they ask for the maximum carve-out so that it becomes resident beside
// them (15.1 -> 12.9 us).  The expand kernels must NOT: the hidden conv layers behind them would inherit the split for the whole
// chain and lose more (9.9 -> 10.3 us per layer; profiles/r02/r02_v_*, r02_w_*).
constexpr size_t HALF_SM_PLUS = 114 * 1024;      // static + dynamic request that allows one CTA per SM only (2 x (114 + 1) KB > 228 KB per SM) and
                                                 // still leaves room for one 112 KB CTA of the direct solver (the adjoint sweep runs it next)

template <int CIN, int P>
int launch_expand2_p(const ThinArgs& a, cudaStream_t st, bool one_wave) {
    constexpr size_t kStatic = (size_t)25 * CIN * 8 * sizeof(float4) + (size_t)CIN * (2 * P + 4) * PW * sizeof(float);
    const size_t smem = one_wave ? HALF_SM_PLUS - kStatic : 0;
    static size_t attr_bytes = 0;      // one process per GPU; raised by the first (eager, never captured) call of a shape
    if (smem > attr_bytes) {
        SOL_CUDA(cudaFuncSetAttribute(k_conv5x5_expand2<CIN, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_bytes = smem;
    }
    SOL_CUDA(launch_kernel(k_conv5x5_expand2<CIN, P>, dim3(cdiv(a.X, TW), cdiv(a.Y, 2 * P), a.B), dim3(128 * P), smem, st, a));
    SOL_LAUNCHED();
    return SOL_OK;
}

template <int CIN>
int launch_expand2(const ThinArgs& a, cudaStream_t st) {
    ThinGeom g = thin_geom(a.B, a.Y, a.X);
    g.one_wave = g.one_wave && (g_thin_cap_mask & 1);
    switch (g.P) {
        case 1: return launch_expand2_p<CIN, 1>(a, st, g.one_wave);
        case 2: return launch_expand2_p<CIN, 2>(a, st, g.one_wave);
        case 3: return launch_expand2_p<CIN, 3>(a, st, g.one_wave);
        default: return launch_expand2_p<CIN, 4>(a, st, g.one_wave);
    }
}

template <int COUT, int P>
int launch_reduce2_p(const ThinArgs& a, cudaStream_t st, bool one_wave) {
    constexpr size_t need = (size_t)reduce2_smem_floats<COUT, P>() * sizeof(float);
    static_assert(need <= HALF_SM_PLUS, "shared-memory budget");
    const size_t smem = one_wave ? HALF_SM_PLUS : need;
    static size_t attr_bytes = 0;
    if (smem > attr_bytes) {
        SOL_CUDA(cudaFuncSetAttribute(k_conv5x5_reduce2<COUT, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_bytes = smem;
    }
    static bool carve_done = false;
    if (!carve_done) {
        SOL_CUDA(cudaFuncSetAttribute(k_conv5x5_reduce2<COUT, P>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        carve_done = true;
    }
    SOL_CUDA(launch_kernel(k_conv5x5_reduce2<COUT, P>, dim3(cdiv(a.X, TW), cdiv(a.Y, 2 * P), a.B), dim3(128 * P), smem, st, a));
    SOL_LAUNCHED();
    return SOL_OK;
}

template <int COUT>
int launch_reduce2(const ThinArgs& a, cudaStream_t st) {
    ThinGeom g = thin_geom(a.B, a.Y, a.X);
    g.one_wave = g.one_wave && (g_thin_cap_mask & (COUT == 2 ? 2 : 4));
    switch (g.P) {
        case 1: return launch_reduce2_p<COUT, 1>(a, st, g.one_wave);
        case 2: return launch_reduce2_p<COUT, 2>(a, st, g.one_wave);
        case 3: return launch_reduce2_p<COUT, 3>(a, st, g.one_wave);
        default: return launch_reduce2_p<COUT, 4>(a, st, g.one_wave);
    }
}

}  // namespace

void set_thin_cap_mask(int m) { g_thin_cap_mask = m; }
int g_thin_path = 0;        // option "thin_path": 0 = the row-pair kernels of this file, 1 = the first-generation kernels of sol_conv.cu

// Returns SOL_ERR_UNSUPPORTED for channel counts this file does not cover (the caller falls back to the generic kernel).
int launch_conv5x5_thin(cudaStream_t st, int B, int Y, int X, int Cin, int Cout, const float* in, const float* w, const float* bias,
                        const float* addend, const float* ref, int act, float slope, float* out, unsigned int* amax_out, bool weights_ready) {
    ThinArgs a;
    a.weights_ready = weights_ready ? 1 : 0;
    a.in = in; a.w = w; a.bias = bias; a.addend = addend; a.ref = ref; a.out = out;
    a.B = B; a.Y = Y; a.X = X; a.act = act; a.slope = slope; a.amax_out = amax_out;
    if (Cout == 32 && Cin == 2) return launch_expand2<2>(a, st);
    if (Cout == 32 && Cin == 3) return launch_expand2<3>(a, st);
    if (Cout == 32 && Cin == 4) return launch_expand2<4>(a, st);
    if (Cin == 32 && Cout == 2) return launch_reduce2<2>(a, st);
    if (Cin == 32 && Cout == 3) return launch_reduce2<3>(a, st);
    if (Cin == 32 && Cout == 4) return launch_reduce2<4>(a, st);
    return SOL_ERR_UNSUPPORTED;
}

}  // namespace sol
