// 5x5 'same' convolutions of the correction CNN (keras Conv2D, reference
// karman-2d/karman_train.py:101-138), NHWC activations, Keras kernel layout [5,5,Cin,Cout].
//
// This file holds the fp32 SIMT kernels: the exact-fp32 baseline every other conv path in the
// repo is validated against.
//   k_conv5x5_c32   32->32 layers (10 of the 12 layers; also the data gradient via flipped weights)
//                   register tile 8 pixels x 4 couts per thread, input tile + halo in shared
//                   memory (float4 over cin), weights streamed through L1 (100 KB, L1-resident),
//                   5-tap sliding window along x so each shared load feeds 20+ FMAs.
//   k_conv5x5_expand / k_conv5x5_reduce   first / last layers (Cin <= 4 -> 32, 32 -> Cout <= 4)
//   k_wgrad_c32     weight gradient of the 32->32 layers: every thread owns 80 (tap,cin,cout)
//                   accumulators in registers for the whole pixel range of its CTA; per-CTA partial
//                   sums go to a private slot (no atomics, deterministic), reduced by
//                   k_wgrad_finalize once per optimiser step.
//   k_wgrad_expand / k_wgrad_reduce   weight gradients of the thin layers (register-tiled, persistent)
#include "sol_internal.cuh"

SOL_TRACE_TU()

namespace sol {

constexpr int THIN_SLOT = 25 * 4 * 32 + 32;      // floats of one CTA's partial slot in the deterministic thin weight-gradient mode

struct ConvArgs {
    const float* in;
    const float* w;
    const float* bias;
    const float* addend;
    const float* ref;
    float* out;
    int B, Y, X;
    int act;
    float slope;
    unsigned int* amax_out;     // optional: running max|out| (bit pattern) of the output tensor, for the fp16 weight-gradient GEMM
};

__device__ __forceinline__ float apply_act(float v, int act, float slope, float ref) {
    if (act == SOL_ACT_LRELU) return v > 0.0f ? v : slope * v;
    if (act == SOL_ACT_DLRELU) return ref > 0.0f ? v : slope * v;
    return v;
}

// ------------------------------------------------------------------------------------------------
// 32 -> 32
// ------------------------------------------------------------------------------------------------
template <int TH>
__global__ void __launch_bounds__(TH * 32, (TH <= 4 ? 3 : 2)) k_conv5x5_c32(const ConvArgs a) {
    pdl_sync();
    constexpr int TW = 32, PW = TW + 4, PH = TH + 4, C = 32;
    extern __shared__ float4 tile4[];   // [PH][PW][8]
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
    const float* inb = a.in + (size_t)b * a.Y * a.X * C;
    for (int idx = tid; idx < PH * PW * 8; idx += TH * 32) {
        const int c4 = idx & 7, pix = idx >> 3;
        const int tyy = pix / PW, txx = pix - tyy * PW;
        const int gy = y0 + tyy - 2, gx = x0 + txx - 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gy >= 0 && gy < a.Y && gx >= 0 && gx < a.X)
            v = __ldg(reinterpret_cast<const float4*>(inb + ((size_t)gy * a.X + gx) * C) + c4);
        tile4[idx] = v;
    }
    __syncthreads();

    const int cog = tid & 7, pg = tid >> 3, row = pg >> 2, xseg = pg & 3;
    float acc[8][4];
#pragma unroll
    for (int p = 0; p < 8; ++p) { acc[p][0] = acc[p][1] = acc[p][2] = acc[p][3] = 0.0f; }
    const float4* w4 = reinterpret_cast<const float4*>(a.w);   // [(tap*32 + ci)*8 + cog]

#pragma unroll 1
    for (int dy = 0; dy < 5; ++dy) {
        const float4* trow = tile4 + ((row + dy) * PW + xseg * 8) * 8;
#pragma unroll 1
        for (int c4 = 0; c4 < 8; ++c4) {
            float4 iv[12];
#pragma unroll
            for (int t = 0; t < 12; ++t) iv[t] = trow[t * 8 + c4];
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) {
                const float4* wp = w4 + ((dy * 5 + dx) * 32 + c4 * 4) * 8 + cog;
                const float4 w0 = __ldg(wp), w1 = __ldg(wp + 8), w2 = __ldg(wp + 16), w3 = __ldg(wp + 24);
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const float4 v = iv[p + dx];
                    acc[p][0] = fmaf(v.x, w0.x, fmaf(v.y, w1.x, fmaf(v.z, w2.x, fmaf(v.w, w3.x, acc[p][0]))));
                    acc[p][1] = fmaf(v.x, w0.y, fmaf(v.y, w1.y, fmaf(v.z, w2.y, fmaf(v.w, w3.y, acc[p][1]))));
                    acc[p][2] = fmaf(v.x, w0.z, fmaf(v.y, w1.z, fmaf(v.z, w2.z, fmaf(v.w, w3.z, acc[p][2]))));
                    acc[p][3] = fmaf(v.x, w0.w, fmaf(v.y, w1.w, fmaf(v.z, w2.w, fmaf(v.w, w3.w, acc[p][3]))));
                }
            }
        }
    }

    const int gy = y0 + row;
    if (gy >= a.Y) return;
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.bias) bv = __ldg(reinterpret_cast<const float4*>(a.bias) + cog);
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const int gx = x0 + xseg * 8 + p;
        if (gx >= a.X) continue;
        const size_t o4 = (((size_t)b * a.Y + gy) * a.X + gx) * 8 + cog;
        float4 v = make_float4(acc[p][0] + bv.x, acc[p][1] + bv.y, acc[p][2] + bv.z, acc[p][3] + bv.w);
        if (a.addend) {
            const float4 ad = __ldg(reinterpret_cast<const float4*>(a.addend) + o4);
            v.x += ad.x; v.y += ad.y; v.z += ad.z; v.w += ad.w;
        }
        float4 rf = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.act == SOL_ACT_DLRELU) rf = __ldg(reinterpret_cast<const float4*>(a.ref) + o4);
        v.x = apply_act(v.x, a.act, a.slope, rf.x);
        v.y = apply_act(v.y, a.act, a.slope, rf.y);
        v.z = apply_act(v.z, a.act, a.slope, rf.z);
        v.w = apply_act(v.w, a.act, a.slope, rf.w);
        reinterpret_cast<float4*>(a.out)[o4] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// thin layers.  Both kernels are issue-bound (shared-memory loads + FMAs), so the weights are read
// as uniform 128-bit shared loads and tiles are small enough that the bench grid gives 192 CTAs.
// ------------------------------------------------------------------------------------------------
// Cin <= 4 -> 32: one output pixel x 32 couts per thread, tile 32 (x) x 4 (y), one warp per tile row.
// Input tile is planar [ci][8][36] (conflict-free scalar reads), weights [tap][ci][32] as float4.
template <int CIN>
__global__ void __launch_bounds__(128) k_conv5x5_expand(const ConvArgs a) {
    pdl_sync();
    constexpr int TW = 32, TH = 4, PW = TW + 4, PH = TH + 4, COUT = 32;
    __shared__ float4 ws4[25 * CIN * 8];
    __shared__ float tin[CIN * PH * PW];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
    const float* inb = a.in + (size_t)b * a.Y * a.X * CIN;
    {
        // all global loads of a thread are issued before the first shared store (compile-time trip counts)
        constexpr int TOTAL = PH * PW * CIN, ITER = (TOTAL + 127) / 128;
        float v[ITER];
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int idx = tid + it * 128;
            const int row = idx / (PW * CIN), rem = idx - row * (PW * CIN);     // [row][px][ci] order = global order
            const int px = rem / CIN;
            const int gy = y0 + row - 2, gx = x0 + px - 2;
            v[it] = 0.0f;
            if (idx < TOTAL && gy >= 0 && gy < a.Y && gx >= 0 && gx < a.X) v[it] = __ldg(inb + ((size_t)gy * a.X + gx) * CIN + (rem - px * CIN));
        }
        constexpr int TOTALW = 25 * CIN * 8, ITERW = (TOTALW + 127) / 128;
        float4 wv[ITERW];
        const float4* wg = reinterpret_cast<const float4*>(a.w);
#pragma unroll
        for (int it = 0; it < ITERW; ++it) { const int idx = tid + it * 128; wv[it] = (idx < TOTALW) ? __ldg(wg + idx) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int idx = tid + it * 128;
            if (idx < TOTAL) {
                const int row = idx / (PW * CIN), rem = idx - row * (PW * CIN);
                const int px = rem / CIN, ci = rem - px * CIN;
                tin[(ci * PH + row) * PW + px] = v[it];
            }
        }
#pragma unroll
        for (int it = 0; it < ITERW; ++it) { const int idx = tid + it * 128; if (idx < TOTALW) ws4[idx] = wv[it]; }
    }
    __syncthreads();
    const int tx = tid & 31, ty = tid >> 5;
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.0f;
#pragma unroll 1
    for (int dy = 0; dy < 5; ++dy) {
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
                const float v = tin[(ci * PH + ty + dy) * PW + tx + dx];
                const float4* wq = ws4 + ((dy * 5 + dx) * CIN + ci) * 8;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 w = wq[q];
                    acc[4 * q + 0] = fmaf(v, w.x, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(v, w.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(v, w.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(v, w.w, acc[4 * q + 3]);
                }
            }
        }
    }
    const int gy = y0 + ty, gx = x0 + tx;
    const bool inside = gy < a.Y && gx < a.X;
    unsigned int amax = 0u;
    if (inside) {
        const size_t o4 = (((size_t)b * a.Y + gy) * a.X + gx) * 8;
        float4* out4 = reinterpret_cast<float4*>(a.out) + o4;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float4 f = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
            if (a.bias) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias) + q);
                f.x += bv.x; f.y += bv.y; f.z += bv.z; f.w += bv.w;
            }
            if (a.addend) {
                const float4 ad = __ldg(reinterpret_cast<const float4*>(a.addend) + o4 + q);
                f.x += ad.x; f.y += ad.y; f.z += ad.z; f.w += ad.w;
            }
            float4 rf = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.act == SOL_ACT_DLRELU) rf = __ldg(reinterpret_cast<const float4*>(a.ref) + o4 + q);
            f.x = apply_act(f.x, a.act, a.slope, rf.x); f.y = apply_act(f.y, a.act, a.slope, rf.y);
            f.z = apply_act(f.z, a.act, a.slope, rf.z); f.w = apply_act(f.w, a.act, a.slope, rf.w);
            out4[q] = f;
            amax = max(amax, max(max(__float_as_uint(f.x) & 0x7fffffffu, __float_as_uint(f.y) & 0x7fffffffu),
                                 max(__float_as_uint(f.z) & 0x7fffffffu, __float_as_uint(f.w) & 0x7fffffffu)));
        }
    }
    if (a.amax_out) {       // one atomic per warp (warp-uniform branch: every lane reaches the reduction)
        amax = __reduce_max_sync(0xffffffffu, amax);
        if ((tid & 31) == 0 && amax) atomicMax(a.amax_out, amax);
    }
}

// 32 -> Cout <= 4: tile 16 (x) x 8 (y), 128 threads.  Each half of the CTA reduces 16 of the 32 input
// channels; a thread owns two vertically adjacent pixels so that one 6-row column of float4 inputs and
// one set of weights feed 5 taps x 2 pixels.  The halves are summed through shared memory.
template <int COUT>
__global__ void __launch_bounds__(128) k_conv5x5_reduce(const ConvArgs a) {
    pdl_sync();
    constexpr int TW = 16, TH = 8, PW = TW + 4, PH = TH + 4, C = 32, PS = 36;   // PS: padded pixel stride (floats)
    extern __shared__ float4 smem4[];
    float* ws = reinterpret_cast<float*>(smem4);                 // [25][32][COUT]
    float* tin = ws + 25 * C * COUT;                             // [PH*PW][PS]
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
    const float* inb = a.in + (size_t)b * a.Y * a.X * C;
    {
        constexpr int TOTAL = PH * PW * 8, ITER = TOTAL / 128;   // 1920 float4 = 15 per thread
        static_assert(TOTAL % 128 == 0, "staging geometry");
        float4 v[ITER];
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int idx = tid + it * 128;
            const int c4 = idx & 7, pix = idx >> 3;
            const int row = pix / PW, px = pix - row * PW;
            const int gy = y0 + row - 2, gx = x0 + px - 2;
            v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy >= 0 && gy < a.Y && gx >= 0 && gx < a.X) v[it] = __ldg(reinterpret_cast<const float4*>(inb + ((size_t)gy * a.X + gx) * C) + c4);
        }
        constexpr int TOTALW = 25 * C * COUT / 4, ITERW = (TOTALW + 127) / 128;
        float4 wv[ITERW];
        const float4* wg = reinterpret_cast<const float4*>(a.w);
#pragma unroll
        for (int it = 0; it < ITERW; ++it) { const int idx = tid + it * 128; wv[it] = (idx < TOTALW) ? __ldg(wg + idx) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int idx = tid + it * 128;
            *reinterpret_cast<float4*>(tin + (idx >> 3) * PS + (idx & 7) * 4) = v[it];
        }
#pragma unroll
        for (int it = 0; it < ITERW; ++it) { const int idx = tid + it * 128; if (idx < TOTALW) smem4[idx] = wv[it]; }
    }
    __syncthreads();
    const int half = tid >> 6, t = tid & 63;
    const int tx = t & 15, ty0 = (t >> 4) * 2;
    float acc[2][COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) { acc[0][c] = 0.0f; acc[1][c] = 0.0f; }
#pragma unroll 1
    for (int dx = 0; dx < 5; ++dx) {
#pragma unroll 1
        for (int cq = 0; cq < 4; ++cq) {
            const int c4 = half * 4 + cq;
            float4 iv[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) iv[r] = *reinterpret_cast<const float4*>(tin + ((ty0 + r) * PW + tx + dx) * PS + c4 * 4);
#pragma unroll
            for (int dy = 0; dy < 5; ++dy) {
                float wv[4 * COUT];                                  // [ci 0..3][co]
                const float4* wq = reinterpret_cast<const float4*>(ws + ((dy * 5 + dx) * C + c4 * 4) * COUT);
#pragma unroll
                for (int q = 0; q < COUT; ++q) {
                    const float4 w = wq[q];
                    wv[4 * q] = w.x; wv[4 * q + 1] = w.y; wv[4 * q + 2] = w.z; wv[4 * q + 3] = w.w;
                }
#pragma unroll
                for (int pz = 0; pz < 2; ++pz) {
                    const float4 v = iv[dy + pz];
#pragma unroll
                    for (int co = 0; co < COUT; ++co)
                        acc[pz][co] = fmaf(v.x, wv[co], fmaf(v.y, wv[COUT + co], fmaf(v.z, wv[2 * COUT + co], fmaf(v.w, wv[3 * COUT + co], acc[pz][co]))));
                }
            }
        }
    }
    __syncthreads();                       // everyone is done with the input tile: reuse it for the half sums
    float* red = tin;                      // [64][2*COUT]
    if (half == 1) {
#pragma unroll
        for (int pz = 0; pz < 2; ++pz)
#pragma unroll
            for (int co = 0; co < COUT; ++co) red[(pz * COUT + co) * 64 + t] = acc[pz][co];
    }
    __syncthreads();
    if (half == 1) return;
#pragma unroll
    for (int pz = 0; pz < 2; ++pz) {
        const int gy = y0 + ty0 + pz, gx = x0 + tx;
        if (gy >= a.Y || gx >= a.X) continue;
        const size_t o = (((size_t)b * a.Y + gy) * a.X + gx) * COUT;
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            float v = acc[pz][co] + red[(pz * COUT + co) * 64 + t];
            if (a.bias) v += __ldg(a.bias + co);
            if (a.addend) v += __ldg(a.addend + o + co);
            const float rf = (a.act == SOL_ACT_DLRELU) ? __ldg(a.ref + o + co) : 0.0f;
            a.out[o + co] = apply_act(v, a.act, a.slope, rf);
        }
    }
}

template <int CIN>
static int launch_expand(const ConvArgs& a, cudaStream_t st) {
    SOL_CUDA(launch_kernel(k_conv5x5_expand<CIN>, dim3(cdiv(a.X, 32), cdiv(a.Y, 4), a.B), dim3(128), 0, st, a));
    SOL_LAUNCHED();
    return SOL_OK;
}

template <int COUT>
static int launch_reduce(const ConvArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)(25 * 32 * COUT + 12 * 20 * 36) * sizeof(float);
    auto kern = k_conv5x5_reduce<COUT>;
    static bool attr_done = false;   // one process per GPU: set once, outside any later graph capture
    if (smem > 48 * 1024 && !attr_done) {
        SOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    SOL_CUDA(launch_kernel(kern, dim3(cdiv(a.X, 16), cdiv(a.Y, 8), a.B), dim3(128), smem, st, a));
    SOL_LAUNCHED();
    return SOL_OK;
}

// ------------------------------------------------------------------------------------------------
// any other (Cin, Cout) — the layers of model_mercury (karman_train.py:92-99: 32 -> 64 -> 2) and their data
// gradients.  Tile 16 (x) x 8 (y) pixels, one pixel per thread, 8 couts per pass in registers; the input tile +
// halo sits in shared memory as [PH][PW][Cin], weights are warp-uniform 128-bit loads.  Plain fp32 FMAs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_conv5x5_generic(const ConvArgs a, int Cin, int Cout) {
    pdl_sync();
    constexpr int TW = 16, TH = 8, PW = TW + 4, PH = TH + 4;
    extern __shared__ float gen_tile[];      // [PH*PW][Cin + 1]  (+1: conflict-free across the 16 pixels of a row)
    const int CP = Cin + 1;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, b = blockIdx.z;
    const float* inb = a.in + (size_t)b * a.Y * a.X * Cin;
    for (int idx = tid; idx < PH * PW * Cin; idx += 128) {
        const int c = idx % Cin, pix = idx / Cin;
        const int tyy = pix / PW, txx = pix - tyy * PW;
        const int gy = y0 + tyy - 2, gx = x0 + txx - 2;
        float v = 0.0f;
        if (gy >= 0 && gy < a.Y && gx >= 0 && gx < a.X) v = __ldg(inb + ((size_t)gy * a.X + gx) * Cin + c);
        gen_tile[pix * CP + c] = v;
    }
    __syncthreads();
    const int px = tid & 15, py = tid >> 4;
    const int gy = y0 + py, gx = x0 + px;
    const bool inside = gy < a.Y && gx < a.X;
    const size_t opix = ((size_t)b * a.Y + gy) * a.X + gx;
    for (int co0 = 0; co0 < Cout; co0 += 8) {
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
        const int nco = Cout - co0 < 8 ? Cout - co0 : 8;
        for (int tap = 0; tap < 25; ++tap) {
            const int dy = tap / 5, dx = tap - dy * 5;
            const float* tp = gen_tile + ((py + dy) * PW + px + dx) * CP;
            const float* wp = a.w + (size_t)tap * Cin * Cout + co0;
            if (nco == 8 && (Cout & 3) == 0) {
                for (int ci = 0; ci < Cin; ++ci) {
                    const float v = tp[ci];
                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)ci * Cout));
                    const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)ci * Cout) + 1);
                    acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
                    acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
                }
            } else {
                for (int ci = 0; ci < Cin; ++ci) {
                    const float v = tp[ci];
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k < nco) acc[k] = fmaf(v, __ldg(wp + (size_t)ci * Cout + k), acc[k]);
                }
            }
        }
        if (inside) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (k >= nco) break;
                const size_t o = opix * Cout + co0 + k;
                float v = acc[k];
                if (a.bias) v += __ldg(a.bias + co0 + k);
                if (a.addend) v += __ldg(a.addend + o);
                const float rf = (a.act == SOL_ACT_DLRELU) ? __ldg(a.ref + o) : 0.0f;
                a.out[o] = apply_act(v, a.act, a.slope, rf);
            }
        }
    }
}

static int launch_conv_generic(const ConvArgs& a, cudaStream_t st, int Cin, int Cout) {
    const size_t smem = (size_t)12 * 20 * (Cin + 1) * sizeof(float);
    if (smem > 200 * 1024) return fail(SOL_ERR_UNSUPPORTED, "conv5x5: too many input channels for the generic kernel");
    static size_t attr_smem = 48 * 1024;
    if (smem > attr_smem) {
        SOL_CUDA(cudaFuncSetAttribute(k_conv5x5_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    SOL_CUDA(launch_kernel(k_conv5x5_generic, dim3(cdiv(a.X, 16), cdiv(a.Y, 8), a.B), dim3(128), smem, st, a, Cin, Cout));
    SOL_LAUNCHED();
    return SOL_OK;
}

// weight gradient for any (Cin, Cout): one CTA per (tap, cin); thread = (cout, pixel group), the pixel groups stride
// over the image rows and are summed through shared memory in a fixed order (deterministic, no atomics); the CTA of
// the centre tap and cin 0 also sums the bias gradient.
__global__ void __launch_bounds__(256) k_wgrad_generic(const float* __restrict__ in, const float* __restrict__ g, float* __restrict__ dW,
                                                       float* __restrict__ db, int B, int Y, int X, int Cin, int Cout, int CP, int accumulate) {
    pdl_sync();
    __shared__ float red[2][256];
    const int tap = blockIdx.x / Cin, ci = blockIdx.x - tap * Cin;
    const int dy = tap / 5 - 2, dx = tap % 5 - 2;
    const bool do_bias = (tap == 12 && ci == 0);
    const int co = threadIdx.x % CP, pg = threadIdx.x / CP, P = 256 / CP;
    float acc = 0.0f, bacc = 0.0f;
    if (co < Cout) {
        for (int row = pg; row < B * Y; row += P) {
            const int b = row / Y, y = row - b * Y;
            const int yy = y + dy;
            const bool rowok = yy >= 0 && yy < Y;
            const float* gr = g + (size_t)row * X * Cout + co;
            const float* ir = in + (((size_t)b * Y + yy) * X + dx) * Cin + ci;
            for (int x = 0; x < X; ++x) {
                const float gv = __ldg(gr + (size_t)x * Cout);
                const int xx = x + dx;
                if (rowok && xx >= 0 && xx < X) acc = fmaf(__ldg(ir + (size_t)x * Cin), gv, acc);
                bacc += gv;
            }
        }
    }
    red[0][threadIdx.x] = acc; red[1][threadIdx.x] = bacc;
    __syncthreads();
    if (pg == 0 && co < Cout) {
        for (int q = 1; q < P; ++q) { acc += red[0][q * CP + co]; bacc += red[1][q * CP + co]; }
        const size_t o = ((size_t)tap * Cin + ci) * Cout + co;
        dW[o] = accumulate ? dW[o] + acc : acc;
        if (do_bias) db[co] = accumulate ? db[co] + bacc : bacc;
    }
}

int launch_conv5x5(cudaStream_t st, int B, int Y, int X, int Cin, int Cout, const float* in, const float* w, const float* bias,
                   const float* addend, const float* ref, int act, float slope, float* out, unsigned int* amax_out, bool weights_ready) {
    ConvArgs a;
    a.in = in; a.w = w; a.bias = bias; a.addend = addend; a.ref = ref; a.out = out;
    a.B = B; a.Y = Y; a.X = X; a.act = act; a.slope = slope; a.amax_out = amax_out;
    if (amax_out && !(Cout == 32 && Cin >= 2 && Cin <= 4)) return fail(SOL_ERR_UNSUPPORTED, "conv5x5: only the Cin <= 4 -> 32 kernels track max|out|");
    if (act == SOL_ACT_DLRELU && !ref) return fail(SOL_ERR_INVALID, "conv5x5: SOL_ACT_DLRELU needs ref");
    if (Cin == 32 && Cout == 32) {
        // pick the tile height that gives at least ~one CTA per SM
        const long ctas8 = (long)cdiv(X, 32) * cdiv(Y, 8) * B;
        if (ctas8 >= 148) {
            const size_t smem = (size_t)12 * 36 * 8 * sizeof(float4);
            static bool attr_done = false;
            if (!attr_done) {
                SOL_CUDA(cudaFuncSetAttribute(k_conv5x5_c32<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                attr_done = true;
            }
            SOL_CUDA(launch_kernel(k_conv5x5_c32<8>, dim3(cdiv(X, 32), cdiv(Y, 8), B), dim3(256), smem, st, a));
        } else {
            const size_t smem = (size_t)8 * 36 * 8 * sizeof(float4);
            SOL_CUDA(launch_kernel(k_conv5x5_c32<4>, dim3(cdiv(X, 32), cdiv(Y, 4), B), dim3(128), smem, st, a));
        }
        SOL_LAUNCHED();
        return SOL_OK;
    }
    if (((uintptr_t)in & 15) && Cin == 32) return fail(SOL_ERR_INVALID, "conv5x5: 32-channel input must be 16-byte aligned");
    if (((uintptr_t)w & 15) || ((uintptr_t)out & 15 && Cout == 32)) return fail(SOL_ERR_INVALID, "conv5x5: weights / 32-channel output must be 16-byte aligned");
    if (g_thin_path == 0 && ((Cout == 32 && Cin >= 2 && Cin <= 4) || (Cin == 32 && Cout >= 2 && Cout <= 4))) {
        if (((uintptr_t)in & 15) && Cin == 32) return fail(SOL_ERR_INVALID, "conv5x5: 32-channel input must be 16-byte aligned");
        if (Cout == 32 && ((addend && ((uintptr_t)addend & 15)) || (ref && ((uintptr_t)ref & 15)))) return fail(SOL_ERR_INVALID, "conv5x5: addend / ref must be 16-byte aligned");
        return launch_conv5x5_thin(st, B, Y, X, Cin, Cout, in, w, bias, addend, ref, act, slope, out, amax_out, weights_ready);
    }
    if (Cout == 32 && Cin == 2) return launch_expand<2>(a, st);
    if (Cout == 32 && Cin == 3) return launch_expand<3>(a, st);
    if (Cout == 32 && Cin == 4) return launch_expand<4>(a, st);
    if (Cin == 32 && Cout == 2) return launch_reduce<2>(a, st);
    if (Cin == 32 && Cout == 3) return launch_reduce<3>(a, st);
    if (Cin == 32 && Cout == 4) return launch_reduce<4>(a, st);
    return launch_conv_generic(a, st, Cin, Cout);
}

// ------------------------------------------------------------------------------------------------
// wT[tap][co][ci] = w[24 - tap][ci][co]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_flip_weights(int Cin, int Cout, const float* __restrict__ w, float* __restrict__ wT) {
    pdl_sync();
    const int n = 25 * Cin * Cout;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int ci = idx % Cin;
        const int co = (idx / Cin) % Cout;
        const int tap = idx / (Cin * Cout);
        wT[idx] = w[((24 - tap) * Cin + ci) * Cout + co];
    }
}

int launch_flip_weights(cudaStream_t st, int Cin, int Cout, const float* w, float* wT) {
    const int n = 25 * Cin * Cout;
    SOL_CUDA(launch_kernel(k_flip_weights, dim3(cdiv(n, 256)), dim3(256), 0, st, Cin, Cout, w, wT));
    SOL_LAUNCHED();
    return SOL_OK;
}

// ------------------------------------------------------------------------------------------------
// weight gradient, 32 -> 32
// ------------------------------------------------------------------------------------------------
constexpr int WG_E = 25 * 32 * 32 + 32;   // entries per partial slot (weights + bias)
constexpr int WG_MAX_CTAS = 148;

struct WgradArgs {
    const float* in;
    const float* g;
    float* part;
    int B, Y, X;
    int accumulate;
};

__global__ void __launch_bounds__(320, 1) k_wgrad_c32(const WgradArgs a) {
    pdl_sync();
    extern __shared__ float4 wsm[];
    const int X = a.X, Y = a.Y;
    const int PWX = X + 4;
    float4* tin = wsm;                 // [5][PWX][8]
    float4* tg = wsm + 5 * PWX * 8;    // [X][8]
    const int tid = threadIdx.x;
    const int co4 = tid & 7, ci4 = (tid >> 3) & 7, dy = tid >> 6;
    float acc[5][4][4];
    float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < 5; ++d)
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc[d][q][0] = acc[d][q][1] = acc[d][q][2] = acc[d][q][3] = 0.0f; }
    const int rows_total = a.B * Y;
    const int r_begin = (int)((long)blockIdx.x * rows_total / gridDim.x);
    const int r_end = (int)((long)(blockIdx.x + 1) * rows_total / gridDim.x);
    const bool do_bias = (ci4 == 0 && dy == 0);
#pragma unroll 1
    for (int r = r_begin; r < r_end; ++r) {
        const int b = r / Y, y = r - b * Y;
        __syncthreads();
        for (int idx = tid; idx < 5 * PWX * 8; idx += 320) {
            const int c4 = idx & 7, pix = idx >> 3;
            const int rr = pix / PWX, xx = pix - rr * PWX;
            const int gy = y + rr - 2, gx = xx - 2;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy >= 0 && gy < Y && gx >= 0 && gx < X)
                v = __ldg(reinterpret_cast<const float4*>(a.in + (((size_t)b * Y + gy) * X + gx) * 32) + c4);
            tin[idx] = v;
        }
        for (int idx = tid; idx < X * 8; idx += 320)
            tg[idx] = __ldg(reinterpret_cast<const float4*>(a.g + ((size_t)b * Y + y) * X * 32) + idx);
        __syncthreads();
        const float4* inrow = tin + dy * PWX * 8;
#pragma unroll 1
        for (int x0 = 0; x0 < X; x0 += 4) {
            float4 wv[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) wv[t] = inrow[(x0 + t) * 8 + ci4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 g4 = tg[(x0 + u) * 8 + co4];
                if (do_bias) { bacc[0] += g4.x; bacc[1] += g4.y; bacc[2] += g4.z; bacc[3] += g4.w; }
#pragma unroll
                for (int d = 0; d < 5; ++d) {
                    const float4 v = wv[u + d];
                    acc[d][0][0] = fmaf(v.x, g4.x, acc[d][0][0]); acc[d][0][1] = fmaf(v.x, g4.y, acc[d][0][1]);
                    acc[d][0][2] = fmaf(v.x, g4.z, acc[d][0][2]); acc[d][0][3] = fmaf(v.x, g4.w, acc[d][0][3]);
                    acc[d][1][0] = fmaf(v.y, g4.x, acc[d][1][0]); acc[d][1][1] = fmaf(v.y, g4.y, acc[d][1][1]);
                    acc[d][1][2] = fmaf(v.y, g4.z, acc[d][1][2]); acc[d][1][3] = fmaf(v.y, g4.w, acc[d][1][3]);
                    acc[d][2][0] = fmaf(v.z, g4.x, acc[d][2][0]); acc[d][2][1] = fmaf(v.z, g4.y, acc[d][2][1]);
                    acc[d][2][2] = fmaf(v.z, g4.z, acc[d][2][2]); acc[d][2][3] = fmaf(v.z, g4.w, acc[d][2][3]);
                    acc[d][3][0] = fmaf(v.w, g4.x, acc[d][3][0]); acc[d][3][1] = fmaf(v.w, g4.y, acc[d][3][1]);
                    acc[d][3][2] = fmaf(v.w, g4.z, acc[d][3][2]); acc[d][3][3] = fmaf(v.w, g4.w, acc[d][3][3]);
                }
            }
        }
    }
    float4* part4 = reinterpret_cast<float4*>(a.part + (size_t)blockIdx.x * WG_E);
#pragma unroll
    for (int d = 0; d < 5; ++d)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int o4 = (((dy * 5 + d) * 32 + ci4 * 4 + q) * 32 + co4 * 4) >> 2;
            float4 v = make_float4(acc[d][q][0], acc[d][q][1], acc[d][q][2], acc[d][q][3]);
            if (a.accumulate) { const float4 o = part4[o4]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
            part4[o4] = v;
        }
    if (do_bias) {
        const int o4 = (25 * 32 * 32 + co4 * 4) >> 2;
        float4 v = make_float4(bacc[0], bacc[1], bacc[2], bacc[3]);
        if (a.accumulate) { const float4 o = part4[o4]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
        part4[o4] = v;
    }
}

__global__ void __launch_bounds__(256) k_wgrad_finalize(int nctas, const float* __restrict__ part, float* __restrict__ dW,
                                                        float* __restrict__ db, int accumulate) {
    pdl_sync();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= WG_E) return;
    float s = 0.0f;
    for (int c = 0; c < nctas; ++c) s += part[(size_t)c * WG_E + e];
    float* dst = (e < 25 * 32 * 32) ? (dW + e) : (db + (e - 25 * 32 * 32));
    *dst = accumulate ? (*dst + s) : s;
}

// All ten 32 -> 32 layers in ONE launch (blockIdx.y = layer): the layers' [weights | bias] blocks are WG_E floats each and
// adjacent in the Keras-ordered gradient buffer, so layer l writes out + l * WG_E.  Same fixed summation order per entry
// as k_wgrad_finalize: CTA slots in four interleaved chains (c = 0, 4, 8, ..; 1, 5, ..; ...) combined at the end.
struct FinalizeMulti { int nctas[10]; };
__global__ void __launch_bounds__(256) k_wgrad_finalize_multi(const FinalizeMulti f, const float* __restrict__ part, size_t part_stride,
                                                              float* __restrict__ out) {
    pdl_sync();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= WG_E) return;
    const int l = blockIdx.y, n = f.nctas[l];
    const float* p = part + (size_t)l * part_stride + e;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    int c = 0;
    for (; c + 3 < n; c += 4) {
        s0 += p[(size_t)c * WG_E]; s1 += p[(size_t)(c + 1) * WG_E]; s2 += p[(size_t)(c + 2) * WG_E]; s3 += p[(size_t)(c + 3) * WG_E];
    }
    for (; c < n; ++c) s0 += p[(size_t)c * WG_E];
    out[(size_t)l * WG_E + e] = (s0 + s1) + (s2 + s3);
}

// ------------------------------------------------------------------------------------------------
// weight gradient of the thin layers, persistent over the tiles of `steps` x B images.  Image (step, b)
// lives at in + step*in_step_stride + b*Y*X*CIN (and likewise for g), so the same kernels serve the
// per-step call (steps = 1) and the deferred call over the whole unrolled sweep.
// Register-tiled outer products: a thread owns 40 accumulators (5 dx taps x 8 values) and walks along x
// with a 5-deep sliding window, so one new input value and one g vector feed 40 FMAs.  240 threads =
// NOWN owners x P row groups; the row groups are summed through shared memory, then one atomicAdd per
// weight and CTA.
// ------------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(256, 2) k_wgrad_expand(const float* __restrict__ in, const float* __restrict__ g, float* dW, float* db,
                                                         int steps, int B, int Y, int X, size_t in_step_stride, size_t g_step_stride,
                                                         float* part) {
    pdl_sync();
    constexpr int COUT = 32;
    constexpr int NOWN = 5 * CIN * 4;              // owner = (dy, ci, cout octet)
    constexpr int P = 240 / NOWN;                  // row groups: 6 / 4 / 3 for CIN = 2 / 3 / 4
    constexpr int TR = 2 * P, TW = 32, PR = TR + 4, PW = TW + 4;
    static_assert(NOWN * P == 240, "thread geometry");
    extern __shared__ float4 wg_sm4[];
    float* tg = reinterpret_cast<float*>(wg_sm4);  // [TR*TW][32]
    float* tin = tg + TR * TW * COUT;              // [CIN][PR][PW] planar
    const int tid = threadIdx.x;
    const bool worker = tid < 240;
    const int pg = tid / NOWN, own = tid - pg * NOWN;
    const int cq = own & 3, ci = (own >> 2) % CIN, dy = own / (4 * CIN);
    const int tiles_x = (X + TW - 1) / TW, tiles_y = (Y + TR - 1) / TR;
    const int tiles_img = tiles_x * tiles_y;
    const int ntiles = tiles_img * steps * B;
    float acc[5][8];
#pragma unroll
    for (int d = 0; d < 5; ++d)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[d][j] = 0.0f;
    float dbacc0 = 0.0f, dbacc1 = 0.0f;            // threads 240..255: bias gradient of couts 2(tid-240), +1
#pragma unroll 1
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int img = tile / tiles_img, rem = tile - img * tiles_img;
        const int tyi = rem / tiles_x, txi = rem - tyi * tiles_x;
        const int step = img / B, b = img - step * B;
        const int x0 = txi * TW, y0 = tyi * TR;
        const float* inb = in + (size_t)step * in_step_stride + (size_t)b * Y * X * CIN;
        const float* gb = g + (size_t)step * g_step_stride + (size_t)b * Y * X * COUT;
        __syncthreads();     // previous tile fully consumed
        {
            constexpr int TOTG = TR * TW * 8, ITG = TOTG / 256;       // float4 loads of g: TR per thread
            static_assert(TOTG % 256 == 0, "g staging geometry");
            float4 v[ITG];
#pragma unroll
            for (int k = 0; k < ITG; ++k) {
                const int idx = tid + k * 256;
                const int c4 = idx & 7, pix = idx >> 3;
                const int gy = y0 + pix / TW, gx = x0 + (pix & (TW - 1));
                v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gy < Y && gx < X) v[k] = __ldg(reinterpret_cast<const float4*>(gb + ((size_t)gy * X + gx) * COUT) + c4);
            }
            constexpr int TOTI = PR * PW * CIN, ITI = (TOTI + 255) / 256;
            float u[ITI];
#pragma unroll
            for (int k = 0; k < ITI; ++k) {
                const int idx = tid + k * 256;
                const int row = idx / (PW * CIN), r2 = idx - row * (PW * CIN);
                const int px = r2 / CIN;
                const int gy = y0 + row - 2, gx = x0 + px - 2;
                u[k] = 0.0f;
                if (idx < TOTI && gy >= 0 && gy < Y && gx >= 0 && gx < X) u[k] = __ldg(inb + ((size_t)gy * X + gx) * CIN + (r2 - px * CIN));
            }
#pragma unroll
            for (int k = 0; k < ITG; ++k) wg_sm4[tid + k * 256] = v[k];
#pragma unroll
            for (int k = 0; k < ITI; ++k) {
                const int idx = tid + k * 256;
                if (idx < TOTI) {
                    const int row = idx / (PW * CIN), r2 = idx - row * (PW * CIN);
                    const int px = r2 / CIN;
                    tin[((r2 - px * CIN) * PR + row) * PW + px] = u[k];
                }
            }
        }
        __syncthreads();
        if (worker) {
#pragma unroll 1
            for (int rr = 0; rr < 2; ++rr) {
                const int y = pg * 2 + rr;
                const float* irow = tin + (ci * PR + y + dy) * PW;
                const float4* grow = reinterpret_cast<const float4*>(tg + (size_t)y * TW * COUT + cq * 8);
                float w0 = irow[0], w1 = irow[1], w2 = irow[2], w3 = irow[3];
#pragma unroll 4
                for (int x = 0; x < TW; ++x) {
                    const float w4 = irow[x + 4];
                    const float4 ga = grow[x * 8], gc = grow[x * 8 + 1];
                    const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gc.x, gc.y, gc.z, gc.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        acc[0][j] = fmaf(w0, gv[j], acc[0][j]);
                        acc[1][j] = fmaf(w1, gv[j], acc[1][j]);
                        acc[2][j] = fmaf(w2, gv[j], acc[2][j]);
                        acc[3][j] = fmaf(w3, gv[j], acc[3][j]);
                        acc[4][j] = fmaf(w4, gv[j], acc[4][j]);
                    }
                    w0 = w1; w1 = w2; w2 = w3; w3 = w4;
                }
            }
        } else {
            const int co = 2 * (tid - 240);
#pragma unroll 4
            for (int pix = 0; pix < TR * TW; ++pix) {
                const float2 t2 = *reinterpret_cast<const float2*>(tg + pix * COUT + co);
                dbacc0 += t2.x; dbacc1 += t2.y;
            }
        }
    }
    // sum the row groups: groups 1..P-1 park their accumulators in shared memory, group 0 adds and publishes
    __syncthreads();
    float* red = reinterpret_cast<float*>(wg_sm4);           // [(P-1)][40][NOWN]
    static_assert((P - 1) * 40 * NOWN <= TR * TW * COUT + CIN * PR * PW, "reduction buffer must fit the tiles");
    if (worker && pg > 0) {
#pragma unroll
        for (int d = 0; d < 5; ++d)
#pragma unroll
            for (int j = 0; j < 8; ++j) red[((pg - 1) * 40 + d * 8 + j) * NOWN + own] = acc[d][j];
    }
    __syncthreads();
    if (worker && pg == 0) {
#pragma unroll
        for (int d = 0; d < 5; ++d)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float s = acc[d][j];
#pragma unroll
                for (int q = 0; q < P - 1; ++q) s += red[(q * 40 + d * 8 + j) * NOWN + own];
                const int e = ((dy * 5 + d) * CIN + ci) * COUT + cq * 8 + j;
                // deterministic mode: a private slot per CTA, summed in CTA order by k_thin_finalize; else one atomic per value
                if (part) part[(size_t)blockIdx.x * THIN_SLOT + e] = s; else atomicAdd(dW + e, s);
            }
    } else if (!worker) {
        if (part) {
            part[(size_t)blockIdx.x * THIN_SLOT + 25 * CIN * COUT + 2 * (tid - 240)] = dbacc0;
            part[(size_t)blockIdx.x * THIN_SLOT + 25 * CIN * COUT + 2 * (tid - 240) + 1] = dbacc1;
        } else {
            atomicAdd(db + 2 * (tid - 240), dbacc0);
            atomicAdd(db + 2 * (tid - 240) + 1, dbacc1);
        }
    }
}

// 32 -> COUT <= 2: owner = (dy, cin quad), 40 accumulators = 5 dx x 4 cin x COUT
template <int COUT>
__global__ void __launch_bounds__(256, 2) k_wgrad_reduce(const float* __restrict__ in, const float* __restrict__ g, float* dW, float* db,
                                                         int steps, int B, int Y, int X, size_t in_step_stride, size_t g_step_stride,
                                                         float* part) {
    pdl_sync();
    static_assert(COUT == 2, "accumulator tile is written for two output channels");
    constexpr int CIN = 32, NOWN = 40, P = 6, TR = 2 * P, TW = 32, PR = TR + 4, PW = TW + 4, PS = 36;   // PS: padded pixel stride
    extern __shared__ float4 wg_sm4[];
    float* tin = reinterpret_cast<float*>(wg_sm4);   // [PR*PW][PS]
    float* tg = tin + PR * PW * PS;                  // [TR*TW][COUT]
    const int tid = threadIdx.x;
    const bool worker = tid < 240;
    const int pg = tid / NOWN, own = tid - pg * NOWN;
    const int c4 = own & 7, dy = own >> 3;
    const int tiles_x = (X + TW - 1) / TW, tiles_y = (Y + TR - 1) / TR;
    const int tiles_img = tiles_x * tiles_y;
    const int ntiles = tiles_img * steps * B;
    float acc[5][4][COUT];
#pragma unroll
    for (int d = 0; d < 5; ++d)
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc[d][k][0] = 0.0f; acc[d][k][1] = 0.0f; }
    float dbacc = 0.0f;                              // threads 240, 241: bias gradient
#pragma unroll 1
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int img = tile / tiles_img, rem = tile - img * tiles_img;
        const int tyi = rem / tiles_x, txi = rem - tyi * tiles_x;
        const int step = img / B, b = img - step * B;
        const int x0 = txi * TW, y0 = tyi * TR;
        const float* inb = in + (size_t)step * in_step_stride + (size_t)b * Y * X * CIN;
        const float* gb = g + (size_t)step * g_step_stride + (size_t)b * Y * X * COUT;
        __syncthreads();
        {
            constexpr int TOTI = PR * PW * 8, CH = 6;                  // 4608 float4 = 18 per thread, 6 in flight
            static_assert(TOTI % (256 * CH) == 0, "input staging geometry");
#pragma unroll 1
            for (int it0 = 0; it0 < TOTI / 256; it0 += CH) {
                float4 v[CH];
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    const int idx = tid + (it0 + k) * 256;
                    const int q4 = idx & 7, pix = idx >> 3;
                    const int row = pix / PW, px = pix - row * PW;
                    const int gy = y0 + row - 2, gx = x0 + px - 2;
                    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (gy >= 0 && gy < Y && gx >= 0 && gx < X) v[k] = __ldg(reinterpret_cast<const float4*>(inb + ((size_t)gy * X + gx) * CIN) + q4);
                }
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    const int idx = tid + (it0 + k) * 256;
                    *reinterpret_cast<float4*>(tin + (idx >> 3) * PS + (idx & 7) * 4) = v[k];
                }
            }
            constexpr int TOTG = TR * TW * COUT, ITG = TOTG / 256;     // 768 floats = 3 per thread
            static_assert(TOTG % 256 == 0, "g staging geometry");
#pragma unroll
            for (int k = 0; k < ITG; ++k) {
                const int idx = tid + k * 256;
                const int pix = idx / COUT;
                const int gy = y0 + pix / TW, gx = x0 + (pix & (TW - 1));
                tg[idx] = (gy < Y && gx < X) ? __ldg(gb + ((size_t)gy * X + gx) * COUT + (idx - pix * COUT)) : 0.0f;
            }
        }
        __syncthreads();
        if (worker) {
#pragma unroll 1
            for (int rr = 0; rr < 2; ++rr) {
                const int y = pg * 2 + rr;
                const float* irow = tin + (size_t)((y + dy) * PW) * PS + c4 * 4;
                const float2* grow = reinterpret_cast<const float2*>(tg + y * TW * COUT);
                float4 w0 = *reinterpret_cast<const float4*>(irow), w1 = *reinterpret_cast<const float4*>(irow + PS),
                       w2 = *reinterpret_cast<const float4*>(irow + 2 * PS), w3 = *reinterpret_cast<const float4*>(irow + 3 * PS);
#pragma unroll 4
                for (int x = 0; x < TW; ++x) {
                    const float4 w4 = *reinterpret_cast<const float4*>(irow + (x + 4) * PS);
                    const float2 gv = grow[x];
                    const float4 ww[5] = {w0, w1, w2, w3, w4};
#pragma unroll
                    for (int d = 0; d < 5; ++d) {
                        acc[d][0][0] = fmaf(ww[d].x, gv.x, acc[d][0][0]); acc[d][0][1] = fmaf(ww[d].x, gv.y, acc[d][0][1]);
                        acc[d][1][0] = fmaf(ww[d].y, gv.x, acc[d][1][0]); acc[d][1][1] = fmaf(ww[d].y, gv.y, acc[d][1][1]);
                        acc[d][2][0] = fmaf(ww[d].z, gv.x, acc[d][2][0]); acc[d][2][1] = fmaf(ww[d].z, gv.y, acc[d][2][1]);
                        acc[d][3][0] = fmaf(ww[d].w, gv.x, acc[d][3][0]); acc[d][3][1] = fmaf(ww[d].w, gv.y, acc[d][3][1]);
                    }
                    w0 = w1; w1 = w2; w2 = w3; w3 = w4;
                }
            }
        } else if (tid < 240 + COUT) {
#pragma unroll 4
            for (int pix = 0; pix < TR * TW; ++pix) dbacc += tg[pix * COUT + (tid - 240)];
        }
    }
    __syncthreads();
    float* red = tin;                                  // [(P-1)][40][NOWN]
    static_assert((P - 1) * 40 * NOWN <= PR * PW * PS, "reduction buffer must fit the input tile");
    if (worker && pg > 0) {
#pragma unroll
        for (int d = 0; d < 5; ++d)
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int co = 0; co < COUT; ++co) red[((pg - 1) * 40 + (d * 4 + k) * COUT + co) * NOWN + own] = acc[d][k][co];
    }
    __syncthreads();
    if (worker && pg == 0) {
#pragma unroll
        for (int d = 0; d < 5; ++d)
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int co = 0; co < COUT; ++co) {
                    float s = acc[d][k][co];
#pragma unroll
                    for (int q = 0; q < P - 1; ++q) s += red[(q * 40 + (d * 4 + k) * COUT + co) * NOWN + own];
                    const int e = ((dy * 5 + d) * CIN + c4 * 4 + k) * COUT + co;
                    if (part) part[(size_t)blockIdx.x * THIN_SLOT + e] = s; else atomicAdd(dW + e, s);
                }
    } else if (!worker && tid < 240 + COUT) {
        if (part) part[(size_t)blockIdx.x * THIN_SLOT + 25 * CIN * COUT + (tid - 240)] = dbacc; else atomicAdd(db + (tid - 240), dbacc);
    }
}

// deterministic mode of the thin-layer weight gradients: dW / db += sum over the CTA slots, in CTA order
__global__ void __launch_bounds__(256) k_thin_finalize(int nctas, const float* __restrict__ part, int nW, int nb, float* __restrict__ dW,
                                                       float* __restrict__ db) {
    pdl_sync();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nW + nb) return;
    float s = 0.0f;
    for (int c = 0; c < nctas; ++c) s += part[(size_t)c * THIN_SLOT + e];
    float* dst = e < nW ? dW + e : db + (e - nW);
    *dst += s;
}

template <int CIN>
static int launch_wgrad_expand(cudaStream_t st, int steps, int B, int Y, int X, const float* in, size_t in_step_stride, const float* g,
                               size_t g_step_stride, float* dW, float* db, int max_ctas, float* part) {
    constexpr int P = 240 / (20 * CIN), TR = 2 * P;
    const size_t smem = (size_t)(TR * 32 * 32 + CIN * (TR + 4) * 36) * sizeof(float);
    auto kern = k_wgrad_expand<CIN>;
    static bool attr_done = false;
    if (smem > 48 * 1024 && !attr_done) {
        SOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    const int ntiles = cdiv(X, 32) * cdiv(Y, TR) * B * steps;
    const int cap = max_ctas > 0 ? max_ctas : 2 * 148;
    const int grid = ntiles < cap ? ntiles : cap;
    SOL_CUDA(launch_kernel(kern, dim3(grid), dim3(256), smem, st, in, g, dW, db, steps, B, Y, X, in_step_stride, g_step_stride, part));
    SOL_LAUNCHED();
    if (part) {
        SOL_CUDA(launch_kernel(k_thin_finalize, dim3(cdiv(25 * CIN * 32 + 32, 256)), dim3(256), 0, st, grid, (const float*)part, 25 * CIN * 32, 32, dW, db));
        SOL_LAUNCHED();
    }
    return SOL_OK;
}

static int launch_wgrad_reduce2(cudaStream_t st, int steps, int B, int Y, int X, const float* in, size_t in_step_stride, const float* g,
                                size_t g_step_stride, float* dW, float* db, int max_ctas, float* part) {
    const size_t smem = (size_t)(16 * 36 * 36 + 12 * 32 * 2) * sizeof(float);
    auto kern = k_wgrad_reduce<2>;
    static bool attr_done = false;
    if (!attr_done) {
        SOL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    const int ntiles = cdiv(X, 32) * cdiv(Y, 12) * B * steps;
    const int cap = max_ctas > 0 ? max_ctas : 2 * 148;
    const int grid = ntiles < cap ? ntiles : cap;
    SOL_CUDA(launch_kernel(kern, dim3(grid), dim3(256), smem, st, in, g, dW, db, steps, B, Y, X, in_step_stride, g_step_stride, part));
    SOL_LAUNCHED();
    if (part) {
        SOL_CUDA(launch_kernel(k_thin_finalize, dim3(cdiv(25 * 32 * 2 + 2, 256)), dim3(256), 0, st, grid, (const float*)part, 25 * 32 * 2, 2, dW, db));
        SOL_LAUNCHED();
    }
    return SOL_OK;
}

size_t wgrad_thin_part_floats() { return (size_t)2 * 148 * THIN_SLOT; }

// thin-layer weight gradient accumulated INTO dW/db over `steps` unrolled steps: one atomic per value and CTA, or — with a `part`
// scratch of wgrad_thin_part_floats() floats (deterministic mode) — private CTA slots summed in CTA order
int launch_wgrad_thin_multi(cudaStream_t st, int steps, int B, int Y, int X, int Cin, int Cout, const float* in, size_t in_step_stride,
                            const float* g, size_t g_step_stride, float* dW, float* db, int max_ctas, float* part) {
    if (part && (max_ctas <= 0 || max_ctas > 2 * 148)) max_ctas = 2 * 148;
    if (((uintptr_t)in & 15) && Cin == 32) return fail(SOL_ERR_INVALID, "wgrad: 32-channel input must be 16-byte aligned");
    if (((uintptr_t)g & 15) && Cout == 32) return fail(SOL_ERR_INVALID, "wgrad: 32-channel gradient must be 16-byte aligned");
    if (Cout == 32 && Cin == 2) return launch_wgrad_expand<2>(st, steps, B, Y, X, in, in_step_stride, g, g_step_stride, dW, db, max_ctas, part);
    if (Cout == 32 && Cin == 3) return launch_wgrad_expand<3>(st, steps, B, Y, X, in, in_step_stride, g, g_step_stride, dW, db, max_ctas, part);
    if (Cout == 32 && Cin == 4) return launch_wgrad_expand<4>(st, steps, B, Y, X, in, in_step_stride, g, g_step_stride, dW, db, max_ctas, part);
    if (Cin == 32 && Cout == 2) return launch_wgrad_reduce2(st, steps, B, Y, X, in, in_step_stride, g, g_step_stride, dW, db, max_ctas, part);
    return fail(SOL_ERR_UNSUPPORTED, "wgrad: unsupported (Cin, Cout) pair");
}

size_t wgrad_workspace_floats(int Cin, int Cout) {
    if (Cin == 32 && Cout == 32) return (size_t)WG_MAX_CTAS * WG_E;
    return 0;
}

int launch_wgrad_finalize_n(cudaStream_t st, int nctas, const float* partials, float* dW, float* db, int accumulate) {
    SOL_CUDA(launch_kernel(k_wgrad_finalize, dim3(cdiv(WG_E, 256)), dim3(256), 0, st, nctas, partials, dW, db, accumulate));
    SOL_LAUNCHED();
    return SOL_OK;
}

int launch_wgrad_finalize_multi(cudaStream_t st, const int* nctas10, const float* partials, size_t part_stride, float* out) {
    FinalizeMulti f;
    for (int l = 0; l < 10; ++l) f.nctas[l] = nctas10[l];
    SOL_CUDA(launch_kernel(k_wgrad_finalize_multi, dim3(cdiv(WG_E, 256), 10), dim3(256), 0, st, f, partials, part_stride, out));
    SOL_LAUNCHED();
    return SOL_OK;
}

static int wgrad_ctas(int B, int Y) {
    const int rows = B * Y;
    return rows < WG_MAX_CTAS ? rows : WG_MAX_CTAS;
}

int launch_wgrad(cudaStream_t st, int B, int Y, int X, int Cin, int Cout, const float* in, const float* g_out, float* dW, float* db,
                 int accumulate, float* partials, bool finalize) {
    // (rows that are not a multiple of 4 pixels: the generic kernel below, accumulating straight into dW / db)
    if (Cin == 32 && Cout == 32 && X % 4 == 0) {
        if (!partials) return fail(SOL_ERR_WORKSPACE, "wgrad 32->32 needs the partials workspace");
        const int nctas = wgrad_ctas(B, Y);
        if (in) {
            WgradArgs a;
            a.in = in; a.g = g_out; a.part = partials; a.B = B; a.Y = Y; a.X = X;
            // stand-alone call (finalize=true): partials are scratch; engine call: partials carry the
            // running sum over the unrolled steps and `accumulate` applies to them
            a.accumulate = finalize ? 0 : accumulate;
            const size_t smem = ((size_t)5 * (X + 4) * 8 + (size_t)X * 8) * sizeof(float4);
            static size_t attr_smem = 0;
            if (smem > attr_smem) {
                SOL_CUDA(cudaFuncSetAttribute(k_wgrad_c32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                attr_smem = smem;
            }
            SOL_CUDA(launch_kernel(k_wgrad_c32, dim3(nctas), dim3(320), smem, st, a));
            SOL_LAUNCHED();
        }
        if (finalize) {
            SOL_CUDA(launch_kernel(k_wgrad_finalize, dim3(cdiv(WG_E, 256)), dim3(256), 0, st, nctas, partials, dW, db, accumulate));
            SOL_LAUNCHED();
        }
        return SOL_OK;
    }
    if (!in) return SOL_OK;   // finalize-only call on a layer without partial sums: nothing to do
    if (!dW || !db) return fail(SOL_ERR_INVALID, "wgrad: dW / db required");
    const bool thin = (Cout == 32 && Cin >= 2 && Cin <= 4) || (Cin == 32 && Cout == 2);
    if (!thin) {
        int CP = 32;
        while (CP < Cout) CP *= 2;
        if (CP > 256) return fail(SOL_ERR_UNSUPPORTED, "wgrad: more than 256 output channels");
        SOL_CUDA(launch_kernel(k_wgrad_generic, dim3(25 * Cin), dim3(256), 0, st, in, g_out, dW, db, B, Y, X, Cin, Cout, CP, accumulate));
        SOL_LAUNCHED();
        return SOL_OK;
    }
    if (!accumulate) {
        SOL_CUDA(cudaMemsetAsync(dW, 0, (size_t)25 * Cin * Cout * sizeof(float), st));
        SOL_CUDA(cudaMemsetAsync(db, 0, (size_t)Cout * sizeof(float), st));
    }
    return launch_wgrad_thin_multi(st, 1, B, Y, X, Cin, Cout, in, 0, g_out, 0, dW, db);
}

}  // namespace sol
