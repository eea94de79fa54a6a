// Measures the kernel-to-kernel gap of a dependent chain inside a replayed CUDA graph on B200, for kernels that
// successively add the resource features of the tensor-core convolution (large dynamic shared memory, TMEM
// allocation, bulk output stores), with plain stream order and with programmatic dependent launch.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gap_probe gap_probe.cu
#include <cuda_runtime.h>
#include <cuda.h>
#include "../../solver_in_the_loop_b200/csrc/sol_tc_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

struct Args {
    long long* stamps;   // [launch][cta][4]: start, released, end, after TMEM dealloc (globaltimer ns)
    float* out;          // 3 MB output
    int launch, features, spin_clks, pdl;
};

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

using namespace sol::tc;

__global__ void __launch_bounds__(192, 2) k_probe(const __grid_constant__ CUtensorMap map, const Args a) {
    extern __shared__ unsigned char smem[];
    __shared__ unsigned tmem_slot;
    __shared__ __align__(8) unsigned long long bars[2];
    const int cta = blockIdx.x, nct = gridDim.x;
    long long* st = a.stamps + ((size_t)a.launch * nct + cta) * 8;
    if (threadIdx.x == 0) st[0] = (long long)gtime();
    unsigned tmem = 0;
    if (a.features & 2) {
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&tmem_slot)), "r"(256u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem = tmem_slot;
    }
    if ((a.features & 24) && threadIdx.x == 0) {
        mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) st[1] = (long long)gtime();
    if (a.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (a.features & 1) smem[threadIdx.x * 64] = (unsigned char)cta;     // touch the dynamic shared memory
    const unsigned sbase = (smem_u32(smem) + 1023u) & ~1023u;
    if (a.features & 8) {        // 8 x 8 KB TMA loads (like the weight stream) + wait
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(smem_u32(&bars[0]), 8 * 8192);
            for (int i = 0; i < 8; ++i) tma_load_2d(sbase + i * 8192, &map, smem_u32(&bars[0]), 0, ((cta + i) % 25) * 64);
        }
        mbar_wait(smem_u32(&bars[0]), 0);
    }
    if (a.features & 16) {        // operands = 1.0 so that a complete accumulator holds 8 x (number of MMAs into it)
        float* fa = reinterpret_cast<float*>(smem + (sbase - smem_u32(smem)));
        for (int i = threadIdx.x; i < 4096; i += 192) { fa[i] = 1.0f; fa[16384 + i % 1024] = 1.0f; }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
    if ((a.features & 16) && threadIdx.x == 32) {     // 64 tf32 UMMAs M=128 N=32 on whatever the shared memory holds + commit + wait
        const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        const unsigned long long dA = make_desc(sbase, 1024, 0), dB = make_desc(sbase + 65536, 1024, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int nmma = (a.features & 32) ? 800 : 64;
        for (int i = 0; i < nmma; ++i) umma_tf32(tmem + 32u * (i & 1), dA + 2 * (i & 3), dB + 2 * (i & 3), idesc, i > 1 ? 1u : 0u);
        umma_commit(smem_u32(&bars[1]));
        mbar_wait(smem_u32(&bars[1]), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        st[4] = (long long)gtime();                  // commit observed
    }
    if ((a.features & 16) && (threadIdx.x >> 5) == 1) {       // whole warp 1: read accumulator 0 right after the commit was observed
        __syncwarp();
        unsigned v0, v1, v2, v3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(tmem + (32u << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (threadIdx.x == 32) { st[5] = (long long)gtime(); st[6] = (long long)__uint_as_float(v0); }
    }
    const long long t0 = clock64();
    while (clock64() - t0 < a.spin_clks) { }
    if (a.features & 4) {      // 16 KB of output per CTA (3 MB per launch), float4 stores
        float4* o = reinterpret_cast<float4*>(a.out) + (size_t)cta * 1024;
        for (int i = threadIdx.x; i < 1024; i += 192) o[i] = make_float4(1.f, 2.f, 3.f, (float)a.launch);
    }
    __syncthreads();
    if (threadIdx.x == 0) st[2] = (long long)gtime();
    if (a.features & 2) {
        if (threadIdx.x < 32) {
            if (a.features & 64) {      // read the accumulator back first (like an epilogue)
                unsigned v0, v1, v2, v3;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(tmem));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (v0 == 0x12345u && v1 + v2 + v3 == 7u) st[1] = 0;
            }
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
            if (threadIdx.x == 0) st[3] = (long long)gtime();
        }
    }
}

int main() {
    const int NL = 24, NCTA = 192;
    long long* stamps; float* out;
    CK(cudaMalloc(&stamps, sizeof(long long) * NL * NCTA * 8));
    CK(cudaMalloc(&out, 4 << 20));
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    struct Cfg { const char* name; int features; size_t smem; int ctas; };
    const Cfg cfgs[] = {{"bare kernel, 192 CTAs", 0, 0, 192}, {"+109 KB dynamic smem", 1, 109 * 1024, 192}, {"+TMEM alloc/dealloc", 3, 109 * 1024, 192},
                        {"+3 MB output stores", 7, 109 * 1024, 192}, {"same, 148 CTAs (1 per SM)", 7, 109 * 1024, 148},
                        {"stores only, no smem/TMEM", 4, 0, 192}, {"smem+TMEM+stores+TMA loads", 15, 109 * 1024, 192},
                        {"smem+TMEM+stores+TMA+64 UMMAs", 31, 109 * 1024, 192}, {"smem+TMEM+64 UMMAs (no TMA)", 19, 109 * 1024, 192},
                        {"smem+TMEM+800 UMMAs", 19 + 32, 109 * 1024, 192}, {"smem+TMEM+64 UMMAs+tcgen05.ld", 19 + 64, 109 * 1024, 192},
                        {"smem+TMEM+800 UMMAs+tcgen05.ld", 19 + 96, 109 * 1024, 192}, {"800 UMMAs, 148 CTAs", 19 + 32, 109 * 1024, 148}};
    // weight-like tensor: [1600 rows][32 floats]
    float* wbuf; CK(cudaMalloc(&wbuf, 1600 * 128)); CK(cudaMemset(wbuf, 0, 1600 * 128));
    CUtensorMap map;
    {
        sol::tc::EncodeTiledFn enc = nullptr; void* pf = nullptr; cudaDriverEntryPointQueryResult qr;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &pf, cudaEnableDefault, &qr)); enc = (sol::tc::EncodeTiledFn)pf;
        cuuint64_t dims[2] = {32, 1600}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {32, 64}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, wbuf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    }
    for (int pdl = 0; pdl < 2; ++pdl)
        for (const Cfg& c : cfgs) {
            CK(cudaMemset(stamps, 0, sizeof(long long) * NL * NCTA * 8));
            cudaGraph_t g; cudaGraphExec_t ge;
            CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            for (int l = 0; l < NL; ++l) {
                Args a{stamps, out, l, c.features, (c.features & 32) ? 2000 : 16000, pdl};
                cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(c.ctas); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = c.smem; cfg.stream = s;
                cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
                CK(cudaLaunchKernelEx(&cfg, k_probe, map, a));
            }
            CK(cudaStreamEndCapture(s, &g));
            CK(cudaGraphInstantiate(&ge, g, 0));
            for (int r = 0; r < 4; ++r) CK(cudaGraphLaunch(ge, s));
            CK(cudaStreamSynchronize(s));
            std::vector<long long> h((size_t)NL * NCTA * 8);
            CK(cudaMemcpy(h.data(), stamps, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            double gap_start = 0, gap_rel = 0, dur = 0, period = 0, dealloc = 0, t_commit = 0, t_ld = 0, accv = 0; int n = 0;
            for (int l = 1; l < NL; ++l) {
                long long pend = 0, start = 1LL << 62, rel = 1LL << 62, end = 0, s0 = 1LL << 62, pe0 = 0;
                for (int k = 0; k < c.ctas; ++k) {
                    const long long* p = &h[((size_t)(l - 1) * c.ctas + k) * 8];   // NOTE: stamps are indexed with nct = gridDim.x
                    const long long* q = &h[((size_t)l * c.ctas + k) * 8];
                    if (p[3]) dealloc += (p[3] - p[2]) / 1e3 / c.ctas;
                    if (p[4]) { t_commit += (p[4] - p[1]) / 1e3 / c.ctas; t_ld += (p[5] - p[1]) / 1e3 / c.ctas; accv += (double)p[6] / c.ctas; }
                    pend = std::max(pend, p[2]); start = std::min(start, q[0]); rel = std::min(rel, q[1]); end = std::max(end, q[2]);
                }
                (void)s0; (void)pe0;
                gap_start += (start - pend) / 1e3; gap_rel += (rel - pend) / 1e3; dur += (end - rel) / 1e3;
                ++n;
            }
            long long e_first = 0, e_last = 0;
            for (int k = 0; k < c.ctas; ++k) { e_first = std::max(e_first, h[((size_t)1 * c.ctas + k) * 8 + 2]); e_last = std::max(e_last, h[((size_t)(NL - 1) * c.ctas + k) * 8 + 2]); }
            period = (e_last - e_first) / 1e3 / (NL - 2);
            printf("pdl %d  %-32s: last end -> first start %6.2f us, -> first released %6.2f us, released -> end %6.2f us, period %6.2f us, end -> after dealloc %5.2f us | released -> commit seen %5.2f us, -> accumulator read %5.2f us, value %.0f\n", pdl, c.name,
                   gap_start / n, gap_rel / n, dur / n, period, dealloc / n, t_commit / n, t_ld / n, accv / n);
            CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
        }
    return 0;
}
