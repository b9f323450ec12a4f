// microbench.cu -- B200 pipe rates that the UKF kernel design depends on (FP64 FMA, FP64 mma.sync, rcp/rsq seed,
// shuffles, 64-bit shared loads).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cuda_runtime.h>
#include <cstdio>

#define ITERS 2048

__global__ void k_dfma(double* out, double a, double b)
{
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    double s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 1.2345) out[0] = s;
}

template <int CH>
__global__ void k_dmma884(double* out, double a, double b)
{
    double c[CH][2];
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    double s = 0;
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    if (s == 1.2345) out[0] = s;
}

template <int CH>
__global__ void k_dmma16816(double* out, double a, double b)
{
    double c[CH][4];
    for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%4,%4,%4,%4,%4,%4,%4}, {%5,%5,%5,%5}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b));
    double s = 0;
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 1.2345) out[0] = s;
}

template <int NCH>
__global__ void k_dfma_chains(double* out, double a, double b)
{
    double x[NCH];
    for (int i = 0; i < NCH; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NCH; ++i) x[i] = fma(x[i], a, b);
    double s = 0;
    for (int i = 0; i < NCH; ++i) s += x[i];
    if (s == 1.2345) out[0] = s;
}

/* 4 DFMA chains interleaved with 1 DMMA chain per iteration: do the two share one pipe? */
__global__ void k_mix(double* out, double a, double b)
{
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    double c[2][2] = {{1, 2}, {3, 4}};
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
        for (int i = 0; i < 2; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = c[0][0] + c[0][1] + c[1][0] + c[1][1];
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 1.2345) out[0] = s;
}

__global__ void k_rcp(double* out, double a)
{
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i + a;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int hi = __double2hiint(x[i]);
            float r;
            asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(x[i]) : "d"(x[i]));
            (void)hi; (void)r;
        }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 1.2345) out[0] = s;
}

__global__ void k_shfl(double* out)
{
    int x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1 + (i & 3));
    int s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345) out[0] = s;
}

__global__ void k_lds64(double* out)
{
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] += sm[(idx + i * 256 + it) & 4095];
    double t = 0;
    for (int i = 0; i < 8; ++i) t += s[i];
    if (t == 1.2345) out[0] = t;
}

template <class F>
static float timeit(F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    double* out;
    cudaMalloc(&out, 64);
    const int blocks = sms * 4, threads = 256;
    const double warps = double(blocks) * threads / 32;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, clk_khz);
    auto rep = [&](const char* name, float ms, double instr_per_warp, double flops_per_instr) {
        double inst = warps * instr_per_warp;
        double per_clk_sm = inst / (ms * 1e-3) / (clk_khz * 1e3) / sms;
        printf("%-22s %8.3f ms  %7.3f warp-instr/clk/SM  %8.2f TFLOP/s\n", name, ms, per_clk_sm, inst * flops_per_instr / (ms * 1e-3) / 1e12);
    };
    rep("dfma", timeit([&] { k_dfma<<<blocks, threads>>>(out, 0.999, 1e-9); }), ITERS * 8.0, 64);
    rep("dmma m8n8k4 x2", timeit([&] { k_dmma884<2><<<blocks, threads>>>(out, 0.999, 1e-9); }), ITERS * 2.0, 512);
    rep("dmma m8n8k4 x4", timeit([&] { k_dmma884<4><<<blocks, threads>>>(out, 0.999, 1e-9); }), ITERS * 4.0, 512);
    rep("dmma m8n8k4 x8", timeit([&] { k_dmma884<8><<<blocks, threads>>>(out, 0.999, 1e-9); }), ITERS * 8.0, 512);
    rep("dmma m8n8k4 x1 (lat)", timeit([&] { k_dmma884<1><<<sms, 32>>>(out, 0.999, 1e-9); }), 0, 0);
    {
        float ms = timeit([&] { k_dmma884<1><<<sms, 32>>>(out, 0.999, 1e-9); });
        printf("dmma m8n8k4 dependent-chain latency: %.1f cycles\n", ms * 1e-3 * clk_khz * 1e3 / ITERS);
        ms = timeit([&] { k_dfma<<<sms, 32>>>(out, 0.999, 1e-9); });
        printf("dfma 8-chain per-iteration: %.1f cycles (8 independent)\n", ms * 1e-3 * clk_khz * 1e3 / ITERS);
    }
    {
        float ms;
        ms = timeit([&] { k_dfma_chains<1><<<sms, 32>>>(out, 0.999, 1e-9); });
        printf("single warp, 1 dfma chain : %.2f cycles per dfma (latency)\n", ms * 1e-3 * clk_khz * 1e3 / ITERS / 1);
        ms = timeit([&] { k_dfma_chains<2><<<sms, 32>>>(out, 0.999, 1e-9); });
        printf("single warp, 2 dfma chains: %.2f cycles per dfma\n", ms * 1e-3 * clk_khz * 1e3 / ITERS / 2);
        ms = timeit([&] { k_dfma_chains<4><<<sms, 32>>>(out, 0.999, 1e-9); });
        printf("single warp, 4 dfma chains: %.2f cycles per dfma\n", ms * 1e-3 * clk_khz * 1e3 / ITERS / 4);
        ms = timeit([&] { k_dfma_chains<16><<<sms, 32>>>(out, 0.999, 1e-9); });
        printf("single warp, 16 dfma chains: %.2f cycles per dfma\n", ms * 1e-3 * clk_khz * 1e3 / ITERS / 16);
        ms = timeit([&] { k_dfma_chains<4><<<sms, 128>>>(out, 0.999, 1e-9); });
        printf("4 warps (1/SMSP), 4 chains: %.2f cycles per dfma per warp\n", ms * 1e-3 * clk_khz * 1e3 / ITERS / 4);
        ms = timeit([&] { k_dfma_chains<2><<<sms, 512>>>(out, 0.999, 1e-9); });
        printf("16 warps (4/SMSP), 2 chains: %.2f SM-cycles per warp-dfma\n", ms * 1e-3 * clk_khz * 1e3 / ITERS / 2 / 16);
        ms = timeit([&] { k_dfma_chains<1><<<sms, 512>>>(out, 0.999, 1e-9); });
        printf("16 warps (4/SMSP), 1 chain : %.2f SM-cycles per warp-dfma\n", ms * 1e-3 * clk_khz * 1e3 / ITERS / 1 / 16);
    }
    {
        float ms = timeit([&] { k_mix<<<blocks, threads>>>(out, 0.999, 1e-9); });
        double cyc = ms * 1e-3 * clk_khz * 1e3 * sms / (warps * ITERS);
        printf("mix 8 dfma + 2 dmma per iteration: %.2f SM-cycles per iteration (separate pipes: ~8.6; shared: ~12.7)\n", cyc);
    }
    rep("dmma m16n8k16 x2", timeit([&] { k_dmma16816<2><<<blocks, threads>>>(out, 0.999, 1e-9); }), ITERS * 2.0, 4096);
    rep("dmma m16n8k16 x4", timeit([&] { k_dmma16816<4><<<blocks, threads>>>(out, 0.999, 1e-9); }), ITERS * 4.0, 4096);
    rep("rcp.approx.f64", timeit([&] { k_rcp<<<blocks, threads>>>(out, 1.5); }), ITERS * 8.0, 0);
    rep("shfl.xor b32", timeit([&] { k_shfl<<<blocks, threads>>>(out); }), ITERS * 8.0, 0);
    rep("lds.64", timeit([&] { k_lds64<<<blocks, threads>>>(out); }), ITERS * 8.0, 0);
    return 0;
}
