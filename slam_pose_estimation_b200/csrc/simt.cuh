/*
 * simt.cuh -- the few macros and warp primitives the device code is written against.
 *
 * Product build: nvcc, sm_100a; everything maps to the CUDA builtins / PTX:
 *   warp_shfl        shfl.sync.idx (two 32-bit halves of a double)
 *   warp_dmma        mma.sync.aligned.m8n8k4.row.col.f64 (FP64 tensor-core tile, SASS DMMA.8x8x4)
 *   rcp_seed/rsq_seed  rcp.approx.ftz.f64 / rsqrt.approx.ftz.f64 (MUFU.RCP64H / MUFU.RSQ64H)
 *
 * UKFB_SIMT_EMU build (tests/simt_emu only, never linked into the product library):
 * the same kernel source is compiled by g++ and each CUDA thread of a block is run as
 * a real host thread, with __syncwarp() as a barrier and the warp primitives emulated
 * through a per-warp exchange buffer.  It exists so that the warp-cooperative indexing
 * (shared-memory maps, tile ownership, hazards) can be exercised under
 * -fsanitize=address,undefined in the GPU-less build container.  It is a
 * development/test harness, not a CPU fallback: the C ABI (ukf_batch.cu) is CUDA-only
 * and fails loudly without a device.
 */
#ifndef UKFB_SIMT_CUH
#define UKFB_SIMT_CUH

#ifndef UKFB_SIMT_EMU

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define UKFB_HD __host__ __device__ __forceinline__
#define UKFB_D __device__ __forceinline__
#define UKFB_DNI __device__ __noinline__
#define UKFB_GLOBAL __global__
#define UKFB_GRID_CONSTANT __grid_constant__
#define UKFB_CONSTANT __constant__
#define UKFB_LDCG(p) __ldcg(p)
#define UKFB_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
#define UKFB_SMEM_DECL extern __shared__ __align__(16) double ukfb_smem[];
#define UKFB_LDG(p) __ldg(p)
#define UKFB_UNROLL _Pragma("unroll")
#define UKFB_NOUNROLL _Pragma("unroll 1")

namespace ukfb {

UKFB_D void ukfb_sincos(double x, double* s, double* c) { sincos(x, s, c); }

/* hint: bring the line at p into L2 */
UKFB_D void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

/* value of `v` held by lane `src` (all 32 lanes must call) */
UKFB_D double warp_shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

/* upper 32 bits of a double: for positive finite values (and +inf, NaN above them) ordered like the values themselves, so
 * a range check can be an integer comparison instead of an instruction of the FP64 pipe */
UKFB_D int hi_word(double x) { return __double2hiint(x); }

/* Programmatic dependent launch: the next kernel of the stream (if it was launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization) may be scheduled once every block of this grid has issued this or
 * exited -- i.e. into the slots the grid's last, partial wave leaves empty.  Ordering of the DATA is then the kernels' own
 * business: tile_done_add / tile_done_wait below. */
UKFB_D void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
/* every lane of the warp that owned a tile in a launch adds 1 to the tile's counter once its stores are issued (release);
 * every lane of the warp that owns it in the next launch waits until the counter shows all 32 (acquire): one
 * release/acquire pair per lane pair through the read-modify-write chain on the counter */
UKFB_D void tile_done_add(unsigned long long* counter)
{
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(counter), "l"(1ull) : "memory");
}
UKFB_D void tile_done_wait(const unsigned long long* counter, unsigned long long need)
{
    unsigned long long v;
    for (unsigned spins = 0;; ++spins) {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(counter) : "memory");
        if (v >= need) break;
        if (spins > 40000000u) __trap(); /* ~10 s: the previous launch never finished this tile -- fail, do not hang */
        __nanosleep(200);
    }
}

/* C(8x8) += A(8x4) * B(4x8) on the FP64 tensor-core path.  Fragment layout (PTX m8n8k4.f64):
 * lane l holds A[l/4][l%4], B[l%4][l/4], C[l/4][2*(l%4)] and C[l/4][2*(l%4)+1]. */
UKFB_D void warp_dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

/* 1/x and 1/sqrt(x) for normal, positive-or-negative (rcp) / positive (rsq) x: hardware seed
 * (about 20 bits) refined by Newton steps in FMA arithmetic; error below 1 ulp, no
 * special-case branches (callers guarantee the operand range).  Two quadratic steps take 20 bits beyond
 * 53 (the third one the first version carried changed nothing: tests/test_so3_kernels.py on the device). */
#ifndef UKFB_SEED_REFINEMENTS
#define UKFB_SEED_REFINEMENTS 2
#endif
UKFB_D double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
#if UKFB_SEED_REFINEMENTS >= 3
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
#endif
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

/* s = sqrt(x), r = 1/sqrt(x) (Goldschmidt from the hardware seed, one correction of s) */
UKFB_D void fast_sqrt_rsqrt(double x, double& s, double& r)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double e = fma(-g, h, 0.5);
    g = fma(g, e, g), h = fma(h, e, h);
#if UKFB_SEED_REFINEMENTS >= 3
    e = fma(-g, h, 0.5);
    g = fma(g, e, g), h = fma(h, e, h);
#endif
    e = fma(-g, h, 0.5);
    g = fma(g, e, g), h = fma(h, e, h);
    const double d = fma(-g, g, x);
    s = fma(d, h, g);
    r = h + h;
}

} /* namespace ukfb */

#else /* ---------------------------- host emulation ---------------------------- */

#include <cmath>
#include <cstdint>

#include "simt_emu_rt.hpp" /* tests/simt_emu: threadIdx, __syncwarp, atomicAdd, warp exchange buffer */

#define UKFB_HD inline
#define UKFB_D inline
#define UKFB_DNI inline
#define UKFB_GLOBAL
#define UKFB_GRID_CONSTANT
#define UKFB_CONSTANT static const
#define UKFB_LDCG(p) (*(p))
#define UKFB_LAUNCH_BOUNDS(t, b)
#define UKFB_SMEM_DECL double* ukfb_smem = ::simt_emu::smem_base();
#define UKFB_LDG(p) (*(p))
#define UKFB_UNROLL
#define UKFB_NOUNROLL

namespace ukfb {
using std::atan;
using std::fabs;
using std::fma;
using std::sqrt;
inline void prefetch_l2(const void*) {}
inline int hi_word(double x)
{
    long long b;
    static_assert(sizeof(b) == sizeof(x), "");
    __builtin_memcpy(&b, &x, sizeof(b));
    return int(b >> 32);
}
inline void pdl_launch_dependents() {}
inline void tile_done_add(unsigned long long* counter) { __atomic_fetch_add(counter, 1ull, __ATOMIC_RELEASE); }
inline void tile_done_wait(const unsigned long long* counter, unsigned long long need)
{
    while (__atomic_load_n(counter, __ATOMIC_ACQUIRE) < need) {
    }
}
inline void ukfb_sincos(double x, double* s, double* c)
{
    *s = std::sin(x);
    *c = std::cos(x);
}
inline double warp_shfl(double v, int src)
{
    double* x = ::simt_emu::warp_xchg();
    const int lane = threadIdx.x & 31;
    x[lane] = v;
    __syncwarp();
    const double r = x[src & 31];
    __syncwarp();
    return r;
}
inline void warp_dmma(double& c0, double& c1, double a, double b)
{
    double* x = ::simt_emu::warp_xchg();
    const int lane = threadIdx.x & 31;
    x[lane] = a;
    x[32 + lane] = b;
    __syncwarp();
    const int row = lane >> 2, col = 2 * (lane & 3);
    for (int k = 0; k < 4; ++k) {
        c0 += x[row * 4 + k] * x[32 + col * 4 + k];
        c1 += x[row * 4 + k] * x[32 + (col + 1) * 4 + k];
    }
    __syncwarp();
}
inline double fast_rcp(double x) { return 1.0 / x; }
inline void fast_sqrt_rsqrt(double x, double& s, double& r)
{
    s = std::sqrt(x);
    r = 1.0 / s;
}
}  // namespace ukfb

#endif

#endif /* UKFB_SIMT_CUH */
