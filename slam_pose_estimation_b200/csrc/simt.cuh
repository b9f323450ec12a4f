/*
 * simt.cuh -- the few macros the device code is written against.
 *
 * Product build: nvcc, sm_100a; everything maps to the CUDA builtins.
 *
 * UKFB_SIMT_EMU build (tests/simt_emu only, never linked into the product library):
 * the same kernel source is compiled by g++ and each CUDA thread of a block is run as
 * a real host thread, with __syncwarp()/__syncthreads() as barriers.  It exists so
 * that the warp-cooperative indexing (shared-memory maps, tile ownership, hazards)
 * can be exercised under -fsanitize=address,undefined in the GPU-less build
 * container.  It is a development/test harness, not a CPU fallback: the C ABI
 * (ukf_batch.cu) is CUDA-only and fails loudly without a device.
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
#define UKFB_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
#define UKFB_SMEM_DECL extern __shared__ __align__(16) double ukfb_smem[];
#define UKFB_LDG(p) __ldg(p)
#define UKFB_UNROLL _Pragma("unroll")
#define UKFB_NOUNROLL _Pragma("unroll 1")

namespace ukfb {
UKFB_HD void ukfb_sincos(double x, double* s, double* c) { sincos(x, s, c); }
}

#else /* ---------------------------- host emulation ---------------------------- */

#include <cmath>
#include <cstdint>

#include "simt_emu_rt.hpp" /* tests/simt_emu: threadIdx, __syncwarp, atomicAdd, ... */

#define UKFB_HD inline
#define UKFB_D inline
#define UKFB_DNI inline
#define UKFB_GLOBAL
#define UKFB_LAUNCH_BOUNDS(t, b)
#define UKFB_SMEM_DECL double* ukfb_smem = ::simt_emu::smem_base();
#define UKFB_LDG(p) (*(p))
#define UKFB_UNROLL
#define UKFB_NOUNROLL

namespace ukfb {
using std::atan;
using std::sqrt;
using std::fabs;
inline void ukfb_sincos(double x, double* s, double* c)
{
    *s = std::sin(x);
    *c = std::cos(x);
}
}

#endif

#endif /* UKFB_SIMT_CUH */
