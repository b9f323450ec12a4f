/*
 * simt_emu_rt.hpp -- minimal CUDA-thread emulation for the GPU-less build container.
 * TEST INFRASTRUCTURE ONLY (see slam_pose_estimation_b200/csrc/simt.cuh).
 *
 * Every CUDA thread of a block runs as a real host thread; __syncwarp() is a 32-thread
 * barrier.  Blocks run one after another.  Kernels must not exit part of a warp early.
 */
#ifndef SIMT_EMU_RT_HPP
#define SIMT_EMU_RT_HPP

#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <thread>
#include <vector>

struct emu_dim3 {
    unsigned x = 1, y = 1, z = 1;
};
inline thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace simt_emu {
inline thread_local std::barrier<>* warp_barrier = nullptr;
inline thread_local double* smem_ptr = nullptr;
inline thread_local double* xchg_ptr = nullptr; /* 64 doubles per warp: shuffle / mma operand exchange */
inline double* warp_xchg() { return xchg_ptr; }
inline double* smem_base() { return smem_ptr; }

template <class Kernel, class Params>
void launch(Kernel kernel, unsigned grid, unsigned block, std::size_t smem_bytes, const Params& params)
{
    const unsigned nwarps = (block + 31) / 32;
    for (unsigned b = 0; b < grid; ++b) {
        std::vector<double> smem((smem_bytes + 7) / 8 + 2, 0.0);
        std::vector<std::unique_ptr<std::barrier<>>> bars;
        for (unsigned w = 0; w < nwarps; ++w) {
            const unsigned cnt = (w + 1) * 32 <= block ? 32 : block - w * 32;
            bars.emplace_back(new std::barrier<>(cnt));
        }
        std::vector<double> xchg(64 * nwarps, 0.0);
        std::vector<std::thread> th;
        th.reserve(block);
        for (unsigned t = 0; t < block; ++t) {
            th.emplace_back([&, t, b] {
                threadIdx.x = t;
                blockIdx.x = b;
                blockDim.x = block;
                gridDim.x = grid;
                warp_barrier = bars[t / 32].get();
                smem_ptr = smem.data();
                xchg_ptr = xchg.data() + 64 * (t / 32);
                kernel(params);
            });
        }
        for (auto& x : th) x.join();
    }
}
}  // namespace simt_emu

inline void __syncwarp() { simt_emu::warp_barrier->arrive_and_wait(); }
inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v)
{
    return __atomic_fetch_add(a, v, __ATOMIC_RELAXED);
}
using std::fabs;

#endif
