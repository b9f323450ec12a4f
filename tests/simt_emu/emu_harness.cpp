/*
 * emu_harness.cpp -- runs the product's device code (ukf_device.cuh) on host threads.
 * TEST INFRASTRUCTURE ONLY: exercises the warp-cooperative indexing of the kernel in
 * the GPU-less build container (optionally under ASan/UBSan).  Not a CPU fallback --
 * nothing in the product links or loads this.
 */
#define UKFB_SIMT_EMU 1
#include "../../slam_pose_estimation_b200/csrc/ukf_device.cuh"

using namespace ukfb;

constexpr int EMU_WPB = 2;

template <class F, int G>
static void run(const StepParams& p)
{
    const long long per_block = (long long)EMU_WPB * G;
    const unsigned grid = unsigned((p.B + per_block - 1) / per_block);
    simt_emu::launch(ukf_step_kernel<F, G, EMU_WPB, 1>, grid, EMU_WPB * 32, sizeof(double) * EMU_WPB * Smem<F, G>::TOTAL, p);
}

extern "C" int emu_step(int filter_kind, int G, const StepParams* p)
{
    if (filter_kind == 0) {
        switch (G) {
            case 4: run<PoseF, 4>(*p); return 0;
            case 8: run<PoseF, 8>(*p); return 0;
            case 16: run<PoseF, 16>(*p); return 0;
        }
    } else {
        switch (G) {
            case 4: run<OriF, 4>(*p); return 0;
            case 8: run<OriF, 8>(*p); return 0;
            case 16: run<OriF, 16>(*p); return 0;
        }
    }
    return -1;
}

extern "C" int emu_sizeof_params(void) { return int(sizeof(StepParams)); }
