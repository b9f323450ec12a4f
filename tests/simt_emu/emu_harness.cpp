/*
 * emu_harness.cpp -- runs the product's device code (ukf_device.cuh) on host threads.
 * TEST INFRASTRUCTURE ONLY: exercises the warp-cooperative indexing of the kernel in
 * the GPU-less build container (optionally under ASan/UBSan).  Not a CPU fallback --
 * nothing in the product links or loads this.
 */
#define UKFB_SIMT_EMU 1
#include "../../slam_pose_estimation_b200/csrc/ukf_device.cuh"
#include "../../slam_pose_estimation_b200/csrc/ukf_thread.cuh"
#include "../../slam_pose_estimation_b200/csrc/ukf_pose_fast.cuh"
#include "../../slam_pose_estimation_b200/csrc/ukf_ori_fast.cuh"

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

/* the lane-per-filter kernel (ukf_thread.cuh): one 32-thread block per tile of 32 filters */
extern "C" int emu_thread_step(int filter_kind, const StepParams* p)
{
    const unsigned grid = unsigned((p->B + TILE - 1) / TILE);
    if (filter_kind == 0)
        simt_emu::launch(ukf_thread_kernel<PoseF>, grid, TILE, sizeof(double) * TSmem<PoseF>::TOTAL, *p);
    else
        simt_emu::launch(ukf_thread_kernel<OriF>, grid, TILE, sizeof(double) * TSmem<OriF>::TOTAL, *p);
    return 0;
}

/* the structure-exploiting PoseUKF kernel (ukf_pose_fast.cuh), same tiles */
extern "C" int emu_pose_fast_step(const StepParams* p)
{
    const unsigned grid = unsigned((p->B + TILE - 1) / TILE);
    simt_emu::launch(p->tile_done ? ukf_pose_fast_kernel<true, true> : ukf_pose_fast_kernel<true, false>, grid, TILE, sizeof(double) * PF_PER_LANE * TILE, *p);
    return 0;
}

extern "C" void emu_pose_fast_fallbacks(unsigned long long* out5)
{
    for (int i = 0; i < 5; ++i) out5[i] = pf_fallbacks[i];
}

/* the structure-exploiting OrientationUKF kernel (ukf_ori_fast.cuh), same tiles */
extern "C" int emu_ori_fast_step(const StepParams* p)
{
    const unsigned grid = unsigned((p->B + TILE - 1) / TILE);
    if (p->ori_params)
        simt_emu::launch(p->tile_done ? ukf_ori_fast_kernel<true, true> : ukf_ori_fast_kernel<true, false>, grid, TILE, sizeof(double) * OF_PER_LANE * TILE, *p);
    else
        simt_emu::launch(p->tile_done ? ukf_ori_fast_kernel<false, true> : ukf_ori_fast_kernel<false, false>, grid, TILE, sizeof(double) * OF_PER_LANE * TILE, *p);
    return 0;
}

extern "C" void emu_ori_fast_fallbacks(unsigned long long* out5)
{
    for (int i = 0; i < 5; ++i) out5[i] = of_fallbacks[i];
}

/* the SO(3) kernels of so3.cuh as the host build evaluates them (same layout as ukfb_selftest_so3) */
extern "C" void emu_selftest_so3(long long n, const double* v, const double* x, double* out)
{
    for (long long i = 0; i < n; ++i) {
        double q[4], w[3], sq, rs;
        so3_exp(v + 3 * i, 1.0, q);
        so3_log(q, w);
        fast_sqrt_rsqrt(x[i], sq, rs);
        double* o = out + 21 * i;
        o[0] = q[0], o[1] = q[1], o[2] = q[2], o[3] = q[3], o[4] = w[0], o[5] = w[1], o[6] = w[2];
        o[7] = fast_rcp(x[i]), o[8] = sq, o[9] = rs;
        /* the branch-free pair of the fast kernels (ukf_pose_fast.cuh): polynomial exp, reciprocal-free log */
        double qf[4], wf[3];
        bool slow = false;
        pf_exp(v + 3 * i, 1.0, qf, slow);
        pf_log(qf, wf, slow);
        o[10] = wf[0], o[11] = wf[1], o[12] = wf[2], o[13] = slow ? 1.0 : 0.0;
        /* the any-angle pair of the fast kernels, as the pair type they use */
        D2 v2[3] = {D2(v[3 * i], -v[3 * i]), D2(v[3 * i + 1], -v[3 * i + 1]), D2(v[3 * i + 2], -v[3 * i + 2])}, q2[4], w2[3];
        bool hard = false;
        pf_exp_wide<D2>(v2, 1.0, q2, hard);
        pf_log_wide<D2>(q2, w2);
        o[14] = q2[0].a, o[15] = q2[1].a, o[16] = q2[2].a, o[17] = q2[3].a;
        o[18] = w2[0].a, o[19] = w2[1].a, o[20] = hard ? 1.0 : w2[2].a;
    }
}

extern "C" int emu_sizeof_params(void) { return int(sizeof(StepParams)); }
