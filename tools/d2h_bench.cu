// tools/d2h_bench.cu -- the box ceiling for the end-to-end figure: concurrent pinned device<->host copies on 1..N GPUs.
// One host thread per GPU, each with its own device buffer, its own page-locked host buffer (cudaHostAlloc, portable)
// and its own stream; all threads start together and issue `iters` back-to-back cudaMemcpyAsync of `bytes`.  Printed:
// aggregate GB/s per direction and with both directions at once, for every GPU count 1, 2, 4, ... up to the GPUs
// present.  The sizes of interest are the per-step copies of bench.py's e2e leg at 1 Mi PoseUKF filters: 109 051 904 B
// (B x 13 means out), 58 720 256 B (pose only), 25 165 824 B (B x 3 measurements in).
// build: nvcc -O2 -std=c++17 -o tools/d2h_bench tools/d2h_bench.cu      run: tools/d2h_bench [iters]
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#define CK(x)                                                                               \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                        \
            exit(1);                                                                        \
        }                                                                                   \
    } while (0)

struct Lane {
    int dev;
    char *d_out, *d_in, *h_out, *h_in;
    cudaStream_t s_out, s_in;
};

static double run(std::vector<Lane>& lanes, int n, size_t bytes_out, size_t bytes_in, int iters)
{
    std::atomic<int> ready(0);
    std::atomic<bool> go(false);
    std::vector<std::thread> th;
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; ++i)
        th.emplace_back([&, i] {
            Lane& l = lanes[i];
            CK(cudaSetDevice(l.dev));
            ready++;
            while (!go.load()) std::this_thread::yield();
            for (int k = 0; k < iters; ++k) {
                if (bytes_out) CK(cudaMemcpyAsync(l.h_out, l.d_out, bytes_out, cudaMemcpyDeviceToHost, l.s_out));
                if (bytes_in) CK(cudaMemcpyAsync(l.d_in, l.h_in, bytes_in, cudaMemcpyHostToDevice, l.s_in));
            }
            CK(cudaStreamSynchronize(l.s_out));
            CK(cudaStreamSynchronize(l.s_in));
        });
    while (ready.load() < n) std::this_thread::yield();
    t0 = std::chrono::steady_clock::now();
    go = true;
    for (auto& t : th) t.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char** argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 40;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    const size_t big = 109051904;
    std::vector<Lane> lanes(ndev);
    for (int i = 0; i < ndev; ++i) {
        Lane& l = lanes[i];
        l.dev = i;
        CK(cudaSetDevice(i));
        CK(cudaMalloc(&l.d_out, big));
        CK(cudaMalloc(&l.d_in, big));
        CK(cudaHostAlloc(&l.h_out, big, cudaHostAllocPortable));
        CK(cudaHostAlloc(&l.h_in, big, cudaHostAllocPortable));
        for (size_t o = 0; o < big; o += 4096) l.h_out[o] = 1, l.h_in[o] = 1; /* touch the pages */
        CK(cudaStreamCreateWithFlags(&l.s_out, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&l.s_in, cudaStreamNonBlocking));
    }
    printf("{\"gpus_present\": %d, \"iters\": %d, \"host_threads\": %u}\n", ndev, iters, std::thread::hardware_concurrency());
    const size_t outs[3] = {109051904, 58720256, 25165824};
    for (int n = 1; n <= ndev; n *= 2) {
        for (size_t b : outs) {
            run(lanes, n, b, 0, 3); /* warm up */
            const double t_out = run(lanes, n, b, 0, iters);
            const double t_in = run(lanes, n, 0, b, iters);
            const double t_both = run(lanes, n, b, 25165824, iters);
            printf("{\"gpus\": %d, \"bytes\": %zu, \"d2h_gbs_total\": %.1f, \"h2d_gbs_total\": %.1f, \"d2h_ms_per_copy\": %.3f, "
                   "\"d2h_gbs_total_with_25MB_h2d_alongside\": %.1f}\n",
                   n, b, n * double(b) * iters / t_out / 1e9, n * double(b) * iters / t_in / 1e9, t_out / iters * 1e3,
                   n * double(b) * iters / t_both / 1e9);
            fflush(stdout);
        }
    }
    return 0;
}
