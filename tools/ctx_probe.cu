// Developer probe: what does CUDA start-up cost on a multi-GPU box, and does it serialise across processes / threads?
//   ctx_probe procs N [lock]     N processes forked before any CUDA call, process r sees only device r; lock: 0 none, 1 around
//                                cuInit + context creation, 2 around context creation only
//   ctx_probe threads N [lock]   one process, N threads, thread r creates the primary context of device r
// Prints per rank: t(cuInit done), t(context ready), t(first kernel done), all relative to the common start.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <thread>
#include <vector>
#include <mutex>
#include <sys/file.h>
#include <sys/wait.h>
#include <unistd.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
__global__ void k(int *p) { if (p) *p = 1; }
int main(int argc, char **argv) {
    const std::string mode = argc > 1 ? argv[1] : "procs";
    const int n = argc > 2 ? atoi(argv[2]) : 1, lock = argc > 3 ? atoi(argv[3]) : 0;
    const double t0 = now();
    if (mode == "procs") {
        const int fd = open("/tmp/ctx_probe.lock", O_CREAT | O_RDWR, 0600);
        int rank = 0;
        std::vector<pid_t> kids;
        for (int r = 1; r < n; r++) { pid_t p = fork(); if (p == 0) { rank = r; kids.clear(); break; } kids.push_back(p); }
        setenv("CUDA_VISIBLE_DEVICES", std::to_string(rank).c_str(), 1);
        const int lfd = open("/tmp/ctx_probe.lock", O_RDWR);
        if (lock == 1) flock(lfd, LOCK_EX);
        cuInit(0);
        const double t1 = now();
        if (lock == 2) flock(lfd, LOCK_EX);
        cudaSetDevice(0); cudaFree(0);
        const double t2 = now();
        if (lock) flock(lfd, LOCK_UN);
        int *d; cudaMalloc(&d, 4); k<<<1, 1>>>(d); cudaDeviceSynchronize();
        const double t3 = now();
        printf("procs n=%d lock=%d rank %d: cuInit %.3f  ctx %.3f  kernel %.3f  (%s)\n", n, lock, rank, t1 - t0, t2 - t0, t3 - t0, cudaGetErrorString(cudaGetLastError()));
        fflush(stdout);
        if (rank > 0) _exit(0);
        for (pid_t p : kids) waitpid(p, nullptr, 0);
        printf("procs n=%d lock=%d all done %.3f\n", n, lock, now() - t0);
        (void)fd;
    } else {
        cuInit(0);
        const double t1 = now();
        std::mutex mu;
        std::vector<std::thread> th;
        for (int r = 0; r < n; r++) th.emplace_back([&, r]() {
            if (lock) mu.lock();
            cudaSetDevice(r); cudaFree(0);
            const double t2 = now();
            if (lock) mu.unlock();
            int *d; cudaMalloc(&d, 4); k<<<1, 1>>>(d); cudaDeviceSynchronize();
            printf("threads n=%d lock=%d dev %d: cuInit %.3f  ctx %.3f  kernel %.3f\n", n, lock, r, t1 - t0, t2 - t0, now() - t0);
        });
        for (auto &t : th) t.join();
        printf("threads n=%d lock=%d all done %.3f\n", n, lock, now() - t0);
    }
    return 0;
}
