// gpuinit_probe.cu -- where does CUDA start-up time go on a multi-GPU box?  (development aid)
//   gpuinit_probe seq      : cudaGetDeviceCount, then one primary context after the other
//   gpuinit_probe threads  : contexts created by one thread per device at once
//   gpuinit_probe procs    : one forked child per device (CUDA_VISIBLE_DEVICES=i) creates a context and holds it
//                            until the parent has created its own contexts on all devices
// Prints wall-clock stamps in ms.  Build: nvcc -O2 -o tools/gpuinit_probe tools/gpuinit_probe.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <sys/wait.h>
#include <unistd.h>

static std::chrono::steady_clock::time_point t0;
static double ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }

int main(int argc, char **argv)
{
    t0 = std::chrono::steady_clock::now();
    const std::string mode = argc > 1 ? argv[1] : "seq";
    const int want = argc > 2 ? atoi(argv[2]) : 64;
    std::vector<pid_t> kids;
    int pipes[64][2];
    if (mode == "procs") {
        // children first, before this process touches CUDA
        for (int i = 0; i < want && i < 64; i++) {
            if (pipe(pipes[i])) return 1;
            pid_t p = fork();
            if (p == 0) {
                char v[16];
                snprintf(v, sizeof v, "%d", i);
                setenv("CUDA_VISIBLE_DEVICES", v, 1);
                for (int j = 0; j <= i; j++)
                    close(pipes[j][1]); // also the write ends inherited from earlier forks, or no child ever sees EOF
                int n = 0;
                if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) {
                    cudaSetDevice(0);
                    cudaFree(0);
                    printf("[%8.1f] child %d: context ready\n", ms(), i);
                    fflush(stdout);
                }
                char c;
                (void)!read(pipes[i][0], &c, 1); // hold the device open until the parent is done
                _exit(0);
            }
            close(pipes[i][0]);
            kids.push_back(p);
        }
        usleep(1000 * (argc > 3 ? atoi(argv[3]) : 0));
    }
    int n = 0;
    cudaGetDeviceCount(&n);
    printf("[%8.1f] cudaGetDeviceCount -> %d\n", ms(), n);
    n = n < want ? n : want;
    if (mode == "threads" || mode == "procs") {
        std::vector<std::thread> th;
        for (int i = 0; i < n; i++)
            th.emplace_back([i] {
                cudaSetDevice(i);
                cudaFree(0);
                printf("[%8.1f] thread %d: context ready\n", ms(), i);
                fflush(stdout);
            });
        for (auto &t : th) t.join();
    } else {
        for (int i = 0; i < n; i++) {
            cudaSetDevice(i);
            cudaFree(0);
            printf("[%8.1f] device %d: context ready\n", ms(), i);
        }
    }
    printf("[%8.1f] all contexts ready (%s)\n", ms(), mode.c_str());
    for (size_t i = 0; i < kids.size(); i++) {
        close(pipes[i][1]);
        waitpid(kids[i], nullptr, 0);
    }
    printf("[%8.1f] exit\n", ms());
    return 0;
}
