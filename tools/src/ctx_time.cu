// Times the pieces of CUDA start-up that b2pt_create goes through (tools/time_phases.py shows 2-6 s there).
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    double t0 = now(), t;
    int n = 0;
    cudaGetDeviceCount(&n); t = now(); printf("cudaGetDeviceCount (driver init)   %8.1f ms\n", t - t0); t0 = t;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); t = now(); printf("cudaGetDeviceProperties            %8.1f ms\n", t - t0); t0 = t;
    cudaSetDevice(0); cudaFree(0); t = now(); printf("cudaSetDevice + cudaFree(0) (ctx)  %8.1f ms\n", t - t0); t0 = t;
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking); t = now(); printf("cudaStreamCreate                   %8.1f ms\n", t - t0); t0 = t;
    cudaEvent_t e; cudaEventCreate(&e); t = now(); printf("cudaEventCreate                    %8.1f ms\n", t - t0); t0 = t;
    void *d; cudaMalloc(&d, 4096); t = now(); printf("cudaMalloc 4 KB                    %8.1f ms\n", t - t0); t0 = t;
    void *h; cudaMallocHost(&h, 4096); t = now(); printf("cudaMallocHost 4 KB                %8.1f ms\n", t - t0); t0 = t;
    void *big; cudaError_t er = cudaMalloc(&big, (size_t)82 << 30); t = now(); printf("cudaMalloc 82 GB (%s)     %8.1f ms\n", cudaGetErrorString(er), t - t0); t0 = t;
    cudaMemset(big, 0, 1 << 20); cudaDeviceSynchronize(); t = now(); printf("first memset + sync                %8.1f ms\n", t - t0); t0 = t;
    cudaFree(big); t = now(); printf("cudaFree 82 GB                     %8.1f ms\n", t - t0); t0 = t;
    return 0;
}
