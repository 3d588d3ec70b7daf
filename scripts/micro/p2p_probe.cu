// Micro-benchmark for the next round's NVLink question: what do kernel-issued peer
// stores achieve between two B200s of this pool, as a function of the grid, the bytes per thread and
// the direction, next to the copy engine?  One process, two devices, plain peer access.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/micro/p2p_probe.cu -o scripts/micro/p2p_probe
//   gpurun --gpus 2 -- scripts/micro/p2p_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x)                                                                                      \
    do {                                                                                           \
        cudaError_t e = (x);                                                                       \
        if (e != cudaSuccess) {                                                                    \
            std::printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__);            \
            std::exit(1);                                                                          \
        }                                                                                          \
    } while (0)

// persistent grid-stride copy, V = 4, 8 or 16 bytes per thread and access
template <typename V> __global__ void copy_kernel(const V *__restrict__ src, V *dst, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}
// the same with 4 independent accesses in flight per thread
template <typename V> __global__ void copy4_kernel(const V *__restrict__ src, V *dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        V a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a, dst[i + stride] = b, dst[i + 2 * stride] = c, dst[i + 3 * stride] = d;
    }
    for (; i < n; i += stride) dst[i] = src[i];
}

template <typename F> static double time_ms(F f, cudaStream_t s, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) f();
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(e0, s));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main() {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) {
        std::printf("needs 2 GPUs\n");
        return 0;
    }
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    std::printf("peer access 0->1: %d\n", can);
    const size_t bytes = 1ull << 30; // 1 GiB per transfer (> L2)
    char *loc0, *loc0b, *rem1;
    CK(cudaSetDevice(1));
    CK(cudaMalloc(&rem1, bytes));
    CK(cudaDeviceEnablePeerAccess(0, 0));
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    CK(cudaMalloc(&loc0, bytes));
    CK(cudaMalloc(&loc0b, bytes));
    CK(cudaMemset(loc0, 1, bytes));
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    auto report = [&](const char *what, int grid, int vec, double ms) {
        std::printf("%-34s grid %4d  %2d B/thread  %8.3f ms  %7.1f GB/s\n", what, grid, vec, ms, bytes / ms / 1e6);
    };
    report("copy engine local -> peer", 0, 0,
           time_ms([&] { CK(cudaMemcpyPeerAsync(rem1, 1, loc0, 0, bytes, s)); }, s, 5));
    report("copy engine peer -> local", 0, 0,
           time_ms([&] { CK(cudaMemcpyPeerAsync(loc0b, 0, rem1, 1, bytes, s)); }, s, 5));
    const int grids[] = {37, 74, 148, 296, 592, 1184};
    for (int g : grids) {
        report("kernel local -> local (uint4)", g, 16,
               time_ms([&] { copy_kernel<uint4><<<g, 256, 0, s>>>((const uint4 *)loc0, (uint4 *)loc0b, bytes / 16); }, s, 5));
        report("kernel STORE to peer (uint4)", g, 16,
               time_ms([&] { copy_kernel<uint4><<<g, 256, 0, s>>>((const uint4 *)loc0, (uint4 *)rem1, bytes / 16); }, s, 5));
        report("kernel STORE to peer x4 (uint4)", g, 16,
               time_ms([&] { copy4_kernel<uint4><<<g, 256, 0, s>>>((const uint4 *)loc0, (uint4 *)rem1, bytes / 16); }, s, 5));
        report("kernel STORE to peer (uint2)", g, 8,
               time_ms([&] { copy_kernel<uint2><<<g, 256, 0, s>>>((const uint2 *)loc0, (uint2 *)rem1, bytes / 8); }, s, 5));
        report("kernel LOAD from peer x4 (uint4)", g, 16,
               time_ms([&] { copy4_kernel<uint4><<<g, 256, 0, s>>>((const uint4 *)rem1, (uint4 *)loc0b, bytes / 16); }, s, 5));
    }
    // both directions at once (what an exchange does): device 1 stores into device 0 meanwhile
    CK(cudaSetDevice(1));
    char *loc1;
    CK(cudaMalloc(&loc1, bytes));
    CK(cudaMemset(loc1, 2, bytes));
    cudaStream_t s1;
    CK(cudaStreamCreate(&s1));
    for (int g : {74, 148, 296}) {
        CK(cudaSetDevice(1));
        for (int i = 0; i < 12; ++i) copy4_kernel<uint4><<<g, 256, 0, s1>>>((const uint4 *)loc1, (uint4 *)loc0b, bytes / 16);
        CK(cudaSetDevice(0));
        report("bidirectional: STORE to peer x4", g, 16,
               time_ms([&] { copy4_kernel<uint4><<<g, 256, 0, s>>>((const uint4 *)loc0, (uint4 *)rem1, bytes / 16); }, s, 5));
        CK(cudaSetDevice(1));
        CK(cudaStreamSynchronize(s1));
    }
    std::printf("done\n");
    return 0;
}
