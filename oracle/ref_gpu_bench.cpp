// TEST / MEASUREMENT INFRASTRUCTURE -- never linked into the product.
//
// On-box GPU comparator: the UNMODIFIED reference headers compiled in
// CUDA mode (thrust + cuBLAS, -arch=sm_100) and timed on the same B200 on the bench workloads:
// the label permutation, a periodic shift and the distillation contraction (BASELINE configs 1/2/5
// shapes).  Built by oracle/Makefile into oracle/_ref/ref_gpu_bench where /root/reference exists;
// bench.py runs it (N = 1 only) and reports its numbers beside ours as `reference_gpu`.
//
// What it calls is the reference's public API exactly as SURVEY.md Appendix A does (dist.h:3583 copy,
// dist.h:3701 contraction, platform.h:809 createGpuContext, blas.h:965 sync); this file contains only
// the timing harness.
#include "superbblas.h"
#include <chrono>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace superbblas;
using CD = std::complex<double>;
using CF = std::complex<float>;

#define CK(x)                                                                                      \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) {                                                                   \
            std::fprintf(stderr, "CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__);  \
            std::exit(1);                                                                          \
        }                                                                                          \
    } while (0)

template <typename R> __global__ void fill_kernel(R *p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned x = (unsigned)(i * 2654435761u) ^ (unsigned)(i >> 32);
        x ^= x >> 15, x *= 2246822519u, x ^= x >> 13;
        p[i] = (R)((double)(x & 0xffffff) / 8388608.0 - 1.0); // uniform in [-1, 1)
    }
}

template <typename T> T *device_tensor(size_t n) {
    T *p = nullptr;
    CK(cudaMalloc(&p, n * sizeof(T)));
    using R = typename T::value_type;
    fill_kernel<R><<<1184, 256>>>((R *)p, 2 * n);
    CK(cudaDeviceSynchronize());
    return p;
}

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

/// cold = first call (the reference builds and caches its index vectors there), warm = mean of reps
template <typename F> static void time_it(F f, const Context &ctx, int reps, double &cold_ms, double &warm_ms) {
    sync(ctx);
    double t = now();
    f();
    sync(ctx);
    cold_ms = (now() - t) * 1e3;
    f();
    sync(ctx);
    t = now();
    for (int i = 0; i < reps; ++i) f();
    sync(ctx);
    warm_ms = (now() - t) * 1e3 / reps;
}

static bool first_item = true;
static void emit(const std::string &name, double cold_ms, double warm_ms, double work, const char *unit) {
    std::printf("%s\"%s\": {\"ms\": %.6g, \"cold_ms\": %.6g, \"%s\": %.6g}", first_item ? "" : ", ",
                name.c_str(), warm_ms, cold_ms, unit, work / (warm_ms * 1e-3));
    first_item = false;
    std::fflush(stdout);
}

template <typename T> static void bench_permute(const Context &ctx, int L, int Lt, int reps, const char *tag) {
    Coor<6> dim0{L, L, L, Lt, 4, 3}, dim1{3, 4, Lt, L, L, L};
    PartitionItem<6> p0{Coor<6>{}, dim0}, p1{Coor<6>{}, dim1};
    const size_t n = (size_t)L * L * L * Lt * 12;
    T *a = device_tensor<T>(n), *b = device_tensor<T>(n);
    const T *src = a;
    T *dst = b;
    double cold, warm;
    time_it([&] {
        copy<6, 6>(T{1}, &p0, 1, "xyztsc", {}, dim0, dim0, &src, nullptr, &ctx, &p1, 1, "cstzyx", {}, dim1,
                   &dst, nullptr, &ctx, FastToSlow, Copy);
    }, ctx, reps, cold, warm);
    emit(std::string("permute_xyztsc_cstzyx_") + tag, cold, warm, 2.0 * n * sizeof(T) / 1e9, "GB/s");
    CK(cudaFree(a));
    CK(cudaFree(b));
}

template <typename T> static void bench_shift(const Context &ctx, int mu, int reps, const char *tag) {
    Coor<6> dim{64, 64, 32, 32, 4, 3};
    PartitionItem<6> p{Coor<6>{}, dim};
    const size_t n = (size_t)64 * 64 * 32 * 32 * 12;
    T *a = device_tensor<T>(n), *b = device_tensor<T>(n);
    const T *src = a;
    T *dst = b;
    Coor<6> from1{};
    from1[mu] = 1;
    double cold, warm;
    time_it([&] {
        copy<6, 6>(T{1}, &p, 1, "xyztsc", {}, dim, dim, &src, nullptr, &ctx, &p, 1, "xyztsc", from1, dim, &dst,
                   nullptr, &ctx, FastToSlow, Copy);
    }, ctx, reps, cold, warm);
    emit(std::string("shift_") + "xyzt"[mu] + "_" + tag, cold, warm, 2.0 * n * sizeof(T) / 1e9, "GB/s");
    CK(cudaFree(a));
    CK(cudaFree(b));
}

template <typename T> static void bench_contraction(const Context &ctx, int L, int Lt, int nv, int reps, const char *tag) {
    Coor<6> dimv{3, L, L, L, Lt, nv};
    Coor<3> dimr{Lt, nv, nv};
    PartitionItem<6> pv{Coor<6>{}, dimv};
    PartitionItem<3> pr{Coor<3>{}, dimr};
    const size_t n = (size_t)3 * L * L * L * Lt * nv, nr = (size_t)Lt * nv * nv;
    T *a = device_tensor<T>(n), *b = device_tensor<T>(n), *c = device_tensor<T>(nr);
    const T *pa = a, *pb = b;
    T *pc = c;
    double cold, warm;
    time_it([&] {
        contraction<6, 6, 3>(T{1}, &pv, {}, dimv, dimv, 1, "cxyztn", true, &pa, &ctx, &pv, {}, dimv, dimv, 1,
                             "cxyztm", false, &pb, &ctx, T{0}, &pr, {}, dimr, dimr, 1, "tnm", &pc, &ctx,
                             FastToSlow);
    }, ctx, reps, cold, warm);
    emit(std::string("contraction_") + tag, cold, warm, 8.0 * Lt * nv * nv * 3.0 * L * L * L / 1e12, "TFLOP/s");
    CK(cudaFree(a));
    CK(cudaFree(b));
    CK(cudaFree(c));
}

int main(int argc, char **argv) {
    int reps = 5;
    for (int i = 1; i < argc; ++i)
        if (std::strncmp(argv[i], "--reps=", 7) == 0) reps = std::atoi(argv[i] + 7);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        std::printf("{\"unavailable\": \"no CUDA device\"}\n");
        return 0;
    }
    Context ctx = createGpuContext(0);
    std::printf("{");
    try {
        bench_permute<CD>(ctx, 32, 64, reps, "c128");
        bench_permute<CF>(ctx, 32, 64, reps, "c64");
        bench_shift<CF>(ctx, 0, reps, "c64");
        bench_shift<CF>(ctx, 3, reps, "c64");
        bench_shift<CD>(ctx, 0, reps, "c128");
        bench_shift<CD>(ctx, 3, reps, "c128");
        clearCaches();
        bench_contraction<CD>(ctx, 32, 64, 64, reps, "config2_c128");
        bench_contraction<CF>(ctx, 32, 64, 64, reps, "config2_c64");
    } catch (const std::exception &e) {
        std::printf("%s\"error\": \"%s\"", first_item ? "" : ", ", e.what());
    }
    std::printf(", \"how\": \"unmodified reference headers, CUDA mode (thrust + cuBLAS), nvcc -O3 -arch=sm_100, "
                "one B200, %d warm repetitions after 2 untimed calls\"}\n", reps);
    return 0;
}
