// Peak issue rate of mma.sync.m8n8k4.f64 on sm_100a as a function of warps per SM sub-partition:
// every warp runs chains of DMMAs on NACC independent accumulator blocks, no memory traffic.
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void dmma_loop(double *out, int iters) {
    double acc[NACC][2];
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[i][0]), "+d"(acc[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
    if (s == 12345.678) out[0] = s;
}

template <int NACC> void run(int warps_per_block, int blocks_per_sm) {
    int sms = 148;
    double *out;
    cudaMalloc(&out, 8);
    const int iters = 20000 / NACC * 8;
    dim3 grid(sms * blocks_per_sm), block(32 * warps_per_block);
    dmma_loop<NACC><<<grid, block>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    dmma_loop<NACC><<<grid, block>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 8 * 8 * 4 * NACC * (double)iters * warps_per_block * blocks_per_sm * sms;
    printf("NACC=%2d warps/block=%d blocks/SM=%d warps/SMSP=%.1f : %.2f TFLOP/s\n", NACC, warps_per_block,
           blocks_per_sm, warps_per_block * blocks_per_sm / 4.0, flops / ms / 1e9);
    cudaFree(out);
}

int main() {
    run<16>(4, 1);
    run<16>(4, 2);
    run<16>(4, 3);
    run<16>(4, 4);
    run<16>(8, 2);
    run<4>(4, 2);
    run<4>(4, 4);
    run<4>(4, 8);
    run<64>(4, 1);
    run<64>(4, 2);
    run<1>(4, 8);
    run<2>(4, 8);
    return 0;
}
