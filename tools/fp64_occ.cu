// fp64_occ.cu -- DFMA throughput as a function of resident warps per SM and per-thread ILP
// (how many warps x independent chains the FP64 pipe needs before it saturates).
#include <cuda_runtime.h>
#include <stdio.h>
template <int ILP> __global__ void k(double *out, int iters, double a, double b)
{
    double r[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) r[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < ILP; ++i) r[i] = fma(r[i], a, b);
    }
    double acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc += r[i];
    if (acc == 123.456) out[0] = acc;
}
template <int ILP> void run(int warps)
{
    double *out;
    cudaMalloc(&out, 8);
    const int iters = 20000;
    k<ILP><<<148, warps * 32>>>(out, 10, 1.0000001, 1e-9);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep)
    {
        cudaEventRecord(e0);
        k<ILP><<<148, warps * 32>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double tf = 148.0 * warps * 32 * ILP * 2.0 * iters / (best * 1e-3) / 1e12;
    printf("warps/SM=%2d ILP=%2d  %.2f TFLOP/s\n", warps, ILP, tf);
    cudaFree(out);
}
int main()
{
    for (int w : {4, 8, 10, 12, 16, 24, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); run<16>(w); }
    return 0;
}
