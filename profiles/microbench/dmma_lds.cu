// How many warps does the fp64 tensor pipe need when every DMMA takes its B operand from shared memory?
// Emulates the inner loop of tree_kernel4: per k step NT x (LDS.64 B, DMMA.8x8x4 on an independent accumulator pair),
// A in a register.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_lds dmma_lds.cu && ./dmma_lds
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NT>
__global__ void k(double* out, int iters) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1e-9 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double acc[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n) { acc[n][0] = 0; acc[n][1] = n; }
    double a = 1.0 + lane * 1e-6;
    for (int it = 0; it < iters; ++it) {
        // the kernel's fragment addressing: 128-byte rows, 16-byte units swizzled by the row, k-permuted halves (conflict-free)
        const int x = lane >> 2, kk = lane & 3, q = it & 3, z = x ^ ((kk >> 1) << 2);
        const char* b = reinterpret_cast<const char*>(sm) + ((it >> 2) & 1) * 32768 + x * 128 + ((q ^ z) << 4) + ((kk & 1) << 3);
#pragma unroll
        for (int n = 0; n < NT; ++n) dmma(acc[n][0], acc[n][1], a, *reinterpret_cast<const double*>(b + n * 1024));
    }
    double s = 0;
#pragma unroll
    for (int n = 0; n < NT; ++n) s += acc[n][0] + acc[n][1];
    if (s == 12345.678) out[0] = s;
}
template <int NT>
void run(int warps, int ctas_per_sm) {
    int dev = 0, sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int iters = 20000;
    double* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(k<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NT><<<sms * ctas_per_sm, warps * 32, 64 * 1024>>>(d, 100);
    cudaEventRecord(e0);
    k<NT><<<sms * ctas_per_sm, warps * 32, 64 * 1024>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 256 * NT * (double)iters * warps * ctas_per_sm * sms;
    printf("NT %2d  warps/CTA %2d  CTAs/SM %d  (warps/SM %2d): %7.2f TFLOP/s  %s\n", NT, warps, ctas_per_sm, warps * ctas_per_sm, flops / ms * 1e-9,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}
int main() {
    run<25>(4, 1); run<25>(8, 1); run<13>(4, 1); run<13>(8, 1); run<13>(16, 1); run<13>(8, 2); run<7>(16, 1); run<7>(32, 1); run<25>(4, 2);
    return 0;
}
