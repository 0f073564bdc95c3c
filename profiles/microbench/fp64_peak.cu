// fp64 throughput probes for B200 (sm_100a): DFMA vs DMMA (mma.sync f64) shapes, plus
// shared-memory-fed DMMA.  Register-only loops, one CTA wave per SM.  Prints TFLOP/s.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* c, const double* a, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void k_dmma884(double* out, int iters, double a, double b) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void k_dmma1684(double* out, int iters, double a0, double b) {
    double c[NACC][4]; double a[2] = {a0, a0 * 0.5};
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma1684(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void k_dmma1688(double* out, int iters, double a0, double b0) {
    double c[NACC][4]; double a[4] = {a0, a0 * 0.5, a0 * 0.25, a0 * 2}; double b[2] = {b0, b0 * 0.5};
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma1688(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void k_dmma16816(double* out, int iters, double a0, double b0) {
    double c[NACC][4]; double a[8]; double b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a0 * (i + 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = b0 * (i + 1);
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

// DMMA m8n8k4 fed from shared memory: warp tile 32(M) x 32(N): per k4 step 4 A + 4 B LDS.64, 16 mma
__global__ void k_dmma884_smem(double* out, int iters) {
    extern __shared__ double sm[];
    const int KC = 64;  // k-chunk held in smem: A[32*nwarpM][KC], B[32][KC]
    double* As = sm;                 // [64][KC+? pad]
    double* Bs = sm + 64 * (KC + 4);
    for (int i = threadIdx.x; i < 64 * (KC + 4) * 2; i += blockDim.x) sm[i] = 1e-3 * (i % 7);
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int wm = warp & 1;  // two 32-row blocks
    double c[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { c[i][j][0] = 0; c[i][j][1] = 0; }
    const int LD = KC + 4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int k = 0; k < KC; k += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[(wm * 32 + i * 8 + (lane >> 2)) * LD + k + (lane & 3)];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(j * 8 + (lane >> 2)) * LD + k + (lane & 3)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j][0] + c[i][j][1];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) launch();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s  SMs %d  clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, 1024));
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int cps : {1, 2}) {
            if (threads * cps > 2048) continue;
            int grid = sms * cps;
            double warps = (double)grid * threads / 32;
            {
                double ms = time_ms([&] { k_dfma<16><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, 3);
                printf("DFMA x16           thr %4d cps %d : %7.2f TFLOP/s\n", threads, cps, 2.0 * 16 * iters * grid * threads / ms * 1e-9);
            }
            {
                double ms = time_ms([&] { k_dmma884<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, 3);
                printf("DMMA m8n8k4  x8    thr %4d cps %d : %7.2f TFLOP/s\n", threads, cps, 2.0 * 256 * 8 * iters * warps / ms * 1e-9);
            }
            {
                double ms = time_ms([&] { k_dmma1684<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, 3);
                printf("DMMA m16n8k4 x8    thr %4d cps %d : %7.2f TFLOP/s\n", threads, cps, 2.0 * 512 * 8 * iters * warps / ms * 1e-9);
            }
            {
                double ms = time_ms([&] { k_dmma1688<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, 3);
                printf("DMMA m16n8k8 x8    thr %4d cps %d : %7.2f TFLOP/s\n", threads, cps, 2.0 * 1024 * 8 * iters * warps / ms * 1e-9);
            }
            {
                double ms = time_ms([&] { k_dmma16816<8><<<grid, threads>>>(out, iters / 2, 1.0000001, 1e-9); }, 3);
                printf("DMMA m16n8k16 x8   thr %4d cps %d : %7.2f TFLOP/s\n", threads, cps, 2.0 * 2048 * 8 * (iters / 2) * warps / ms * 1e-9);
            }
        }
    }
    // fewer independent accumulators (latency probe)
    {
        int grid = sms, threads = 256; double warps = (double)grid * threads / 32;
        double ms = time_ms([&] { k_dmma884<1><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, 3);
        printf("DMMA m8n8k4 x1 (dep chain) thr 256: %7.2f TFLOP/s  => %.1f ns per dependent mma\n",
               2.0 * 256 * iters * warps / ms * 1e-9, ms * 1e6 / iters);
        ms = time_ms([&] { k_dmma884<2><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, 3);
        printf("DMMA m8n8k4 x2 thr 256: %7.2f TFLOP/s\n", 2.0 * 256 * 2 * iters * warps / ms * 1e-9);
        ms = time_ms([&] { k_dmma884<4><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); }, 3);
        printf("DMMA m8n8k4 x4 thr 256: %7.2f TFLOP/s\n", 2.0 * 256 * 4 * iters * warps / ms * 1e-9);
    }
    // smem-fed
    for (int threads : {128, 256, 512}) {
        int grid = sms; size_t smem = 64 * 68 * 2 * sizeof(double);
        CK(cudaFuncSetAttribute(k_dmma884_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int it2 = 2000;
        double ms = time_ms([&] { k_dmma884_smem<<<grid, threads, smem>>>(out, it2); }, 3);
        double warps = (double)grid * threads / 32;
        printf("DMMA m8n8k4 smem-fed 32x32 warp tile thr %4d: %7.2f TFLOP/s\n", threads,
               2.0 * 256 * 16 * (64 / 4) * it2 * warps / ms * 1e-9);
    }
    // sustained (~2 s) DMMA to see the power-capped rate
    {
        int grid = sms * 2, threads = 512; double warps = (double)grid * threads / 32;
        double ms = time_ms([&] { k_dmma884<8><<<grid, threads>>>(out, iters * 4, 1.0000001, 1e-9); }, 40);
        printf("DMMA m8n8k4 sustained: %7.2f TFLOP/s (%.1f ms per launch x40)\n", 2.0 * 256 * 8 * iters * 4 * warps / ms * 1e-9, ms);
        ms = time_ms([&] { k_dfma<16><<<grid, threads>>>(out, iters * 4, 1.0000001, 1e-9); }, 40);
        printf("DFMA sustained: %7.2f TFLOP/s (%.1f ms per launch x40)\n", 2.0 * 16 * iters * 4 * grid * threads / ms * 1e-9, ms);
    }
    return 0;
}
