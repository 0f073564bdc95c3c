// TMEM as warp-private scratch (sm_100a): every thread parks 2*NT doubles in tensor memory with tcgen05.st and
// reads them back with tcgen05.ld (32x32b shape: thread i of warp w <-> TMEM lane 32*(w%4)+i, its own columns).
// Checks the round trip and times st / ld against the same traffic through L2 (global scratch).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_scratch tmem_scratch.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const double (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
                 :: "r"(taddr),
                    "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
                    "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3]))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, double (&v)[4]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

constexpr int NT = 26;   // 2*NT doubles per thread = 104 columns

__global__ void __launch_bounds__(256, 1) k_tmem(double* out, int* bad, int iters, long long* cyc) {
    __shared__ uint32_t tbase_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" :: "r"((uint32_t)__cvta_generic_to_shared(&tbase_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t tbase = tbase_s;
    // warp w: lanes 32*(w%4).., columns (w/4)*256 ..
    const uint32_t my = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 256);
    double v[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n) { v[n][0] = (double)(threadIdx.x + 1000 * n + 100000 * (int)blockIdx.x); v[n][1] = -v[n][0] * 3.0; }
    // correctness: two regions (chain at col 0, slot at col 104)
#pragma unroll
    for (int c = 0; c < NT / 2; ++c) {
        double t[4] = {v[2 * c][0], v[2 * c][1], v[2 * c + 1][0], v[2 * c + 1][1]};
        tmem_st8(my + 8 * c, t);
        double u[4] = {t[0] * 2, t[1] * 2, t[2] * 2, t[3] * 2};
        tmem_st8(my + 104 + 8 * c, u);
    }
    tmem_wait_st();
    int nbad = 0;
    for (int c = NT / 2 - 1; c >= 0; --c) {   // dynamic column index
        double t[4], u[4];
        tmem_ld8(my + 8 * c, t);
        tmem_ld8(my + 104 + 8 * c, u);
        const double e[4] = {(double)(threadIdx.x + 1000 * (2 * c) + 100000 * (int)blockIdx.x), 0, (double)(threadIdx.x + 1000 * (2 * c + 1) + 100000 * (int)blockIdx.x), 0};
        if (t[0] != e[0] || t[1] != -e[0] * 3.0 || t[2] != e[2] || t[3] != -e[2] * 3.0) ++nbad;
        if (u[0] != 2 * t[0] || u[1] != 2 * t[1] || u[2] != 2 * t[2] || u[3] != 2 * t[3]) ++nbad;
    }
    if (nbad) atomicAdd(bad, nbad);
    __syncthreads();
    // timing: st of 104 columns, then ld of 104 columns, iters times
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < NT / 2; ++c) {
            double t[4] = {v[2 * c][0], v[2 * c][1], v[2 * c + 1][0], v[2 * c + 1][1]};
            tmem_st8(my + 8 * c, t);
        }
        tmem_wait_st();
    }
    long long t1 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int c = 0; c < NT / 2; ++c) {
            double t[4];
            tmem_ld8(my + 8 * c, t);
            v[0][0] += t[0] + t[1] + t[2] + t[3];
        }
    }
    long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
    double s = 0;
#pragma unroll
    for (int n = 0; n < NT; ++n) s += v[n][0] + v[n][1];
    if (s == 12345.678) out[0] = s;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" :: "r"(tbase));
}

// the same traffic through global memory (L2): fragment layout [n][lane] double2, thread-private
__global__ void __launch_bounds__(256, 1) k_l2(double* scratch, double* out, int iters, long long* cyc) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* my = reinterpret_cast<double2*>(scratch + ((size_t)blockIdx.x * 8 + warp) * NT * 64) + lane;
    double v[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n) { v[n][0] = 1e-3 * threadIdx.x + n; v[n][1] = -v[n][0]; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int n = 0; n < NT; ++n) my[n * 32] = make_double2(v[n][0] + it, v[n][1]);
        __syncwarp();
    }
    __threadfence_block();
    long long t1 = clock64();
    for (int it = 0; it < iters; ++it) {
        double2 t[NT];
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];\n" : "=d"(t[n].x), "=d"(t[n].y) : "l"(&my[n * 32]) : "memory");
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) v[0][0] += t[n].x + t[n].y;
    }
    long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
    double s = 0;
#pragma unroll
    for (int n = 0; n < NT; ++n) s += v[n][0] + v[n][1];
    if (s == 12345.678) out[0] = s;
}

int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    double* out; int* bad; long long* cyc; double* scratch;
    CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&bad, 4)); CK(cudaMalloc(&cyc, 16)); CK(cudaMemset(bad, 0, 4));
    CK(cudaMalloc(&scratch, (size_t)sms * 8 * NT * 64 * 8));
    const int iters = 2000;
    k_tmem<<<sms, 256>>>(out, bad, iters, cyc);
    CK(cudaDeviceSynchronize());
    int hbad; long long hc[2];
    CK(cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost));
    const double bytes = 8.0 * NT * 64 * 8;   // per CTA per iteration (8 warps x 13 KB)
    printf("TMEM scratch round trip: %s (%d mismatches)\n", hbad ? "FAILED" : "ok", hbad);
    printf("TMEM st: %.1f cycles per 8-warp burst of %.0f B  => %.1f B/clk/SM\n", (double)hc[0] / iters, bytes, bytes * iters / hc[0]);
    printf("TMEM ld: %.1f cycles per 8-warp burst of %.0f B  => %.1f B/clk/SM (x8 loads, each waited for)\n", (double)hc[1] / iters, bytes, bytes * iters / hc[1]);
    k_l2<<<sms, 256>>>(scratch, out, iters, cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost));
    printf("L2   st: %.1f cycles per burst => %.1f B/clk/SM\n", (double)hc[0] / iters, bytes * iters / hc[0]);
    printf("L2   ld: %.1f cycles per burst => %.1f B/clk/SM\n", (double)hc[1] / iters, bytes * iters / hc[1]);
    return hbad != 0;
}
