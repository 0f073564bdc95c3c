// Device side of the transition-table generator.
//
// The reference builds its tables offline as chains of (N+1)^3 fp64 matrix products (mope/generate_transitions.py:217-228,
// :386-414 matrix powers of the one-generation Wright-Fisher matrix; :29-54 + mope/_transition.pyx:6-33 bottleneck
// products) followed by the symmetric frequency binning (:194-208).  Here:
//   * binom_matrix_kernel  rows of binomial pmfs (the Wright-Fisher matrix :57-67 and the bottleneck factors);
//   * dgemm_tn_kernel      C = A . Bt^T on the fp64 tensor pipe (mma.sync.m8n8k4.f64; tcgen05 has no f64 kind), operands
//                          streamed by TMA into a shared-memory ring exactly like the pruning kernels;
//   * transpose_kernel     Bt of a row-major B (the GEMM wants both operands K-contiguous);
//   * bin_matrix_kernel    bin_matrix(): column sums per bin of the bin's middle row(s).
#pragma once
#include "kernels.cuh"

namespace mope {

// P[i][j] = Binomial(j; n_trials, p_i) for i < rows_valid, j <= n_trials (and j < cols); zero elsewhere in the rows x cols
// matrix (leading dimension ld).  log p_i, log(1 - p_i) and ln C(n_trials, j) come from the host (libm / SciPy gammaln);
// pclass marks p_i == 0 (1) and p_i == 1 (2), where the pmf is an indicator.
__global__ void binom_matrix_kernel(double* __restrict__ P, int ld, int rows, int cols, int rows_valid, int n_trials,
                                    const double* __restrict__ logp, const double* __restrict__ log1mp,
                                    const double* __restrict__ lnchoose, const int8_t* __restrict__ pclass) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= ld || i >= rows) return;
    double v = 0.0;
    if (i < rows_valid && j <= n_trials && j < cols) {
        const int cls = pclass[i];
        if (cls == 1) v = (j == 0) ? 1.0 : 0.0;
        else if (cls == 2) v = (j == n_trials) ? 1.0 : 0.0;
        else v = exp(lnchoose[j] + (double)j * logp[i] + (double)(n_trials - j) * log1mp[i]);
    }
    P[(size_t)i * ld + j] = v;
}

__global__ void transpose_kernel(const double* __restrict__ A, double* __restrict__ At, int rows, int cols, int lda, int ldt) {
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int r = r0 + k, c = c0 + threadIdx.x;
        tile[k][threadIdx.x] = (r < rows && c < cols) ? A[(size_t)r * lda + c] : 0.0;
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        const int c = c0 + k, r = r0 + threadIdx.x;
        if (c < cols && r < rows) At[(size_t)c * ldt + r] = tile[threadIdx.x][k];
    }
}

// out[bi][bj] = sum_{j in bin bj} P[mid(bi)][j], mid = the bin's middle state; bins with an even number of states take
// the mean of the two middle rows (generate_transitions.py:194-208).  Sequential left-to-right sums like np.add.reduceat.
__global__ void bin_matrix_kernel(const double* __restrict__ P, int ld, int n_states, const int32_t* __restrict__ breaks, int F,
                                  double* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= F * F) return;
    const int bi = idx / F, bj = idx % F;
    const int j0 = breaks[bj], j1 = (bj + 1 < F) ? breaks[bj + 1] : n_states;
    const int i0 = breaks[bi], i1 = (bi + 1 < F) ? breaks[bi + 1] : n_states;
    const int len = i1 - i0;
    const int lo = i0 + (len - 1) / 2, hi = i0 + len / 2;   // floor / ceil of (len - 1) / 2
    double s_lo = 0.0;
    for (int j = j0; j < j1; ++j) s_lo += P[(size_t)lo * ld + j];
    double v = s_lo;
    if ((len & 1) == 0) {
        double s_hi = 0.0;
        for (int j = j0; j < j1; ++j) s_hi += P[(size_t)hi * ld + j];
        v = (s_lo + s_hi) / 2;
    }
    out[idx] = v;
}

// ------------------------------------------------------------------------------------------------
// C[M x N] = A[M x K] . Bt[N x K]^T, fp64, row-major with leading dimensions >= the padded extents.
//
// CTA tile 64 x (16 NTW): 8 warps as 4 (M) x 2 (N), a warp owns 16 rows x NTW n-tiles of 8 columns (2 x NTW x 2
// accumulator doubles per thread).  K streams in 16-wide chunks (128 B per row) by TMA with the 128-byte swizzle into a
// GEMM_NS-stage full/empty mbarrier ring, GEMM_PF chunks ahead; fragment reads are the conflict-free {2q, 2q+1, 2q+8,
// 2q+9} pattern of the pruning kernels.  Out-of-range rows / columns of the boxes are zero-filled by the tensor maps, so
// M, N, K need no padding in memory beyond an even leading dimension.  NTW = 7 gives 16 x 9 = 144 tiles for the
// 1001 x 1001 Wright-Fisher matrices on 148 SMs (one wave).
// ------------------------------------------------------------------------------------------------
constexpr int GEMM_NS = 6, GEMM_PF = 4;

struct GemmParams {
    CUtensorMap tmapA;   // [M][K], box 16 x 64
    CUtensorMap tmapB;   // [N][K], box 16 x 16 NTW
    double* C;
    int32_t M, N, K, ldc;
    int32_t tiles_n;
};

template <int NTW>
struct GemmCfg {
    static constexpr int A_BYTES = TM * KC * 8;             // 8 KB
    static constexpr int B_BYTES = 2 * NTW * 8 * KC * 8;    // 16 NTW rows x 128 B
    static constexpr int STAGE_BYTES = A_BYTES + ((B_BYTES + 1023) / 1024) * 1024;
    static constexpr size_t SMEM_BYTES = (size_t)GEMM_NS * STAGE_BYTES + 2 * GEMM_NS * sizeof(unsigned long long) + 1024;
};

template <int NTW>
__global__ void __launch_bounds__(NTHREADS, 1) dgemm_tn_kernel(const __grid_constant__ GemmParams p) {
    using Cfg = GemmCfg<NTW>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + (size_t)GEMM_NS * Cfg::STAGE_BYTES);
    unsigned long long* empty = full + GEMM_NS;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp & 3, wn = warp >> 2;
    const int x = lane >> 2, kk = lane & 3;
    const int z = x ^ ((kk >> 1) << 2);
    const int h = (kk & 1) << 3;
    const int a_base = (wm * 16 + x) * 128 + h;
    const int b_base = Cfg::A_BYTES + (wn * NTW * 8 + x) * 128 + h;
    const int tm = blockIdx.x / p.tiles_n, tn = blockIdx.x % p.tiles_n;
    const int row0 = tm * TM, col0 = tn * 16 * NTW;
    const int nchunks = (p.K + KC - 1) / KC;

    if (threadIdx.x == 0) {
        for (int s = 0; s < GEMM_NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NTHREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int c) {
        const unsigned s = (unsigned)c % GEMM_NS;
        mbar_wait(&empty[s], (((unsigned)c / GEMM_NS) & 1u) ^ 1u);
        mbar_expect_tx(&full[s], (unsigned)(Cfg::A_BYTES + Cfg::B_BYTES));
        unsigned char* st = ring + (size_t)s * Cfg::STAGE_BYTES;
        tma_load_2d(st, &p.tmapA, c * KC, row0, &full[s]);
        tma_load_2d(st + Cfg::A_BYTES, &p.tmapB, c * KC, col0, &full[s]);
    };
    if (threadIdx.x == 0)
        for (int c = 0; c < GEMM_PF && c < nchunks; ++c) issue(c);

    double acc[2][NTW][2];
#pragma unroll
    for (int j = 0; j < NTW; ++j) { acc[0][j][0] = acc[0][j][1] = acc[1][j][0] = acc[1][j][1] = 0.0; }

    for (int c = 0; c < nchunks; ++c) {
        if (threadIdx.x == 0 && c + GEMM_PF < nchunks) issue(c + GEMM_PF);
        __syncwarp();
        const unsigned s = (unsigned)c % GEMM_NS;
        mbar_wait(&full[s], ((unsigned)c / GEMM_NS) & 1u);
        const unsigned char* st = ring + (size_t)s * Cfg::STAGE_BYTES;
#pragma unroll
        for (int q = 0; q < KC / 4; ++q) {
            const int oq = ((q ^ z) << 4);
            const unsigned char* pa = st + a_base + oq;
            const unsigned char* pb = st + b_base + oq;
            const double a0 = *reinterpret_cast<const double*>(pa);
            const double a1 = *reinterpret_cast<const double*>(pa + 1024);
#pragma unroll
            for (int j = 0; j < NTW; ++j) {
                const double b = *reinterpret_cast<const double*>(pb + j * 1024);
                dmma(acc[0][j][0], acc[0][j][1], a0, b);
                dmma(acc[1][j][0], acc[1][j][1], a1, b);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    // epilogue: thread (x, kk) holds rows x, x + 8 of the warp's 16 and columns 8j + 2kk, + 1 of every n-tile
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
        const int r = row0 + wm * 16 + x + 8 * mi;
        if (r >= p.M) continue;
        double* crow = p.C + (size_t)r * p.ldc;
#pragma unroll
        for (int j = 0; j < NTW; ++j) {
            const int col = col0 + (wn * NTW + j) * 8 + 2 * kk;
            if (col + 1 < p.N) *reinterpret_cast<double2*>(crow + col) = make_double2(acc[mi][j][0], acc[mi][j][1]);
            else if (col < p.N) crow[col] = acc[mi][j][0];
        }
    }
}

}  // namespace mope
