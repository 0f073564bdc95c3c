// Selection model and masked-matrix variants.
//
// In the reference's selection model every base-pair position has its own scaled selection coefficient, so the
// transition matrix of a branch is fetched per ROW, P_i = transitions(time_i, alpha_i), and applied as a mat-vec
// (mope/_likes.pyx:59-79 rate version, :155-175 fixed-length version); the posterior-mutation-location code applies
// masked copies of P (`Pzero`, mope/_likes.pyx:82-136,178-229).  Rows that share (time, alpha) -- a position's three rows
// (likelihood, all-zero and all-one ascertainment rows), repeated across families for fixed-length branches -- share P.
// Here every DISTINCT (time, alpha) pair of a branch is interpolated once by interp_kernel (the checked, renormalised
// matrix of the main path, natural column order) and sel_matvec_kernel applies it to all of its rows: one CTA per
// matrix, the operand rows staged in shared memory, one warp per matrix row with the row read coalesced exactly once.
// The work is bound by moving the interpolated matrices (one F x F block per distinct pair), not by arithmetic.
#pragma once
#include "kernels.cuh"

namespace mope {

constexpr int SEL_ROWS = 8;   // operand rows applied per pass over a matrix

// mask: 0 none; 1 "zero" Pzero[0,0] = P[0,0]; 2 "focal node" Pzero[0,1:-1] = P[0,1:-1]; 3 "descendant"
// Pzero[1:-1,1:-1] = P[1:-1,1:-1] (everything else zero).
__device__ __forceinline__ void sel_mask_ranges(int mask, int F, int& i_lo, int& i_hi, int& k_lo, int& k_hi) {
    i_lo = 0; i_hi = F; k_lo = 0; k_hi = F;
    if (mask == 1) { i_hi = 1; k_hi = 1; }
    else if (mask == 2) { i_hi = 1; k_lo = 1; k_hi = F - 1; }
    else if (mask == 3) { i_lo = 1; i_hi = F - 1; k_lo = 1; k_hi = F - 1; }
}

// Y[r][:] = Pmask_j . V[r][:] for the rows r of job j; parent[r][:] *= Y[r][:]; trans[r][:] = Y[r][:] (optional).
// mat_off[j] < 0: identity matrix (N*t below the first table generation, transition_data_2d.py:184-187).
__global__ void __launch_bounds__(256) sel_matvec_kernel(const double* __restrict__ pool, const int64_t* __restrict__ mat_off,
                                                         const int32_t* __restrict__ job_row_off, const int32_t* __restrict__ job_rows,
                                                         const double* __restrict__ V, double* __restrict__ parent, double* __restrict__ trans,
                                                         int F, int Fp, int mask) {
    extern __shared__ double vs[];   // [SEL_ROWS][Fp]
    const int j = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = job_row_off[j], r1 = job_row_off[j + 1];
    const int64_t off = mat_off[j];
    int i_lo, i_hi, k_lo, k_hi;
    sel_mask_ranges(mask, F, i_lo, i_hi, k_lo, k_hi);
    for (int rb = r0; rb < r1; rb += SEL_ROWS) {
        const int nr = min(SEL_ROWS, r1 - rb);
        __syncthreads();
        for (int e = threadIdx.x; e < nr * Fp; e += blockDim.x) {
            const int q = e / Fp, f = e - q * Fp;
            vs[e] = (f < F) ? V[(size_t)job_rows[rb + q] * Fp + f] : 0.0;
        }
        __syncthreads();
        for (int i = warp; i < Fp; i += 8) {
            double acc[SEL_ROWS];
#pragma unroll
            for (int q = 0; q < SEL_ROWS; ++q) acc[q] = 0.0;
            if (i >= i_lo && i < i_hi) {
                if (off < 0) {
                    if (lane == 0 && i >= k_lo && i < k_hi) {
#pragma unroll
                        for (int q = 0; q < SEL_ROWS; ++q) if (q < nr) acc[q] = vs[q * Fp + i];
                    }
                } else {
                    const double* prow = pool + off + (size_t)i * Fp;
                    for (int k = k_lo + lane; k < k_hi; k += 32) {
                        const double p = prow[k];
#pragma unroll
                        for (int q = 0; q < SEL_ROWS; ++q) if (q < nr) acc[q] = fma(p, vs[q * Fp + k], acc[q]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < SEL_ROWS; ++q) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
            }
            if (lane < nr) {
                double y = 0.0;
#pragma unroll
                for (int q = 0; q < SEL_ROWS; ++q) if (q == lane) y = acc[q];
                if (i >= F) y = 0.0;
                const size_t o = (size_t)job_rows[rb + lane] * Fp + i;
                if (parent) parent[o] = (i < F) ? parent[o] * y : 0.0;
                if (trans) trans[o] = y;
            }
        }
    }
}

// node_likes of a leaf in the selection model (likelihoods.py:427-433,465-469): rows [0, L) the leaf's emissions, rows
// [L, 2L) e0 = 1[freq <= min_freq], rows [2L, 3L) eN = 1[freq >= 1 - min_freq]; natural column order, Fp-padded.
__global__ void sel_leaf_kernel(const double* __restrict__ E_leaf /* [R_pad][Fp], k-permuted when perm */, int perm, int L, int F, int Fp,
                                const int8_t* __restrict__ e0, const int8_t* __restrict__ eN, double* __restrict__ out /* [3L][Fp] */) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)3 * L * Fp) return;
    const int r = (int)(idx / Fp), f = (int)(idx % Fp);
    double v = 0.0;
    if (f < F) {
        if (r < L) v = E_leaf[(size_t)r * Fp + (perm ? kperm(f) : f)];
        else if (r < 2 * L) v = e0[f] ? 1.0 : 0.0;
        else v = eN[f] ? 1.0 : 0.0;
    }
    out[idx] = v;
}

__global__ void sel_fill_kernel(double* __restrict__ dst, long long n, double v) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) dst[idx] = v;
}

// asc[i][f] *= 1 - trans[L + i][f] - trans[2L + i][f] for a branch below the root (likelihoods.py:487-493,506-510)
__global__ void sel_asc_kernel(const double* __restrict__ trans, double* __restrict__ asc, int L, int Fp) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)L * Fp) return;
    const int i = (int)(idx / Fp), f = (int)(idx % Fp);
    const double zero = trans[(size_t)(L + i) * Fp + f], one = trans[(size_t)(2 * L + i) * Fp + f];
    asc[idx] *= (1.0 - zero) - one;
}

// loglikes[i] = log(root[i] . stat), log_asc[i] = log(asc[i] . stat) (_likes.compute_root_locus_log_likes; likelihoods.py:517)
__global__ void sel_root_kernel(const double* __restrict__ root, const double* __restrict__ asc, const double* __restrict__ stat, int L, int F,
                                int Fp, double* __restrict__ loglikes, double* __restrict__ log_asc) {
    const int lane = threadIdx.x & 31;
    const int i = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i >= L) return;
    double a = 0.0, b = 0.0;
    for (int f = lane; f < F; f += 32) {
        const double s = stat[f];
        a = fma(root[(size_t)i * Fp + f], s, a);
        b = fma(asc[(size_t)i * Fp + f], s, b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) { loglikes[i] = log(a); log_asc[i] = log(b); }
}

}  // namespace mope
