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
//
// Which rows share a matrix does not depend on the parameters: time = multiplier(row) * length(branch) and
// alpha = coefficient(position(row)) * size(branch), so rows with the same (multiplier value, position) on a branch form
// a GROUP for every parameter vector.  The groups are built once per pedigree (host, mope_b200.cu) and stay on the
// device; per evaluation sel_plan_kernel derives every group's (time, alpha), the range checks of
// TransitionData.get_transition_probabilities_2d and the interpolation job -- nothing is planned on the host.
#pragma once
#include "kernels.cuh"

namespace mope {

constexpr int SEL_ROWS = 4;   // operand rows applied per pass over a matrix (a group usually holds 3: a locus and its two pseudo-rows)

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
// One CTA per matrix.  The kernel is bound by the latency of its global loads, so every warp keeps TWO matrix rows (J
// 32-wide pieces each) and the parent values it will update in flight at once, and only then runs the dot products.
template <int J>
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
        for (int e = threadIdx.x; e < SEL_ROWS * Fp; e += blockDim.x) {
            const int q = e / Fp, f = e - q * Fp;
            vs[e] = (q < nr && f < F) ? V[(size_t)job_rows[rb + q] * Fp + f] : 0.0;   // unused rows / padding: zeros
        }
        __syncthreads();
        // lanes 0..nr-1 update row i of the matrix for the group's rows, lanes 8..8+nr-1 row i + 8
        const int half = lane >> 3, ql = lane & 7;
        const bool upd_lane = half < 2 && ql < nr;
        const size_t row_base = upd_lane ? (size_t)job_rows[rb + ql] * Fp : 0;
        for (int i = warp; i < Fp; i += 16) {
            const int ia = i, ib = i + 8;
            const bool va = ia < Fp, vb = ib < Fp;
            const bool ina = va && ia >= i_lo && ia < i_hi, inb = vb && ib >= i_lo && ib < i_hi;
            double pa[J], pb[J];
#pragma unroll
            for (int jj = 0; jj < J; ++jj) {
                const int k = lane + 32 * jj;
                const bool kin = k >= k_lo && k < k_hi;
                pa[jj] = (ina && off >= 0 && kin) ? pool[off + (size_t)ia * Fp + k] : 0.0;
                pb[jj] = (inb && off >= 0 && kin) ? pool[off + (size_t)ib * Fp + k] : 0.0;
            }
            // the parent values this lane will update, in flight with the matrix rows
            const int my_i = half == 0 ? ia : ib;
            const bool my_valid = upd_lane && (half == 0 ? va : vb);
            double pv = 1.0;
            if (my_valid && parent != nullptr && my_i < F) pv = parent[row_base + my_i];
            double acca[SEL_ROWS], accb[SEL_ROWS];
#pragma unroll
            for (int q = 0; q < SEL_ROWS; ++q) { acca[q] = 0.0; accb[q] = 0.0; }
            if (off < 0) {
                // identity: Y[r][i] = V[r][i] inside the mask
                if (lane == 0) {
#pragma unroll
                    for (int q = 0; q < SEL_ROWS; ++q) {
                        if (ina && ia >= k_lo && ia < k_hi) acca[q] = vs[q * Fp + ia];
                        if (inb && ib >= k_lo && ib < k_hi) accb[q] = vs[q * Fp + ib];
                    }
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < J; ++jj) {
                    const int k = lane + 32 * jj;
                    if (k < Fp) {
#pragma unroll
                        for (int q = 0; q < SEL_ROWS; ++q) {
                            const double v = vs[q * Fp + k];
                            acca[q] = fma(pa[jj], v, acca[q]);
                            accb[q] = fma(pb[jj], v, accb[q]);
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < SEL_ROWS; ++q) {
                if (q < nr) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        acca[q] += __shfl_xor_sync(0xffffffffu, acca[q], o);
                        accb[q] += __shfl_xor_sync(0xffffffffu, accb[q], o);
                    }
                }
            }
            if (my_valid) {
                double y = 0.0;
#pragma unroll
                for (int q = 0; q < SEL_ROWS; ++q) if (q == ql) y = half == 0 ? acca[q] : accb[q];
                if (my_i >= F) y = 0.0;
                const size_t o = row_base + my_i;
                if (parent) parent[o] = (my_i < F) ? pv * y : 0.0;
                if (trans) trans[o] = y;
            }
        }
    }
}

// generic row length: the same in pieces of J = 4 (wide grids)
__global__ void __launch_bounds__(256) sel_matvec_wide_kernel(const double* __restrict__ pool, const int64_t* __restrict__ mat_off,
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
        for (int e = threadIdx.x; e < SEL_ROWS * Fp; e += blockDim.x) {
            const int q = e / Fp, f = e - q * Fp;
            vs[e] = (q < nr && f < F) ? V[(size_t)job_rows[rb + q] * Fp + f] : 0.0;
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
                        for (int q = 0; q < SEL_ROWS; ++q) acc[q] = vs[q * Fp + i];
                    }
                } else {
                    const double* prow = pool + off + (size_t)i * Fp;
                    for (int k = k_lo + lane; k < k_hi; k += 32) {
                        const double p = prow[k];
#pragma unroll
                        for (int q = 0; q < SEL_ROWS; ++q) acc[q] = fma(p, vs[q * Fp + k], acc[q]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < SEL_ROWS; ++q) {
                if (q < nr) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
                }
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

// launcher: the instance whose J pieces cover a row
inline cudaError_t launch_sel_matvec(int n_jobs, cudaStream_t st, const double* pool, const int64_t* mat_off, const int32_t* job_row_off,
                                     const int32_t* job_rows, const double* V, double* parent, double* trans, int F, int Fp, int mask) {
    const size_t smem = (size_t)SEL_ROWS * Fp * sizeof(double);
    if (Fp <= 96) sel_matvec_kernel<3><<<n_jobs, 256, smem, st>>>(pool, mat_off, job_row_off, job_rows, V, parent, trans, F, Fp, mask);
    else if (Fp <= 128) sel_matvec_kernel<4><<<n_jobs, 256, smem, st>>>(pool, mat_off, job_row_off, job_rows, V, parent, trans, F, Fp, mask);
    else if (Fp <= 224) sel_matvec_kernel<7><<<n_jobs, 256, smem, st>>>(pool, mat_off, job_row_off, job_rows, V, parent, trans, F, Fp, mask);
    else sel_matvec_wide_kernel<<<n_jobs, 256, smem, st>>>(pool, mat_off, job_row_off, job_rows, V, parent, trans, F, Fp, mask);
    return cudaGetLastError();
}

// One thread per (branch, group): time = g_mult * branch_length, alpha = locus_alphas[branch][g_pos]; range checks
// (transition_data_2d.py:158-173) OR-ed into *status; the interpolation job of the group (transition_data_2d.py:183-224,
// dst = slot of the group inside the pool chunk it will be processed in) and its matrix offset (-1: identity short-cut).
struct SelPlanParams {
    const double* g_mult;        // [n] multiplier value of the group (1 for branches without a multiplier)
    const int32_t* g_pos;        // [n] unique-position index
    const int32_t* g_branch;     // [n]
    const int32_t* g_local;      // [n] index of the group inside its branch
    const double* branch_len;    // [B]
    const double* alphas;        // [B][n_pos]
    int32_t n, n_pos, pool_blocks;
    const double *gens, *us;
    int32_t n_gens, n_us;
    double N, max_coal, min_mut, max_mut;
    int32_t F, Fp;
    InterpJob* jobs;             // [n]
    int64_t* mat_off;            // [n]
    int32_t* status;
};

__global__ void sel_plan_kernel(SelPlanParams p) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.n) return;
    const int b = p.g_branch[gid];
    const double len = __dmul_rn(p.g_mult[gid], p.branch_len[b]);     // data[multipliername].values * branch_lengths[b]
    const double mut = p.alphas[(size_t)b * p.n_pos + p.g_pos[gid]];
    const size_t FF = (size_t)p.F * p.F;
    int32_t st = 0;
    if (!isfinite(len) || !isfinite(mut)) st |= 32;
    if (len > p.max_coal) st |= 1;
    if (mut < p.min_mut) st |= 2;
    if (mut > p.max_mut) st |= 4;
    InterpJob jb;
    jb.pad = 0;
    jb.flags = 8;
    jb.dst = (int64_t)((size_t)(p.g_local[gid] % p.pool_blocks) * blk_doubles(p.Fp));
    jb.src[0] = jb.src[1] = jb.src[2] = jb.src[3] = 0;
    jb.w[0] = jb.w[1] = jb.w[2] = jb.w[3] = 0.0;
    int64_t moff = -1;
    const double t = __dmul_rn(p.N, len);
    if (!st && !(t < p.gens[0])) {
        const int G = p.n_gens, U = p.n_us;
        const int gi = min(lower_bound_dev(p.gens, G, t), G - 1);
        const double xd = mut / (2.0 * p.N);
        const int xi = min(lower_bound_dev(p.us, U, xd), U - 1);
        const int g1 = gi > 0 ? gi - 1 : G - 1, x1i = xi > 0 ? xi - 1 : U - 1;   // python's index -1 wraps
        const double t1 = p.gens[g1], t2 = p.gens[gi], x1 = p.us[x1i], x2 = p.us[xi];
        const double denom = __dmul_rn(__dsub_rn(t2, t1), __dsub_rn(x2, x1));
        jb.w[0] = __ddiv_rn(__dmul_rn(__dsub_rn(t2, t), __dsub_rn(x2, xd)), denom);
        jb.w[2] = __ddiv_rn(__dmul_rn(__dsub_rn(t, t1), __dsub_rn(x2, xd)), denom);
        jb.w[1] = __ddiv_rn(__dmul_rn(__dsub_rn(t2, t), __dsub_rn(xd, x1)), denom);
        jb.w[3] = __ddiv_rn(__dmul_rn(__dsub_rn(t, t1), __dsub_rn(xd, x1)), denom);
        jb.src[0] = (int64_t)(((size_t)g1 * U + x1i) * FF);
        jb.src[1] = (int64_t)(((size_t)g1 * U + xi) * FF);
        jb.src[2] = (int64_t)(((size_t)gi * U + x1i) * FF);
        jb.src[3] = (int64_t)(((size_t)gi * U + xi) * FF);
        jb.flags = 1;
        moff = jb.dst;
    }
    p.jobs[gid] = jb;
    p.mat_off[gid] = moff;
    if (st) atomicOr(p.status, st);
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
