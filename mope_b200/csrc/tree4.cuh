// tree_kernel4 -- fused pruning pass, "warp-owned rows" variant (used when F <= 208).
//
// Same arithmetic as tree_kernel (kernels.cuh) with a different decomposition: each of the 8 warps of a CTA owns
// 8 complete rows of the 64-row tile and all F output bins, so a message never has to cross warps:
//   * the output of a branch GEMM sits in the warp's mma C fragments (thread (x, kk) holds bins 8n+2kk, 8n+2kk+1 of
//     row x).  With the k index of an mma step taken as {8j + 2kk + h : kk = 0..3} those very registers are the A
//     fragment of the NEXT branch (the parent's branch follows its last child in postorder), so chained messages
//     go through warp-private TENSOR MEMORY columns (tcgen05.st / tcgen05.ld, thread i <-> TMEM lane i) -- no global
//     round trip, no proxy fence, no CTA-wide synchronisation;
//   * only the message of a non-last child is parked (fragment layout) and multiplied in when the last sibling
//     arrives: the busiest slots of the schedule live in tensor memory too (as many as fit the warp's 256 columns),
//     the others in an L2-resident scratch (coalesced 16-byte accesses by the thread that later reads them back);
//   * the only shared resource is the TMA ring with the P chunks (and emission chunks for leaf branches); since no
//     copy depends on an epilogue, lane 0 of warp k streams chunks k, k+8, ... of the item PF chunks ahead without
//     ever stalling on a dependency;
//   * items (walker, 64-row tile) are drawn from an atomic counter; tiles that hold only ascertainment pseudo-rows
//     take the messages of leaf branches and of family-independent subtrees from precomputed vectors (ASCFAST).
// K layout: the transition matrices and the emission tensor are stored with their k (column) index permuted inside
// every 16-wide chunk, kperm() below, so that the chained k sets land on the bank-conflict-free fragment positions
// {2q, 2q+1, 2q+8, 2q+9} of the 128-byte swizzled chunk.
#pragma once
#include "kernels.cuh"

namespace mope {

#ifdef MOPE_PROF   // per-warp cycle accounting for kernel tuning (printed by two CTAs at the end of the kernel)
#define PROF_DECL long long pf_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pf_c0 = clock64(), pf_start = pf_c0;
#define PROF_MARK(i) { const long long pf_c1 = clock64(); pf_t[i] += pf_c1 - pf_c0; pf_c0 = pf_c1; }
#else
#define PROF_DECL
#define PROF_MARK(i)
#endif

constexpr int NS4_MAX = 6;   // most ring stages of any configuration (Cfg4<NTT>::NS)
// Homes of parked messages (Op4::home), chosen per park when the schedule is compiled (mope_b200.cu, compile_schedule4):
//   * HOME_CHAIN: the chain buffer itself.  It is only live from the post-op of a node's last child to the GEMM of the
//     node's own branch, so a park whose whole lifetime is spanned by leaf branches (the first leaf of a cherry, the usual
//     case) can sit there for free;
//   * k >= 0: one of the warp's whole tensor-memory slots, handed out by interval scheduling (earliest end first), which
//     parks as many messages as possible there;
//   * <= HOME_L2: the L2-resident scratch, for the parks that are left.
constexpr int HOME_CHAIN = -1;
constexpr int HOME_L2 = -2;

struct Op4 {
    int32_t leaf;         // >= 0: leaf column, A operand = emission rows (TMA ring); -1: A operand chained from the previous op
    int32_t kind;         // 0 shared matrix (fixed length or bottleneck), 1 per-row times (rate)
    int32_t key;          // matrix key / rate key
    int32_t mult;         // rate: multiplier column
    int32_t mul_slot;     // >= 0: this op is the LAST child of its parent: multiply the parked sibling product in
    int32_t store_slot;   // >= 0: not the last child: park the message (store_first) or fold it into the parked one
    int32_t store_first;
    int32_t home;         // where the parked product of mul_slot / store_slot lives (HOME_*): k >= 0 the k-th tensor-memory slot
                          // of the warp, HOME_CHAIN the chain buffer, <= HOME_L2 slot (HOME_L2 - home) of the L2 scratch
    int32_t last;         // completes the root: root reduction
    int32_t uni;          // ascertainment pseudo-rows: 0 the message depends on the family (an age-scaled branch at or below it),
                          // 1 inside a subtree whose messages are the same for every family, 2 the top branch of such a subtree
    int32_t uroot;        // uni == 2 and not a leaf: index of the per-walker message in Tree4Params::uni_msg
};

struct Tree4Params {
    CUtensorMap tmapE;      // emissions [n_leaves*R_pad][Fp] (k-permuted columns), box 16 x 64
    CUtensorMap tmapP;      // matrix pool [n_blocks*(Fp+BLK_EXTRA)][Fp] (k-permuted columns), box 16 x 8*NTT
    const double* E;
    int64_t E_leaf_stride;
    double* scratch;        // [gridDim.x][8 warps][n_slots + 1][NTT][32][2]: parked messages whose home is the L2 scratch
    const Op4* ops;
    int32_t n_ops, n_slots;
    int32_t R, L, R_pad, Fp, F, NT, n_tiles, W, tile_block;
    const double* pool;
    const int64_t* mat_off;  // [W][n_mat], negative = identity
    int32_t n_mat, n_rate;
    const RateInfo* rate;
    const double* mult;
    const double* gens;
    int32_t n_gens;
    double Npop;
    const double* stat;      // [W][Fp] natural order
    double* tile_ll;
    double* asc_dot;
    double* Yout;
    int32_t debug;           // unused by tree_kernel4 (ablations are compile-time: -DMOPE_ABL); kept for the wide-grid kernel's launcher
    // Messages of ascertainment pseudo-rows that do not depend on the family (no age-scaled branch at or below the branch)
    // are the same for all even / all odd pseudo-rows of a walker.  A first, one-tile-per-walker launch (canonical = 1)
    // writes them to uni_msg [W][n_uroot][2][Fp] (natural column order); in the main launch a pure pseudo-row tile skips
    // the branches inside such subtrees and reads the message of their top branch from there.
    double* uni_msg;
    int32_t n_uroot, canonical;
    int32_t* work_ctr;       // zeroed before the launch: next item to hand out (items differ in cost: identity branches)
    unsigned long long* unit_rows;   // statistics (may be null): [0] += real rows x k columns actually run through the tensor pipe;
                                     // [1] / [2] += (warp, generation segment) pairs of age-scaled branches sat through / multiplied in
    const unsigned long long* emask;   // [n_leaves][R_pad / 8]: K chunks of an 8-row group of E that are not all zero (may be null)
    const int32_t* wstatus;  // device-planned evaluations: per-walker MOPE_ST_* bits, non-zero = skip the walker (may be null)
};

struct Tree4Smem {
    unsigned long long full[NS4_MAX], empty[NS4_MAX];
    Op4 ops[MAX_OPS];
    int32_t prow[MAX_OPS];     // per item: pool row of a shared-matrix op (negative = identity)
    int32_t seg_lo[MAX_OPS], seg_hi[MAX_OPS];   // per item: generation-grid range of a rate op over the tile's rows
    int32_t unit_first[MAX_OPS + 1];            // per item: first GEMM unit (op x generation segment) of every op
    int32_t unit_prow0[MAX_OPS];                // per item: pool row of the op's first unit (later segments: + k*(Fp+1))
    double rowval[TM];         // root reduction: per-row log-likelihood terms
    uint32_t tmem_base;        // tcgen05.alloc result
    long long cur_item;        // the item the CTA works on (handed out dynamically)
};

template <int NTT>
struct Cfg4 {
    static constexpr int PROWS = NTT * 8;
    static constexpr int E_BYTES = TM * KC * 8;          // 8 KB
    static constexpr int P_BYTES = PROWS * KC * 8;
    static constexpr int STAGE_BYTES = E_BYTES + P_BYTES;  // multiples of 1024
    // ring depth: a prefetch distance of 2 chunks with 2 stages of slack is as fast as deeper rings (measured), and the
    // shared memory it does not take stays L1 (the wide configuration gains 0.6 % from the larger L1)
    static constexpr int NS = (NTT > 16) ? 4 : 6;
    static constexpr int PF = NS - 2;   // NS - PF >= 2 keeps the issuing lane from waiting on the slowest warp
    // tensor-memory budget of a warp (256 columns): chain buffer + TM_SLOTS whole parked slots (F = 201: 1, F = 115: 3,
    // F = 65: 5).  Splitting one more slot between the 48 spare columns of the wide configuration and shared memory or
    // the L2 scratch was measured slower than leaving it in the scratch.
    static constexpr int MSG_COLS = NTT * 4;
    static constexpr int TM_SLOTS = 256 / MSG_COLS - 1;
    static constexpr size_t SMEM_BYTES = (size_t)NS * STAGE_BYTES + sizeof(Tree4Smem) + 1024;
    static_assert(SMEM_BYTES <= 227 * 1024, "ring + tables exceed the 227 KB of shared memory a CTA can have");
};

// per-row time weights of an age-scaled branch (transition_data_2d.py:183-224)
__device__ __forceinline__ void rate_row(const Tree4Params& p, const Op4& op, const RateInfo* ri, int r, int& gi, double& a, double& d) {
    gi = -1; a = 0.0; d = 0.0;
    if (r >= p.R) return;
    const double t = __dmul_rn(p.Npop, __dmul_rn(p.mult[(size_t)op.mult * p.R_pad + r], ri->len));
    if (t < p.gens[0]) return;
    int lo = 0, hi = p.n_gens;  // searchsorted left
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (p.gens[mid] < t) lo = mid + 1; else hi = mid; }
    gi = lo < p.n_gens ? lo : p.n_gens - 1;
    const double t2 = p.gens[gi];
    const double t1 = p.gens[gi > 0 ? gi - 1 : p.n_gens - 1];
    const double den = __dsub_rn(t2, t1);
    a = __ddiv_rn(__dsub_rn(t2, t), den);
    d = __ddiv_rn(__dsub_rn(t, t1), den);
    if (gi == 0) a = 0.0;  // wrapped neighbour carries an exactly-zero weight
}

// y / den for the row renormalisation of an age-scaled branch.  den is a blend of stored row sums, 1 + e with e of the
// order of the tables' rounding excess (<= 6e-9 measured on the default grid up to 10 000 generations), so
// 1 / den = 1 - e + e^2 - ... and y (2 - den) = y (1 - e) is the quotient to a relative e^2 < 1e-16, i.e. to the last
// bit or so -- two instructions instead of a division per element (104 per thread and branch).  Larger excesses divide.
__device__ __forceinline__ double renorm_div(double y, double den) {
    return (den < 1.0 + 1e-8) ? y * (2.0 - den) : y / den;
}

// ---- tensor memory as warp-private scratch ----------------------------------------------------------------------
// A message in fragment layout is thread-private (the thread that writes an element is the one that reads it back), so
// it can live in TMEM: with the 32x32b shape thread i of warp w owns TMEM lane 32*(w%4)+i, and every warp gets 256 of
// the 512 columns of its lane quarter.  One 16-wide chunk of a row (4 doubles per thread) is 8 columns.  Measured on
// B200 (profiles/microbench/tmem_scratch.cu): tcgen05.st 317 B/clk/SM, tcgen05.ld 340 B/clk/SM, against 32 B/clk/SM
// for the same bursts stored to global memory.
__device__ __forceinline__ void tmem_st4(uint32_t taddr, double v0, double v1, double v2, double v3) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
                 :: "r"(taddr),
                    "r"(__double2loint(v0)), "r"(__double2hiint(v0)), "r"(__double2loint(v1)), "r"(__double2hiint(v1)),
                    "r"(__double2loint(v2)), "r"(__double2hiint(v2)), "r"(__double2loint(v3)), "r"(__double2hiint(v3))
                 : "memory");
}
// asynchronous: the registers are valid after tmem_wait_ld()
__device__ __forceinline__ void tmem_ld4_issue(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ double u2d(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }
constexpr int TMEM_COLS = 512;        // whole tensor memory of the SM (one CTA per SM)
constexpr int TMEM_WARP_COLS = 256;   // per warp: chain buffer at column 0, parked slots after it (Cfg4)

// One 16-wide K chunk of a branch GEMM for the warp's 8 rows: 4 k steps x NTT n-tiles.  A comes from the emission chunk
// in the stage (LEAF) or from the chained fragment values `ac`; RATE scales it by the row's segment weight.
// Only the first NFULL n-tiles go through the tensor pipe.  With SCALAR the single real column of tile NFULL (F = 8 NFULL
// + 1, e.g. F = 201) is accumulated by one DFMA per k step into `sacc` (partial sum over the thread's own k indices; the
// four kk lanes of a row are added up after the GEMM) instead of spending a whole DMMA on seven padding columns.
template <int NTT, int NFULL, bool SCALAR, bool LEAF, bool RATE, bool NO_LDS>
__device__ __forceinline__ void chunk_mma(double (&acc)[NTT][2], double& sacc, const unsigned char* st, int a_base, int b_base, int z,
                                          int bs_base, int z0, const double (&ac)[4], double wgt, int nq) {
    double a[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (LEAF) a[q] = *reinterpret_cast<const double*>(st + a_base + ((q ^ z) << 4));
        else a[q] = ac[q];
        if (RATE) a[q] *= wgt;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (q >= nq) break;   // all-padding k step (zero columns): nothing to add
        const unsigned char* pb = st + b_base + ((q ^ z) << 4);
        double ps = 0.0;
        if (SCALAR) ps = *reinterpret_cast<const double*>(st + bs_base + ((q ^ z0) << 4));   // P[8 NFULL][k], same for all 8 rows
#pragma unroll
        for (int n = 0; n < NFULL; ++n) {
            const double b = NO_LDS ? wgt : *reinterpret_cast<const double*>(pb + n * 1024);
            dmma(acc[n][0], acc[n][1], a[q], b);
        }
        if (SCALAR) sacc = fma(a[q], ps, sacc);
    }
}

template <int NTT, int NFULL, bool SCALAR, bool ASCFAST>
__global__ void __launch_bounds__(NTHREADS, 1) tree_kernel4(const __grid_constant__ Tree4Params p) {
    using Cfg = Cfg4<NTT>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int NS4 = Cfg::NS, PF4 = Cfg::PF;
    Tree4Smem* ts = reinterpret_cast<Tree4Smem*>(ring + (size_t)NS4 * Cfg::STAGE_BYTES);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = lane >> 2, kk = lane & 3;
    const int Fp = p.Fp;
    const int nchunks = Fp / KC;
    const int blk_rows = Fp + BLK_EXTRA;
    const size_t blk = blk_doubles(Fp);
    double* my_scratch = p.scratch + ((size_t)blockIdx.x * 8 + warp) * (p.n_slots + 1) * NTT * 64;
    // fragment addressing inside a stage (128 B swizzle, see kernels.cuh)
    const int z = x ^ ((kk >> 1) << 2);
    const int hb = (kk & 1) << 3;
    const int a_base = (warp * 8 + x) * 128 + hb;
    const int b_base = Cfg::E_BYTES + x * 128 + hb;
    const int z0 = (kk >> 1) << 2;                               // row 8 NFULL of the chunk (the scalar column's matrix row)
    const int bs_base = Cfg::E_BYTES + hb + NFULL * 1024;
    static_assert(NFULL + (SCALAR ? 1 : 0) <= NTT, "n-tiles");

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS4; ++s) { mbar_init(&ts->full[s], 1); mbar_init(&ts->empty[s], NTHREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.n_ops; i += NTHREADS) ts->ops[i] = p.ops[i];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&ts->tmem_base)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    // Chained operand (the message of a node's last child, operand of the node's own branch = the next op) and the parked
    // slots with a tensor-memory home: warp-private tensor-memory columns
    const uint32_t t_chain = ts->tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * TMEM_WARP_COLS);
    static_assert(Cfg::TM_SLOTS >= 1 && NTT % 2 == 0, "chain buffer + one slot must fit the warp's TMEM columns");
    // acc *= parked, or parked = acc, chunk by chunk (a chunk = two n-tiles = 4 doubles per thread), wherever the park lives
    auto slot_mul = [&](double (&acc)[NTT][2], int home) {
        if (home > HOME_L2) {
            const uint32_t tb = t_chain + (uint32_t)((home + 1) * Cfg::MSG_COLS);
#pragma unroll
            for (int c = 0; c < NTT / 2; ++c) {
                uint32_t r[8];
                tmem_ld4_issue(tb + 8 * c, r);
                tmem_wait_ld();
                acc[2 * c][0] *= u2d(r[0], r[1]); acc[2 * c][1] *= u2d(r[2], r[3]);
                acc[2 * c + 1][0] *= u2d(r[4], r[5]); acc[2 * c + 1][1] *= u2d(r[6], r[7]);
            }
        } else {
            const double2* sl = reinterpret_cast<const double2*>(my_scratch + (size_t)(HOME_L2 - home) * NTT * 64) + lane;
#pragma unroll
            for (int n = 0; n < NTT; ++n) { const double2 v = sl[n * 32]; acc[n][0] *= v.x; acc[n][1] *= v.y; }
        }
    };
    auto slot_store = [&](const double (&acc)[NTT][2], int home) {
        if (home > HOME_L2) {
            const uint32_t tb = t_chain + (uint32_t)((home + 1) * Cfg::MSG_COLS);
#pragma unroll
            for (int c = 0; c < NTT / 2; ++c)
                tmem_st4(tb + 8 * c, acc[2 * c][0], acc[2 * c][1], acc[2 * c + 1][0], acc[2 * c + 1][1]);
            tmem_wait_st();
        } else {
            double2* sl = reinterpret_cast<double2*>(my_scratch + (size_t)(HOME_L2 - home) * NTT * 64) + lane;
#pragma unroll
            for (int n = 0; n < NTT; ++n) sl[n * 32] = make_double2(acc[n][0], acc[n][1]);
        }
    };
    // a consumed park of the L2 scratch is dead: drop its lines from L2 instead of letting them be written back to HBM
    auto slot_discard = [&](int home) {
        if (home > HOME_L2) return;
        const double* slot_base = my_scratch + (size_t)(HOME_L2 - home) * NTT * 64;
        __syncwarp();
        for (int line = lane; line < (NTT * 64 * 8) / 128; line += 32)
            asm volatile("discard.global.L2 [%0], 128;\n" ::"l"(slot_base + line * 16) : "memory");
    };

    // k steps of the last chunk that hold at least one real column (natural k = 16c + 8(q>>1) + (q&1) + 2kk)
    int nq_last = 0;
    for (int q = 0; q < 4; ++q) nq_last += (16 * (nchunks - 1) + 8 * (q >> 1) + (q & 1) < p.F) ? 1 : 0;
#ifdef MOPE_ABL   // compile-time ablations for kernel tuning (results are garbage): 1 skip the MMAs, 2 skip the copies and
                  // ring barriers, 8 skip the shared-memory B loads
    constexpr bool no_mma = MOPE_ABL & 1, no_copy = MOPE_ABL & 2;
#else
    constexpr bool no_mma = false, no_copy = false;
#endif
#ifdef MOPE_ABL
    constexpr bool no_lds = MOPE_ABL & 8;
#else
    constexpr bool no_lds = false;
#endif

    unsigned g = 0;   // chunks consumed so far (ring position)
    PROF_DECL
    const long long n_items = p.canonical ? (long long)p.W : (long long)p.W * p.n_tiles;

    for (;;) {
        // items are handed out dynamically in the L2-friendly order: a CTA that drew cheap items (walkers with many
        // identity branches) simply draws more of them
        if (threadIdx.x == 0) ts->cur_item = (long long)atomicAdd(p.work_ctr, 1);
        __syncthreads();
        const long long item = ts->cur_item;
        if (item >= n_items) break;
        int w, tile;
        {
            const long long per_block = (long long)p.W * p.tile_block;
            const int tb = (int)(item / per_block);
            const int t_first = tb * p.tile_block;
            const int t_count = min(p.tile_block, p.n_tiles - t_first);
            const long long rem = item - (long long)tb * per_block;
            w = (int)(rem / t_count);
            tile = t_first + (int)(rem % t_count);
        }
        if (p.canonical) { w = (int)item; tile = 0; }
        if (p.wstatus != nullptr && p.wstatus[w] != 0) { __syncthreads(); continue; }   // walker outside the tables / prior: NaN from finalize
        // the canonical item of a walker: the first pseudo-row pair (rows L, L + 1 = parities 0, 1) leads the tile
        const int row0 = p.canonical ? p.L : tile * TM;
        const int my_row = row0 + warp * 8 + x;     // the row this thread's fragments belong to
        // A tile of ascertainment pseudo-rows only: every leaf starts from e0 (even rows) or eN (odd rows), so the message of
        // a leaf branch is P . e0 / P . eN -- the low / high sums stored with every matrix block -- and needs no GEMM.
        const bool asc_tile = ASCFAST && row0 >= p.L && p.Yout == nullptr;
        const bool asc_cse = asc_tile && !p.canonical && p.uni_msg != nullptr;   // family-independent subtrees come from uni_msg

        // ---- per-item tables: matrix rows, generation ranges of the rate ops ----
        for (int i = threadIdx.x; i < p.n_ops; i += NTHREADS) {
            const Op4& o = ts->ops[i];
            int pr = 0;
            if (o.kind == 0) {
                const long long off = p.mat_off[(size_t)w * p.n_mat + o.key];
                pr = off < 0 ? -1 : (int)(off / Fp);
            }
            ts->prow[i] = pr;
        }
        for (int i = warp; i < p.n_ops; i += 8) {
            const Op4& o = ts->ops[i];
            if (o.kind != 1) continue;
            const RateInfo* ri = &p.rate[(size_t)w * p.n_rate + o.key];
            int lo = 0x7fffffff, hi = -1;
            for (int hh = 0; hh < 2; ++hh) {
                int gi; double a, d;
                rate_row(p, o, ri, row0 + lane + 32 * hh, gi, a, d);
                if (gi >= 0) { lo = min(lo, gi > 0 ? gi - 1 : 0); hi = max(hi, gi); }
            }
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) {
                lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o2));
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o2));
            }
            if (lane == 0) { ts->seg_lo[i] = lo; ts->seg_hi[i] = hi; }
        }
        __syncthreads();

        // ---- flat list of the item's GEMM units (op x segment) so that ANY warp can issue chunk number k of the item:
        //      the copies are issued round-robin by lane 0 of warp (k mod 8), which spreads the issue cost evenly ----
        if (threadIdx.x == 0) {
            int n = 0;
            for (int i = 0; i < p.n_ops; ++i) {
                const Op4& o = ts->ops[i];
                ts->unit_first[i] = n;
                if (asc_tile && o.leaf >= 0) continue;   // no GEMM: message from the low / high sums
                if (asc_cse && o.uni != 0) continue;     // no GEMM: inside / on top of a family-independent subtree
                n += (o.kind == 0) ? (ts->prow[i] >= 0 ? 1 : 0) : max(0, ts->seg_hi[i] - ts->seg_lo[i] + 1);
            }
            ts->unit_first[p.n_ops] = n;
        }
        __syncthreads();
        const int n_units = ts->unit_first[p.n_ops];
        for (int i = threadIdx.x; i < p.n_ops; i += NTHREADS) {
            const Op4& o = ts->ops[i];
            int r0 = ts->prow[i];
            if (o.kind == 1) {
                r0 = 0;
                if (ts->seg_hi[i] >= ts->seg_lo[i]) {
                    const RateInfo* ri = &p.rate[(size_t)w * p.n_rate + o.key];
                    r0 = (int)(ri->off / Fp) + (ts->seg_lo[i] - ri->gbase) * blk_rows;
                }
            }
            ts->unit_prow0[i] = r0;
        }
        __syncthreads();
        const int total_chunks = n_units * nchunks;
        const unsigned g_item0 = g;
        // Copies are issued round-robin: lane 0 of warp k issues chunks k, k+8, ... of the item.  It tracks the position of
        // its next chunk (op, unit inside the op, chunk inside the unit) incrementally; only the first one is searched for.
        int iss_rel = warp, iss_op = 0, iss_uo = 0, iss_ch = 0;
        if (lane == 0 && iss_rel < total_chunks) {
            const int u = iss_rel / nchunks;
            iss_ch = iss_rel - u * nchunks;
            int lo = 0, hi = p.n_ops - 1;   // last op with unit_first[op] <= u (ops without units have empty ranges)
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (ts->unit_first[mid] <= u) lo = mid; else hi = mid - 1;
            }
            iss_op = lo;
            iss_uo = u - ts->unit_first[lo];
        }
        auto issue_next = [&]() {   // called by lane 0: issue chunk iss_rel, then step 8 chunks ahead
            const Op4& o = ts->ops[iss_op];
            const int prow_u = ts->unit_prow0[iss_op] + iss_uo * blk_rows;
            const int erow = o.leaf >= 0 ? o.leaf * p.R_pad + row0 : -1;
            const unsigned gg = g_item0 + (unsigned)iss_rel;
            const unsigned s = gg % NS4;
            PROF_MARK(2)
            mbar_wait(&ts->empty[s], ((gg / NS4) & 1u) ^ 1u);
            PROF_MARK(6)   // waiting for the stage to be released
            unsigned char* st = ring + (size_t)s * Cfg::STAGE_BYTES;
            mbar_expect_tx(&ts->full[s], (unsigned)(Cfg::P_BYTES + (erow >= 0 ? Cfg::E_BYTES : 0)));
            tma_load_2d(st + Cfg::E_BYTES, &p.tmapP, iss_ch * KC, prow_u, &ts->full[s]);
            if (erow >= 0) tma_load_2d(st, &p.tmapE, iss_ch * KC, erow, &ts->full[s]);
            PROF_MARK(7)   // expect_tx + TMA issue
            iss_rel += 8;
            iss_ch += 8;
            while (iss_ch >= nchunks && iss_op < p.n_ops) {
                iss_ch -= nchunks;
                ++iss_uo;
                while (iss_op < p.n_ops && iss_uo >= ts->unit_first[iss_op + 1] - ts->unit_first[iss_op]) { ++iss_op; iss_uo = 0; }
            }
        };
        if (!no_copy && lane == 0 && warp < PF4 && warp < total_chunks) issue_next();
        __syncwarp();
        int crel = 0;   // chunks of this item consumed so far
        int my_kcols = 0;   // k columns (over all GEMM units = branch x generation segment) this warp ran through the tensor pipe
        PROF_MARK(0)   // item setup

        for (int oi = 0; oi < p.n_ops; ++oi) {
            const Op4 op = ts->ops[oi];
            if (asc_cse && op.uni == 1) continue;   // inside a family-independent subtree: its top branch supplies the message
            const bool identity = (op.kind == 0 && ts->prow[oi] < 0);
            const RateInfo* ri = (op.kind == 1) ? &p.rate[(size_t)w * p.n_rate + op.key] : nullptr;
            int gi = 0;
            double ra = 0.0, rd = 0.0;
            if (op.kind == 1) rate_row(p, op, ri, my_row, gi, ra, rd);

            double acc[NTT][2];
#pragma unroll
            for (int n = 0; n < NTT; ++n) { acc[n][0] = 0.0; acc[n][1] = 0.0; }
            double sacc = 0.0;

            // ---------------- GEMM(s) ----------------
            const bool asc_leaf = asc_tile && op.leaf >= 0;
            const bool asc_uroot = asc_cse && op.uni == 2 && op.leaf < 0;
            const int nseg = (identity || asc_leaf || asc_uroot) ? 0 : (op.kind == 0 ? 1 : max(0, ts->seg_hi[oi] - ts->seg_lo[oi] + 1));
            // leaf branches: K chunks in which the warp's 8 emission rows are all zero add nothing (emission_mask_kernel)
            unsigned long long cmask = ~0ull;
            if (op.leaf >= 0 && nseg > 0 && p.emask != nullptr && !p.canonical) cmask = p.emask[((size_t)op.leaf * p.R_pad + row0) / 8 + warp];
            if (op.kind == 1 && nseg > 0 && p.unit_rows != nullptr && row0 + warp * 8 < p.R) {
                // statistics, outside the GEMM loops: (warp, segment) pairs of this branch and those the warp multiplies in
                int act = 0;
                for (int si = 0; si < nseg; ++si) {
                    const int seg = ts->seg_lo[oi] + si;
                    const bool nz = gi >= 0 && ((seg == gi - 1 && ra != 0.0) || (seg == gi && rd != 0.0));
                    act += __any_sync(0xffffffffu, nz) ? 1 : 0;
                }
                if (lane == 0) { atomicAdd(p.unit_rows + 1, (unsigned long long)nseg); atomicAdd(p.unit_rows + 2, (unsigned long long)act); }
            }
            for (int si = 0; si < nseg; ++si) {
                double wgt = 1.0;
                if (op.kind == 1) {
                    const int seg = ts->seg_lo[oi] + si;
                    wgt = (gi >= 0) ? ((seg == gi - 1 ? ra : 0.0) + (seg == gi ? rd : 0.0)) : 0.0;
                }
                // a warp none of whose 8 rows blends this grid generation only keeps pace with the ring (its products would
                // all be zero); the warp it shares the sub-partition's DMMA pipe with then runs the segment at full rate
                const bool seg_active = (op.kind != 1) || __any_sync(0xffffffffu, wgt != 0.0);
                uint32_t an[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // chained A values of the coming chunk (tcgen05.ld in flight)
                if (op.leaf < 0) tmem_ld4_issue(t_chain, an);
                for (int c = 0; c < nchunks; ++c) {
                    {
                        const int rel = crel + PF4;
                        PROF_MARK(1)   // compute (tail of the previous chunk) + loop control
                        if (!no_copy && lane == 0 && rel == iss_rel && rel < total_chunks) issue_next();
                    }
                    __syncwarp();
                    PROF_MARK(2)   // copy issue
                    double ac[4] = {0.0, 0.0, 0.0, 0.0};
                    if (op.leaf < 0) {
                        tmem_wait_ld();
                        ac[0] = u2d(an[0], an[1]); ac[1] = u2d(an[2], an[3]); ac[2] = u2d(an[4], an[5]); ac[3] = u2d(an[6], an[7]);
                        if (c + 1 < nchunks) tmem_ld4_issue(t_chain + 8 * (c + 1), an);
                    }
                    const unsigned s = g % NS4;
                    PROF_MARK(1)
                    if (!no_copy) mbar_wait(&ts->full[s], (g / NS4) & 1u);
                    PROF_MARK(3)   // waiting for the stage
                    const int nq = (c + 1 == nchunks) ? nq_last : 4;
                    const unsigned char* st = ring + (size_t)s * Cfg::STAGE_BYTES;
                    const bool do_mma = seg_active && ((cmask >> c) & 1ull);
                    my_kcols += do_mma ? min(KC, p.F - KC * c) : 0;
                    if (!no_mma && do_mma) {
                        // the four operand flavours get their own straight-line code: nothing but LDS + DMMA between k steps
                        if (op.leaf >= 0) {
                            if (op.kind == 1) chunk_mma<NTT, NFULL, SCALAR, true, true, no_lds>(acc, sacc, st, a_base, b_base, z, bs_base, z0, ac, wgt, nq);
                            else chunk_mma<NTT, NFULL, SCALAR, true, false, no_lds>(acc, sacc, st, a_base, b_base, z, bs_base, z0, ac, wgt, nq);
                        } else {
                            if (op.kind == 1) chunk_mma<NTT, NFULL, SCALAR, false, true, no_lds>(acc, sacc, st, a_base, b_base, z, bs_base, z0, ac, wgt, nq);
                            else chunk_mma<NTT, NFULL, SCALAR, false, false, no_lds>(acc, sacc, st, a_base, b_base, z, bs_base, z0, ac, wgt, nq);
                        }
                    }
                    __syncwarp();
                    if (!no_copy && lane == 0) mbar_arrive(&ts->empty[s]);
                    ++g;
                    ++crel;
                }
            }

            if (SCALAR) {
                // column 8 NFULL: add the partial sums of the row's four kk lanes; lane kk = 0 owns the column
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
                sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
                acc[NFULL][0] = (kk == 0) ? sacc : 0.0;
            }
            PROF_MARK(1)
            // ---------------- post-op (warp local) ----------------
            const bool row_ident = !asc_uroot && (identity || (op.kind == 1 && gi < 0));
            if (asc_uroot && my_row < p.R) {
                // the whole subtree below this branch is the same for every family: its message was computed once per walker
                const double* v = p.uni_msg + (((size_t)w * p.n_uroot + op.uroot) * 2 + ((my_row - p.L) & 1)) * Fp + 2 * kk;
#pragma unroll
                for (int n = 0; n < NTT; ++n)
                    if (8 * n < Fp) { acc[n][0] = v[8 * n]; acc[n][1] = v[8 * n + 1]; }
            }
            if (asc_leaf && !row_ident) {
                const int which = 1 + ((my_row - p.L) & 1);   // block vector Fp + 1: low sums (P . e0), Fp + 2: high sums (P . eN)
                if (my_row < p.R && op.kind == 0) {
                    const double* v = p.pool + (size_t)ts->prow[oi] * Fp + (size_t)Fp * Fp + (size_t)which * Fp + 2 * kk;
#pragma unroll
                    for (int n = 0; n < NTT; ++n)
                        if (8 * n < Fp) { acc[n][0] = v[8 * n]; acc[n][1] = v[8 * n + 1]; }
                } else if (my_row < p.R) {
                    // blend of the two grid generations the row's time falls between; renormalised below like any other row
                    const double* v2 = p.pool + ri->off + (size_t)(gi - ri->gbase) * blk + (size_t)Fp * Fp + (size_t)which * Fp + 2 * kk;
                    const double* v1 = v2 - blk;
                    const bool two = (gi > 0 && ra != 0.0);
#pragma unroll
                    for (int n = 0; n < NTT; ++n) {
                        if (8 * n < Fp) {
                            double y0 = 0.0, y1 = 0.0;
                            if (two) { y0 = ra * v1[8 * n]; y1 = ra * v1[8 * n + 1]; }
                            acc[n][0] = fma(rd, v2[8 * n], y0);
                            acc[n][1] = fma(rd, v2[8 * n + 1], y1);
                        }
                    }
                }
            }
            // identity short-cut (transition_data_2d.py:184-187) or padding row: the message is the operand itself
            if (op.leaf >= 0) {
                if (row_ident) {
                    const double* e = p.E + (size_t)op.leaf * p.E_leaf_stride + (size_t)my_row * Fp;
#pragma unroll
                    for (int n = 0; n < NTT; ++n) {
                        if (8 * n < Fp) {
                            acc[n][0] = e[kperm(8 * n + 2 * kk)];
                            acc[n][1] = e[kperm(8 * n + 2 * kk + 1)];
                        }
                    }
                }
            } else if (__any_sync(0xffffffffu, row_ident)) {
                // tensor-memory loads are warp-wide: every lane reads, the lanes of identity rows keep the values
#pragma unroll
                for (int c = 0; c < NTT / 2; ++c) {
                    uint32_t r[8];
                    tmem_ld4_issue(t_chain + 8 * c, r);
                    tmem_wait_ld();
                    if (row_ident) {
                        acc[2 * c][0] = u2d(r[0], r[1]); acc[2 * c][1] = u2d(r[2], r[3]);
                        acc[2 * c + 1][0] = u2d(r[4], r[5]); acc[2 * c + 1][1] = u2d(r[6], r[7]);
                    }
                }
            }
            if (!row_ident && op.kind == 1) {
                // row renormalisation of the blended matrix (_util.pyx:138-147)
                const double* s2 = p.pool + ri->off + (size_t)(gi - ri->gbase) * blk + (size_t)Fp * Fp;
                const double* s1 = s2 - blk;
                const bool two = (gi > 0 && ra != 0.0);
#pragma unroll
                for (int n = 0; n < NTT; ++n) {
                    if (8 * n < Fp) {
                        const int col = 8 * n + 2 * kk;
                        double den0 = rd * s2[col], den1 = rd * s2[col + 1];
                        if (two) { den0 += ra * s1[col]; den1 += ra * s1[col + 1]; }
                        if (den0 > 1.0) acc[n][0] = renorm_div(acc[n][0], den0);
                        if (den1 > 1.0) acc[n][1] = renorm_div(acc[n][1], den1);
                    }
                }
            }

            if (p.canonical) {
                // canonical launch: rows L, L + 1 (parities 0, 1) of warp 0 publish the family-independent messages
                if (op.uni == 2 && op.leaf < 0 && warp == 0 && x < 2) {
                    double* v = p.uni_msg + (((size_t)w * p.n_uroot + op.uroot) * 2 + x) * Fp + 2 * kk;
#pragma unroll
                    for (int n = 0; n < NTT; ++n)
                        if (8 * n < Fp) { v[8 * n] = acc[n][0]; v[8 * n + 1] = acc[n][1]; }
                }
            }
            if (p.Yout != nullptr) {
                double* yo = p.Yout + ((size_t)w * p.R_pad + my_row) * Fp + 2 * kk;
#pragma unroll
                for (int n = 0; n < NTT; ++n)
                    if (8 * n < Fp) *reinterpret_cast<double2*>(yo + 8 * n) = make_double2(acc[n][0], acc[n][1]);
                { PROF_MARK(4) continue; }
            }

            __syncwarp();   // the tensor-memory instructions below are warp-wide
            if (op.store_slot >= 0) {
                // not the last child: park the message (or fold it into the siblings parked so far)
                if (!op.store_first) slot_mul(acc, op.home);
                slot_store(acc, op.home);
                { PROF_MARK(4) continue; }
            }
            if (op.mul_slot >= 0) {
                slot_mul(acc, op.home);
                slot_discard(op.home);
            }
            if (!op.last) {
                // chain: this product is the operand of the parent's branch, which is the next op
#pragma unroll
                for (int c = 0; c < NTT / 2; ++c)
                    tmem_st4(t_chain + 8 * c, acc[2 * c][0], acc[2 * c][1], acc[2 * c + 1][0], acc[2 * c + 1][1]);
                tmem_wait_st();
            } else {
                // root reduction (_likes.pyx:312-322, 426-442)
                const double* sv = p.stat + (size_t)w * Fp + 2 * kk;
                double dot = 0.0;
#pragma unroll
                for (int n = 0; n < NTT; ++n)
                    if (8 * n < Fp) dot += acc[n][0] * sv[8 * n] + acc[n][1] * sv[8 * n + 1];
                dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                dot += __shfl_xor_sync(0xffffffffu, dot, 2);
                if (kk == 0) {
                    double v = 0.0;
                    if (my_row < p.L) v = log(dot);
                    else if (my_row < p.R && !p.canonical) p.asc_dot[(size_t)w * (p.R - p.L) + (my_row - p.L)] = dot;
                    ts->rowval[warp * 8 + x] = v;
                }
            }
            PROF_MARK(4)   // post-op
        }  // ops

        if (p.unit_rows != nullptr && lane == 0 && my_kcols > 0) {
            const int real = max(0, min(8, p.R - (row0 + warp * 8)));
            if (real > 0) atomicAdd(p.unit_rows, (unsigned long long)my_kcols * (unsigned)real);
        }

        __syncthreads();   // rowval complete; every warp is done with the item's tables
        if (p.Yout == nullptr && !p.canonical && warp == 0) {
            double part = ts->rowval[lane] + ts->rowval[lane + 32];
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o2);
            if (lane == 0) p.tile_ll[(size_t)w * p.n_tiles + tile] = part;
        }
        __syncthreads();
        PROF_MARK(5)   // item end (barriers, tile sum)
    }  // items
#ifdef MOPE_PROF
    if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 77)) {
        const double tot = (double)(clock64() - pf_start);
        printf("cta %3d warp %d: total %.3e cyc; setup %.2f%% compute %.2f%% issue %.2f%% wait_full %.2f%% post %.2f%% item_end %.2f%% wait_empty %.2f%% tma %.2f%%\n", (int)blockIdx.x, warp, tot,
               100.0 * pf_t[0] / tot, 100.0 * pf_t[1] / tot, 100.0 * pf_t[2] / tot, 100.0 * pf_t[3] / tot, 100.0 * pf_t[4] / tot, 100.0 * pf_t[5] / tot,
               100.0 * pf_t[6] / tot, 100.0 * pf_t[7] / tot);
    }
#endif
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(ts->tmem_base), "n"(TMEM_COLS));
}

}  // namespace mope
