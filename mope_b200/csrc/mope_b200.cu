// mope_b200 host side of the C ABI (include/mope_b200.h).  No torch types; plain pointers and sizes.
//
// Host responsibilities (all tiny, exact, per call):
//   * the scalar part of TransitionData.get_transition_probabilities_2d -- range checks, searchsorted,
//     bilinear weights (transition_data_2d.py:148-238, transition_data_bottleneck.py:96-183) -- for every
//     (walker, matrix key), turned into InterpJob records for interp_kernel;
//   * the branch schedule (postorder walk of likelihoods.py:352-397) compiled once into TreeOp records;
//   * chunking walkers so that the per-chunk matrix pool fits the caller's workspace.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mope_b200.h"
#include "kernels.cuh"
#include "tree4.cuh"
#include "generate.cuh"
#include "selection.cuh"

using namespace mope;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(MOPE_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int round_up_i(int x, int a) { return (x + a - 1) / a * a; }

struct Bump {
    char* base = nullptr;
    size_t cap = 0, used = 0;
    bool dry = false;  // size query only
    template <typename T>
    T* take(size_t n) {
        used = align_up(used, 256);
        T* p = dry ? nullptr : reinterpret_cast<T*>(base + used);
        used += n * sizeof(T);
        return p;
    }
    bool ok() const { return dry || used <= cap; }
};

// selection model: parameter-independent row groups of a pedigree (selection.cuh), built at the first mope_b200_sel_loglike
struct SelCache {
    bool valid = false;
    std::vector<int32_t> pos_idx;            // the position index the cache was built for
    int n_pos = 0;
    std::vector<int> br_first, br_groups;    // per branch: first group (flat index) and number of groups
    int total_groups = 0;
    char* dev = nullptr;                     // one allocation: g_mult, g_pos, g_branch, g_local, g_row_off, g_rows
    double* g_mult = nullptr;
    int32_t *g_pos = nullptr, *g_branch = nullptr, *g_local = nullptr;
    int32_t* g_row_off = nullptr;            // [total_groups + n_branches]: per branch n_groups + 1 offsets into that branch's g_rows
    int32_t* g_rows = nullptr;               // [n_branches][3 L]
};

struct Pedigree {
    int n_leaves = 0, n_branches = 0, L = 0, n_fam = 0, n_mult = 0, n_asc = 0;
    int R = 0, R_pad = 0, n_tiles = 0, n_slots = 0, n_ops = 0;
    double* E = nullptr;
    unsigned long long* emask = nullptr;   // [n_leaves][R_pad / 8] non-zero K chunks of E per 8-row group (emission_mask_kernel)
    double* mult = nullptr;  // [n_mult][R_pad]
    int32_t* fam_asc = nullptr;
    double* fam_count = nullptr;
    double* fam_lgamma = nullptr;
    TreeOp* ops = nullptr;
    std::vector<TreeOp> ops_h;
    // warp-owned-rows kernel (tree4.cuh): postorder schedule with chained messages
    Op4* ops4 = nullptr;
    std::vector<Op4> ops4_h;
    int parks_chain4 = 0, parks_tm4 = 0, parks_l2_4 = 0;   // homes of the parked messages per item (tree4.cuh, HOME_*)
    int n_uroot4 = 0;               // family-independent subtrees of the ascertainment pass (top branches that are not leaves)
    int n_slots4 = 0;
    int asc_offset = 0;
    CUtensorMap tmapE;
    // host copies for the selection-model entry points (per-row plans are made on the host)
    std::vector<int32_t> br_parent_h, br_leaf_h, br_mult_h;
    std::vector<double> row_mult_h;   // [n_mult][L]
    SelCache sel;
};

struct FixedKey { int var; int bott; };
struct RateKey { int var; double min_mult, max_mult; };

}  // namespace

struct mope_ctx {
    int device = 0;
    int num_sms = 0;
    int F = 0, Fp = 0, NT = 0;
    int n_gens = 0, n_us = 0, n_nbs = 0, n_bot_us = 0;
    bool has_bot = false;
    double N = 0;
    std::vector<double> gens, us, nbs, bot_us, freqs;
    double min_coal = 0, max_coal = 0, min_mut = 0, max_mut = 0, min_bmut = 0, max_bmut = 0, min_nb = 0, max_nb = 0;
    // device tables
    double* tables = nullptr;  // drift [G][U][F][F] then bot [NB][UB][F][F]
    int64_t bot_base = 0;
    double* gens_dev = nullptr;
    int32_t* work_ctr = nullptr;   // item counter of tree_kernel4 (zeroed before every launch)
    int8_t* asc_mask = nullptr;    // [2][F] e0 (freq < min_freq) then eN (freq > 1 - min_freq)
    unsigned* lane_masks = nullptr; // [2][32] asc_mask as per-lane column bit masks (interp_kernel), null when F > 1024
    int32_t* breaks_dev = nullptr; // [F] first Wright-Fisher state of every bin (null: no device root prior)
    bool has_breaks = false;
    // constants of the device-side planner (library-owned, a few KB): grids and matrix keys
    char* plan_consts = nullptr;
    double *us_dev = nullptr, *nbs_dev = nullptr, *bot_us_dev = nullptr;
    FixedKeyDev* fkeys_dev = nullptr;
    RateKeyDev* rkeys_dev = nullptr;
    // model
    int n_var = 0, ndim = 0, n_ped = 0, asc_total = 0;
    double min_phred = -1, min_freq = 0, genome = 0, penalty = 0;
    std::vector<Pedigree> peds;
    std::vector<FixedKey> fixed_keys;
    std::vector<RateKey> rate_keys;
    int max_slots = 0, max_tiles = 0, max_asc2 = 0, max_uroot = 0;
    // host staging (pinned, grow-only)
    char* pinned = nullptr;
    size_t pinned_cap = 0, pinned_off = 0;
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};   // recorded when the staging ring leaves half 0 / half 1
    bool pin_ev_valid[2] = {false, false};
    int pin_half = 0;
    // small library-owned device scratch for the operator-level entry points
    char* opbuf = nullptr;
    size_t opbuf_cap = 0;
    int64_t launches = 0;
    int tree_variant = 13;
    bool use_v4 = false;   // F <= 208: warp-owned rows + k-permuted operand layout
    int ntt = 26;          // n-tiles per warp of tree_kernel4 (10, 16 or 26)
    int max_slots4 = 1;
    // statistics of the last eval: branch GEMMs executed / skipped by the identity short-cut (per walker x branch)
    int64_t last_gemm_ops = 0, last_identity_ops = 0;
    cudaStream_t last_stream = nullptr;
    // profiling
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_tree, ev_interp;
    std::vector<cudaEvent_t> ev_pool;
};

namespace {

int ensure_pinned(mope_ctx* c, size_t bytes) {
    if (bytes <= c->pinned_cap) return 0;
    if (c->pinned) cudaFreeHost(c->pinned);
    c->pinned = nullptr;
    c->pinned_cap = 0;
    size_t cap = align_up(bytes + bytes / 4, 1 << 16);
    CUDA_TRY(cudaMallocHost(&c->pinned, cap));
    c->pinned_cap = cap;
    return 0;
}

// Pinned staging as a ring of two halves: consecutive asynchronous calls take consecutive pieces.  When the ring moves on
// to the other half it records an event, and before it re-enters a half it waits for the event recorded when that half
// was left -- work enqueued a whole half-ring of calls ago, i.e. normally long finished.  (A plain stream synchronise on
// wrap-around would stall the host behind the pruning kernel in flight.)
int stage_alloc(mope_ctx* c, size_t bytes, cudaStream_t st, char** out) {
    bytes = align_up(std::max<size_t>(bytes, 64), 256);
    if (bytes * 2 > c->pinned_cap) {
        CUDA_TRY(cudaStreamSynchronize(st));
        if (int rc = ensure_pinned(c, std::max(bytes * 8, (size_t)16 << 20))) return rc;
        c->pinned_off = 0;
        c->pin_half = 0;
        c->pin_ev_valid[0] = c->pin_ev_valid[1] = false;
    }
    const size_t half = (c->pinned_cap / 2) & ~(size_t)255;
    if (c->pinned_off + bytes > (size_t)(c->pin_half + 1) * half) {
        if (!c->pin_ev[c->pin_half]) CUDA_TRY(cudaEventCreateWithFlags(&c->pin_ev[c->pin_half], cudaEventDisableTiming));
        CUDA_TRY(cudaEventRecord(c->pin_ev[c->pin_half], st));
        c->pin_ev_valid[c->pin_half] = true;
        c->pin_half ^= 1;
        c->pinned_off = (size_t)c->pin_half * half;
        if (c->pin_ev_valid[c->pin_half]) CUDA_TRY(cudaEventSynchronize(c->pin_ev[c->pin_half]));
    }
    *out = c->pinned + c->pinned_off;
    c->pinned_off += bytes;
    return 0;
}

int ensure_opbuf(mope_ctx* c, size_t bytes) {
    if (bytes <= c->opbuf_cap) return 0;
    if (c->opbuf) cudaFree(c->opbuf);
    c->opbuf = nullptr;
    c->opbuf_cap = 0;
    size_t cap = align_up(bytes, 1 << 20);
    CUDA_TRY(cudaMalloc(&c->opbuf, cap));
    c->opbuf_cap = cap;
    return 0;
}

// ln C(n, k) per cell exactly as the reference's GSL call chain evaluates it (mope/_binom.pyx:20-31 -> gsl_ran_binomial_pdf ->
// gsl_sf_lnchoose): counts and coverage cast to unsigned, m := min(k, n - k), lgamma(n+1) - lgamma(m+1) - lgamma(n-m+1) with
// libm's lgamma.  Host threads; cells with a NaN count or k > n are not read by the kernel.
void host_lnchoose(const double* counts, const double* coverage, size_t cells, double* out) {
    auto run = [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            const double x = counts[i], nn = coverage[i];
            double v = 0.0;
            if (!(x != x)) {
                const unsigned un = (unsigned)nn;
                unsigned um = (unsigned)x;
                if (um <= un && um != un && um != 0) {
                    if (um * 2u > un) um = un - um;
                    int sg;
                    v = lgamma_r((double)un + 1.0, &sg) - lgamma_r((double)um + 1.0, &sg) - lgamma_r((double)(un - um) + 1.0, &sg);
                }
            }
            out[i] = v;
        }
    };
    const unsigned nth = (unsigned)std::max<size_t>(1, std::min<size_t>({(size_t)16, (size_t)std::thread::hardware_concurrency(), cells >> 14}));
    std::vector<std::thread> th;
    for (unsigned k = 1; k < nth; ++k) th.emplace_back(run, cells * k / nth, cells * (k + 1) / nth);
    run(0, cells / nth);
    for (auto& t : th) t.join();
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 2-D fp64 tensor [rows][Fp] (row pitch Fp*8 bytes), box = 16 columns x box_rows rows, 128-byte swizzle
int make_tmap(CUtensorMap* out, const double* base, uint64_t rows, int Fp, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(MOPE_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (rows == 0) rows = 1;
    cuuint64_t gdim[2] = {(cuuint64_t)Fp, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)Fp * sizeof(double)};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MOPE_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows %llu, Fp %d, box %d)", (int)r,
                                       (unsigned long long)rows, Fp, box_rows);
    return 0;
}

int tree_prows(const mope_ctx* c) { return c->use_v4 ? c->ntt * 8 : 2 * c->tree_variant * 8; }
// doubles of message scratch for a full grid
size_t scratch_doubles(const mope_ctx* c) {
    return c->use_v4 ? (size_t)c->num_sms * 8 * (c->max_slots4 + 1) * c->ntt * 64 : (size_t)c->num_sms * c->max_slots * TM * c->Fp;
}

// searchsorted(side='left')
inline int lower_bound_idx(const std::vector<double>& v, double x) {
    return (int)(std::lower_bound(v.begin(), v.end(), x) - v.begin());
}

int validate(const mope_tables_desc* t, const mope_model_desc* m) {
    if (!t || !m) return fail(MOPE_E_INVALID, "null description");
    if (t->F < 2 || t->n_gens < 2 || t->n_us < 2 || !t->drift || !t->gens || !t->us || !t->freqs)
        return fail(MOPE_E_INVALID, "drift tables need F>=2 and a grid of at least 2x2");
    if (t->bot && (t->n_nbs < 2 || t->n_bot_us < 2 || !t->nbs || !t->bot_us))
        return fail(MOPE_E_INVALID, "bottleneck tables need a grid of at least 2x2");
    // The clamp of _util.check_transition_matrix (mope/_util.pyx:140-143) is a no-op on blends of entries in [0,1]; the
    // factored age-scaled branches rely on that, so the drift tables are checked here, whoever the caller is.
    {
        const size_t n = (size_t)t->n_gens * t->n_us * t->F * t->F;
        const unsigned nth = (unsigned)std::max<size_t>(1, std::min<size_t>({(size_t)16, (size_t)std::thread::hardware_concurrency(), n >> 22}));
        std::vector<double> mn(nth, 0.0), mx(nth, 0.0);
        std::vector<char> bad(nth, 0);
        auto scan = [&](unsigned k) {
            const size_t lo = n * k / nth, hi = n * (k + 1) / nth;
            double a = 0.0, b = 0.0;
            bool nan = false;
            for (size_t i = lo; i < hi; ++i) { const double v = t->drift[i]; a = v < a ? v : a; b = v > b ? v : b; nan |= (v != v); }
            mn[k] = a; mx[k] = b; bad[k] = nan;
        };
        std::vector<std::thread> th;
        for (unsigned k = 1; k < nth; ++k) th.emplace_back(scan, k);
        scan(0);
        for (auto& x : th) x.join();
        for (unsigned k = 0; k < nth; ++k)
            if (mn[k] < 0.0 || mx[k] > 1.0 || bad[k]) return fail(MOPE_E_INVALID, "drift tables hold entries outside [0, 1]");
    }
    for (int i = 1; i < t->n_gens; ++i) if (!(t->gens[i] > t->gens[i - 1])) return fail(MOPE_E_INVALID, "gens not ascending");
    for (int i = 1; i < t->n_us; ++i) if (!(t->us[i] > t->us[i - 1])) return fail(MOPE_E_INVALID, "us not ascending");
    if (m->n_var < 1 || m->n_ped < 1 || !m->peds) return fail(MOPE_E_INVALID, "model needs >=1 variable and >=1 pedigree");
    for (int k = 0; k < m->n_ped; ++k) {
        const mope_pedigree_desc& p = m->peds[k];
        if (p.n_leaves < 1 || p.n_branches < 1 || p.n_loci < 0 || p.n_fam < 1 || p.n_asc < 1 || p.n_mult < 0)
            return fail(MOPE_E_INVALID, "pedigree %d: bad sizes", k);
        for (int b = 0; b < p.n_branches; ++b) {
            if (p.br_parent[b] <= b || p.br_parent[b] > p.n_branches)
                return fail(MOPE_E_INVALID, "pedigree %d: branch %d parent %d violates postorder numbering", k, b, p.br_parent[b]);
            if (p.br_var[b] < 0 || p.br_var[b] >= m->n_var) return fail(MOPE_E_INVALID, "pedigree %d: bad var index", k);
            if (p.br_mult[b] >= p.n_mult) return fail(MOPE_E_INVALID, "pedigree %d: bad multiplier index", k);
            if (p.br_leaf[b] >= p.n_leaves) return fail(MOPE_E_INVALID, "pedigree %d: bad leaf index", k);
            if (p.br_bottleneck[b] && p.br_mult[b] >= 0)
                return fail(MOPE_E_INVALID, "pedigree %d: a multiplier on a bottleneck branch is not defined", k);
        }
        for (int f = 0; f < p.n_fam; ++f)
            if (p.fam_asc[f] < 0 || p.fam_asc[f] >= p.n_asc) return fail(MOPE_E_INVALID, "pedigree %d: bad fam_asc", k);
        // a multiplier (age) that is NA on a row an age-scaled branch uses: the reference raises
        // ValueError('scaled_time is not finite') at the first evaluation (transition_data_2d.py:135-136)
        for (int b = 0; b < p.n_branches; ++b) {
            const int mc = p.br_mult[b];
            if (mc < 0) continue;
            for (int r = 0; r < p.n_loci; ++r)
                if (!std::isfinite(p.row_mult[(size_t)mc * p.n_loci + r]))
                    return fail(MOPE_E_INVALID, "pedigree %d: multiplier column %d is not finite on locus row %d (scaled_time is not finite)", k, mc, r);
            for (int u = 0; u < p.n_asc; ++u)
                if (!std::isfinite(p.asc_mult[(size_t)mc * p.n_asc + u]))
                    return fail(MOPE_E_INVALID, "pedigree %d: multiplier column %d is not finite for family group %d (scaled_time is not finite)", k, mc, u);
        }
    }
    return 0;
}

// Compile the branch schedule.  Every branch b produces a message Y_b = P_b . content(b); content(b) is the
// element-wise product of the messages of b's children (or the leaf emissions).  A node keeps at most two
// pending messages: the third and later children fold into the second one (read-modify-write), so that a
// consumer never multiplies more than two operands.  Slots are recycled as soon as they are consumed.
int compile_schedule(const mope_pedigree_desc& pd, const std::vector<int>& fixed_key_of_branch,
                     const std::vector<int>& rate_key_of_branch, Pedigree& out) {
    const int nb = pd.n_branches, root = pd.n_branches;
    if (nb > MAX_OPS) return fail(MOPE_E_INVALID, "tree has %d branches; this build stages at most %d", nb, MAX_OPS);
    std::vector<int> nchild(nb + 1, 0), seen(nb + 1, 0), done_children(nb + 1, 0);
    std::vector<std::vector<int>> pending(nb + 1);  // message slots waiting at a node
    for (int b = 0; b < nb; ++b) nchild[pd.br_parent[b]]++;

    // Execution order: any topological order gives bit-identical results (every message has fixed operands).
    // Postorder makes most internal branches consume the message written by the op right before them, which
    // stalls the copy pipeline on that epilogue; so among the ready branches prefer the earliest (postorder)
    // one that does not depend on the op just scheduled.
    std::vector<int> order;
    {
        std::vector<char> scheduled(nb, 0);
        int prev = -1;
        for (int n = 0; n < nb; ++n) {
            int pick = -1, fallback = -1;
            for (int b = 0; b < nb; ++b) {
                if (scheduled[b] || done_children[b] != nchild[b]) continue;   // not ready
                if (fallback < 0) fallback = b;
                const bool dep = (prev >= 0 && pd.br_parent[prev] == b);
                if (!dep) { pick = b; break; }
            }
            if (pick < 0) pick = fallback;
            if (pick < 0) return fail(MOPE_E_INVALID, "branch table is not a tree");
            // keep the walk close to postorder: do not start a new leaf while an internal branch further left is
            // ready and independent (the scan above already returns the left-most candidate)
            scheduled[pick] = 1;
            done_children[pd.br_parent[pick]]++;
            order.push_back(pick);
            prev = pick;
        }
    }

    std::vector<int> free_slots;
    int n_slots = 0;
    int prev_dst = -1;
    out.ops_h.clear();
    for (int b : order) {
        TreeOp op{};
        const int par = pd.br_parent[b];
        op.sib[0] = op.sib[1] = -1;
        if (pd.br_leaf[b] >= 0) {
            if (nchild[b] != 0) return fail(MOPE_E_INVALID, "branch %d is marked leaf but has children", b);
            op.nsrc = 1;
            op.src_kind[0] = 0;
            op.src[0] = pd.br_leaf[b];
            op.src_kind[1] = 0;
            op.src[1] = 0;
            op.dep_prev = 0;
        } else {
            if (nchild[b] == 0 || pending[b].empty()) return fail(MOPE_E_INVALID, "internal node %d has no children", b);
            op.nsrc = (int)pending[b].size();
            for (int i = 0; i < op.nsrc; ++i) { op.src_kind[i] = 1; op.src[i] = pending[b][i]; }
            if (op.nsrc == 1) { op.src_kind[1] = 1; op.src[1] = op.src[0]; }
            op.dep_prev = 0;
            for (int i = 0; i < op.nsrc; ++i) if (op.src[i] == prev_dst) op.dep_prev = 1;
        }
        seen[par]++;
        const bool completes_root = (par == root && seen[par] == nchild[par]);
        op.last = completes_root ? 1 : 0;
        if (completes_root) {
            op.dst = -1;
            op.first = 1;
            for (size_t i = 0; i < pending[par].size(); ++i) op.sib[i] = pending[par][i];
        } else if (pending[par].size() < 2) {
            int sl;
            // most recently freed first: that tile is still dirty in L2 and gets overwritten there instead of being
            // written back to HBM
            if (!free_slots.empty()) { sl = free_slots.back(); free_slots.pop_back(); }
            else sl = n_slots++;
            op.dst = sl;
            op.first = 1;
            pending[par].push_back(sl);
        } else {
            op.dst = pending[par][1];
            op.first = 0;
        }
        if (pd.br_mult[b] >= 0) { op.kind = 1; op.key = rate_key_of_branch[b]; op.mult = pd.br_mult[b]; }
        else { op.kind = 0; op.key = fixed_key_of_branch[b]; op.mult = -1; }
        out.ops_h.push_back(op);
        prev_dst = op.dst;
        // the operands are free once this op has run
        if (pd.br_leaf[b] < 0) for (int sl : pending[b]) free_slots.push_back(sl);
    }
    if (out.ops_h.empty() || !out.ops_h.back().last) return fail(MOPE_E_INVALID, "schedule does not end at the root");
    out.n_slots = std::max(n_slots, 1);
    out.n_ops = (int)out.ops_h.size();
    return 0;
}

// Schedule of tree_kernel4: plain postorder (branch b's last child is branch b-1, so its message is still in the
// warp's fragments when b runs).  Messages of non-last children are parked in a slot; a node never holds more than
// one parked product (later siblings fold into it).
int compile_schedule4(const mope_pedigree_desc& pd, const std::vector<int>& fixed_key_of_branch,
                      const std::vector<int>& rate_key_of_branch, int tm_slots, Pedigree& out) {
    const int nb = pd.n_branches, root = pd.n_branches;
    std::vector<int> nchild(nb + 1, 0), seen(nb + 1, 0), slot(nb + 1, -1);
    for (int b = 0; b < nb; ++b) nchild[pd.br_parent[b]]++;
    std::vector<int> free_slots;
    int n_slots = 0;
    out.ops4_h.clear();
    for (int b = 0; b < nb; ++b) {
        Op4 op{};
        const int par = pd.br_parent[b];
        op.leaf = pd.br_leaf[b];
        if (op.leaf < 0) {
            // chained operand: the previous op must be this node's last child
            if (b == 0 || pd.br_parent[b - 1] != b || seen[b] != nchild[b])
                return fail(MOPE_E_INVALID, "branch table is not in postorder (branch %d)", b);
        } else if (nchild[b] != 0) {
            return fail(MOPE_E_INVALID, "branch %d is marked leaf but has children", b);
        }
        if (pd.br_mult[b] >= 0) { op.kind = 1; op.key = rate_key_of_branch[b]; op.mult = pd.br_mult[b]; }
        else { op.kind = 0; op.key = fixed_key_of_branch[b]; op.mult = -1; }
        seen[par]++;
        const bool last_child = (seen[par] == nchild[par]);
        op.mul_slot = -1; op.store_slot = -1; op.store_first = 0; op.last = 0;
        if (!last_child) {
            if (slot[par] < 0) {
                if (!free_slots.empty()) { slot[par] = free_slots.back(); free_slots.pop_back(); }
                else slot[par] = n_slots++;
                op.store_first = 1;
            }
            op.store_slot = slot[par];
        } else {
            op.mul_slot = slot[par];
            if (slot[par] >= 0) { free_slots.push_back(slot[par]); slot[par] = -1; }
            op.last = (par == root) ? 1 : 0;
            if (par != root && par != b + 1) return fail(MOPE_E_INVALID, "branch table is not in postorder (parent of %d)", b);
        }
        out.ops4_h.push_back(op);
    }
    if (out.ops4_h.empty() || !out.ops4_h.back().last) return fail(MOPE_E_INVALID, "schedule does not end at the root");
    out.n_slots4 = std::max(n_slots, 1);
    // Ascertainment pseudo-rows: a branch whose subtree holds no age-scaled branch carries the same message for every
    // family (uni != 0); the top branch of such a subtree (uni == 2) is computed once per walker (tree4.cuh, canonical launch)
    {
        std::vector<char> uniform(nb, 1);
        for (int b = 0; b < nb; ++b) {   // postorder: children come before their parent
            if (out.ops4_h[b].kind != 0) uniform[b] = 0;
            const int par = pd.br_parent[b];
            if (par != root && !uniform[b]) uniform[par] = 0;
        }
        out.n_uroot4 = 0;
        for (int b = 0; b < nb; ++b) {
            Op4& o = out.ops4_h[b];
            const int par = pd.br_parent[b];
            o.uni = !uniform[b] ? 0 : ((par == root || !uniform[par]) ? 2 : 1);
            o.uroot = (o.uni == 2 && o.leaf < 0) ? out.n_uroot4++ : -1;
        }
    }
    // Homes of the parked products (tree4.cuh, HOME_*).  A park lives from the post-op of the op that opens it (store_first)
    // to the post-op of the parent's last child (mul_slot).
    struct Park { int open, close, home; };
    std::vector<Park> parks;
    {
        std::vector<int> open_of_slot(n_slots, -1);
        for (int b = 0; b < nb; ++b) {
            const Op4& o = out.ops4_h[b];
            if (o.store_slot >= 0 && o.store_first) open_of_slot[o.store_slot] = b;
            if (o.mul_slot >= 0) { parks.push_back({open_of_slot[o.mul_slot], b, HOME_L2}); open_of_slot[o.mul_slot] = -1; }
        }
    }
    std::vector<Park*> rest;
    for (Park& pk : parks) {
        // the chain buffer is idle while only leaf branches run (they take their operand from the emission tensor and,
        // not being a last child before `close`, do not write it either)
        bool leaves_only = true;
        for (int b = pk.open + 1; b <= pk.close; ++b) if (out.ops4_h[b].leaf < 0) leaves_only = false;
        if (leaves_only) pk.home = HOME_CHAIN; else rest.push_back(&pk);
    }
    // interval scheduling on tm_slots tensor-memory slots: earliest close first, best fit -- parks the largest possible
    // number of messages there; what is left goes to the L2 scratch (slots recycled most-recently-freed first)
    std::sort(rest.begin(), rest.end(), [](const Park* a, const Park* b) { return a->close < b->close; });
    std::vector<int> tm_busy_until(std::max(tm_slots, 0), -1);
    std::vector<Park*> l2;
    for (Park* pk : rest) {
        int best = -1;
        for (int k = 0; k < (int)tm_busy_until.size(); ++k)
            if (tm_busy_until[k] < pk->open && (best < 0 || tm_busy_until[k] > tm_busy_until[best])) best = k;
        if (best >= 0) { pk->home = best; tm_busy_until[best] = pk->close; } else l2.push_back(pk);
    }
    std::sort(l2.begin(), l2.end(), [](const Park* a, const Park* b) { return a->open < b->open; });
    int n_l2 = 0;
    {
        std::vector<int> busy_until;   // per L2 slot
        for (Park* pk : l2) {
            int best = -1;
            for (int k = 0; k < (int)busy_until.size(); ++k)
                if (busy_until[k] < pk->open && (best < 0 || busy_until[k] > busy_until[best])) best = k;
            if (best < 0) { busy_until.push_back(-1); best = (int)busy_until.size() - 1; }
            busy_until[best] = pk->close;
            pk->home = HOME_L2 - best;
        }
        n_l2 = (int)busy_until.size();
    }
    out.n_slots4 = std::max(n_l2, 1);
    out.parks_chain4 = out.parks_tm4 = out.parks_l2_4 = 0;
    for (const Park& pk : parks) {
        if (pk.home == HOME_CHAIN) out.parks_chain4++; else if (pk.home >= 0) out.parks_tm4++; else out.parks_l2_4++;
        for (int b = pk.open; b <= pk.close; ++b) {
            Op4& o = out.ops4_h[b];
            // every op that touches this park: the opener and folders (store_slot) and the closer (mul_slot)
            if ((o.store_slot >= 0 && pd.br_parent[b] == pd.br_parent[pk.open]) || b == pk.close) o.home = pk.home;
        }
    }
    return 0;
}

struct ArenaPlan {
    size_t bytes = 0;
};

// lays out (or just sizes) everything that lives in the arena
int layout_arena(mope_ctx* c, const mope_tables_desc* t, const mope_model_desc* m, Bump& bump, bool dry) {
    const size_t FF = (size_t)t->F * t->F;
    size_t ntab = (size_t)t->n_gens * t->n_us * FF + (t->bot ? (size_t)t->n_nbs * t->n_bot_us * FF : 0);
    double* tab = bump.take<double>(ntab);
    double* gd = bump.take<double>(t->n_gens);
    int32_t* wc = bump.take<int32_t>(64);
    int8_t* am = bump.take<int8_t>(2 * (size_t)t->F);
    int32_t* brk = bump.take<int32_t>(t->F);
    unsigned* lm = bump.take<unsigned>(64);
    if (!dry && t->F <= 1024) c->lane_masks = lm;
    if (!dry) { c->tables = tab; c->gens_dev = gd; c->work_ctr = wc; c->asc_mask = am; c->breaks_dev = brk; c->bot_base = (int64_t)t->n_gens * t->n_us * FF; }
    const int Fp = round_up_i(t->F, 16);
    for (int k = 0; k < m->n_ped; ++k) {
        const mope_pedigree_desc& pd = m->peds[k];
        const int R = pd.n_loci + 2 * pd.n_asc;
        const int R_pad = round_up_i(std::max(R, 1), TM);
        double* E = bump.take<double>((size_t)pd.n_leaves * R_pad * Fp);
        unsigned long long* em = bump.take<unsigned long long>((size_t)pd.n_leaves * (R_pad / 8));
        double* mu = bump.take<double>((size_t)std::max(pd.n_mult, 1) * R_pad);
        int32_t* fa = bump.take<int32_t>(pd.n_fam);
        double* fc = bump.take<double>(pd.n_fam);
        double* fl = bump.take<double>(pd.n_fam);
        TreeOp* ops = bump.take<TreeOp>(pd.n_branches);
        Op4* ops4 = bump.take<Op4>(pd.n_branches);
        if (!dry) {
            Pedigree& P = c->peds[k];
            P.E = E; P.emask = em; P.mult = mu; P.fam_asc = fa; P.fam_count = fc; P.fam_lgamma = fl; P.ops = ops; P.ops4 = ops4;
        }
    }
    // transient inputs of the emission kernel live at the arena tail; they are dead after create().
    // Size it with the same take() sequence create() uses (each take is 256-byte aligned).
    size_t max_cells = 1;
    for (int k = 0; k < m->n_ped; ++k) max_cells = std::max(max_cells, (size_t)m->peds[k].n_loci * m->peds[k].n_leaves);
    if (dry) {
        bump.take<double>(max_cells);
        bump.take<double>(max_cells);
        bump.take<double>(max_cells);
        bump.take<double>(t->F);
        bump.take<double>(t->F);
        bump.take<int8_t>(t->F);
        bump.take<int8_t>(t->F);
        bump.take<int8_t>(t->F);
    }
    return 0;
}

cudaEvent_t get_event(mope_ctx* c) {
    if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

struct ScopedTimer {
    mope_ctx* c; cudaStream_t st; std::vector<std::pair<cudaEvent_t, cudaEvent_t>>* sink; cudaEvent_t e0 = nullptr;
    ScopedTimer(mope_ctx* c_, cudaStream_t st_, std::vector<std::pair<cudaEvent_t, cudaEvent_t>>* sink_) : c(c_), st(st_), sink(sink_) {
        if (c->profiling) { e0 = get_event(c); cudaEventRecord(e0, st); }
    }
    ~ScopedTimer() {
        if (e0) { cudaEvent_t e1 = get_event(c); cudaEventRecord(e1, st); sink->push_back({e0, e1}); }
    }
};

template <int NTW>
int launch_tree_t(mope_ctx* c, const TreeParams& tp, int grid, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(tree_kernel<NTW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileCfg<NTW>::SMEM_BYTES));
        attr_set = true;
    }
    CUDA_TRY(cudaMemsetAsync(c->work_ctr, 0, sizeof(int32_t), st));
    {
        ScopedTimer timer(c, st, &c->ev_tree);
        TreeParams tq = tp;
        tq.work_ctr = c->work_ctr;
        tq.unit_rows = reinterpret_cast<unsigned long long*>(c->work_ctr + 8);
        tree_kernel<NTW><<<grid, NTHREADS, TileCfg<NTW>::SMEM_BYTES, st>>>(tq);
    }
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    return 0;
}

// one CTA per (matrix, tile of rows): see interp_kernel
template <bool APPLY>
int launch_interp_t(mope_ctx* c, const InterpJob* d_jobs, size_t n_jobs, double* d_pool, cudaStream_t st, bool perm, const ApplyArgs& ap) {
    const int F = c->F, Fp = c->Fp;
    const bool bulk = interp_use_bulk(F);
    const int rows = bulk ? interp_rows_per_tile(F) : interp_rows_per_tile_ldg(F);
    const int tiles = (Fp + rows - 1) / rows;
    const size_t smem = bulk ? (size_t)4 * interp_buf_doubles(rows, F) * sizeof(double) : (size_t)rows * interp_row_stride_ldg(F) * sizeof(double);
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(interp_kernel<true, APPLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(interp_kernel<true, APPLY>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUDA_TRY(cudaFuncSetAttribute(interp_kernel<false, APPLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(interp_kernel<false, APPLY>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    if (!bulk && (size_t)rows * F >= 65536) return fail(MOPE_E_INVALID, "interp: tile of %d x %d elements exceeds the index range of the fast division", rows, F);
    const unsigned magic = (unsigned)((0x100000000ull + (unsigned long long)F - 1) / (unsigned long long)F);
    const unsigned long long grid = (unsigned long long)n_jobs * tiles;
    if (grid > 0x7fffffffull) return fail(MOPE_E_INVALID, "interp: too many matrices in one launch (%zu)", n_jobs);
    if (bulk)
        interp_kernel<true, APPLY><<<(unsigned)grid, INTERP_THREADS, smem, st>>>(d_jobs, (int)n_jobs, c->tables, d_pool, F, Fp, perm ? 1 : 0, c->asc_mask,
                                                                                 c->lane_masks, rows, tiles, magic, ap);
    else
        interp_kernel<false, APPLY><<<(unsigned)grid, INTERP_THREADS, smem, st>>>(d_jobs, (int)n_jobs, c->tables, d_pool, F, Fp, perm ? 1 : 0, c->asc_mask,
                                                                                  c->lane_masks, rows, tiles, magic, ap);
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    return 0;
}
int launch_interp(mope_ctx* c, const InterpJob* d_jobs, size_t n_jobs, double* d_pool, cudaStream_t st, bool perm) {
    return launch_interp_t<false>(c, d_jobs, n_jobs, d_pool, st, perm, ApplyArgs{});
}
// interpolate and apply in one pass (selection model, masked operators): no matrix leaves the SM.  Fp <= 32 * APPLY_MAXJ.
int launch_interp_apply(mope_ctx* c, const InterpJob* d_jobs, size_t n_jobs, cudaStream_t st, const ApplyArgs& ap) {
    return launch_interp_t<true>(c, d_jobs, n_jobs, nullptr, st, false, ap);
}

template <int NTT, int NFULL, bool SCALAR, bool ASCFAST>
int launch_tree4_t(mope_ctx* c, const Tree4Params& tp, int grid, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(tree_kernel4<NTT, NFULL, SCALAR, ASCFAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg4<NTT>::SMEM_BYTES));
        attr_set = true;
    }
    CUDA_TRY(cudaMemsetAsync(c->work_ctr, 0, sizeof(int32_t), st));
    {
        ScopedTimer timer(c, st, &c->ev_tree);
        Tree4Params tq = tp;
        tq.work_ctr = c->work_ctr;
        tq.unit_rows = reinterpret_cast<unsigned long long*>(c->work_ctr + 8);
        tree_kernel4<NTT, NFULL, SCALAR, ASCFAST><<<grid, NTHREADS, Cfg4<NTT>::SMEM_BYTES, st>>>(tq);
    }
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    return 0;
}

template <int NTT, int NFULL, bool SCALAR>
int launch_tree4_a(mope_ctx* c, const Tree4Params& tp, int grid, cudaStream_t st) {
    // the ascertainment fast path (leaf messages of pure pseudo-row tiles from the low / high sums, no GEMM) is a separate
    // instance: pedigrees without such tiles (cfg2: two pseudo-rows) keep the leaner code
    const bool asc_fast = tp.Yout == nullptr && (tp.R_pad - (tp.L + TM - 1) / TM * TM) >= TM && !getenv("MOPE_NO_ASC_FAST");
    if (asc_fast) return launch_tree4_t<NTT, NFULL, SCALAR, true>(c, tp, grid, st);
    return launch_tree4_t<NTT, NFULL, SCALAR, false>(c, tp, grid, st);
}

// n-tiles that go through the tensor pipe: all-padding tiles are skipped, and a last tile with a single real column
// (F = 8k + 1: the grid sizes 201 and 65) is done by scalar FMAs (tree4.cuh).  Specialisations exist for the grid
// sizes in use; any other F runs the generic instance of its size class.
int launch_tree4(mope_ctx* c, const Tree4Params& tp, int grid, cudaStream_t st) {
    const int full = c->F / 8, rem = c->F % 8;
    const int nf = rem <= 1 ? full : full + 1;
    const bool sc = rem == 1;
    const bool generic = getenv("MOPE_TREE4_GENERIC") != nullptr;
    switch (c->ntt) {
        case 10:
            if (!generic && nf == 8 && sc) return launch_tree4_a<10, 8, true>(c, tp, grid, st);
            return launch_tree4_a<10, 10, false>(c, tp, grid, st);
        case 16:
            if (!generic && nf == 15 && !sc) return launch_tree4_a<16, 15, false>(c, tp, grid, st);
            return launch_tree4_a<16, 16, false>(c, tp, grid, st);
        default:
            if (!generic && nf == 25 && sc) return launch_tree4_a<26, 25, true>(c, tp, grid, st);
            return launch_tree4_a<26, 26, false>(c, tp, grid, st);
    }
}

int launch_tree(mope_ctx* c, const TreeParams& tp, int grid, cudaStream_t st) {
    switch (c->tree_variant) {
        case 5: return launch_tree_t<5>(c, tp, grid, st);
        case 8: return launch_tree_t<8>(c, tp, grid, st);
        default: return launch_tree_t<13>(c, tp, grid, st);
    }
}

// ------------------------------------------------------------------------------------------------
// per-walker interpolation plan
// ------------------------------------------------------------------------------------------------
struct WalkerPlan {
    int32_t status = 0;
    size_t n_blocks = 0;  // matrix blocks this walker needs in the pool
};

struct PlanOut {
    std::vector<InterpJob> jobs;   // dst offsets relative to the chunk pool
    std::vector<int64_t> mat_off;  // [Wc][n_mat]
    std::vector<RateInfo> rate;    // [Wc][n_rate]
};

// range checks of one non-rate key (transition_data_2d.py:158-173, transition_data_bottleneck.py:118-129)
int32_t status_fixed(const mope_ctx* c, bool bott, double len, double mut) {
    int32_t st = 0;
    if (bott) {
        if (!c->has_bot) return MOPE_ST_NO_BOTTLENECK_TABLE;
        if (std::isnan(len) || std::isnan(mut)) st |= MOPE_ST_NOT_FINITE;
        if (len < c->min_nb) st |= MOPE_ST_NB_TOO_SMALL;
        if (len > c->max_nb) st |= MOPE_ST_NB_TOO_LARGE;
        if (mut < c->min_bmut) st |= MOPE_ST_MUT_TOO_SMALL;
        if (mut > c->max_bmut) st |= MOPE_ST_MUT_TOO_LARGE;
    } else {
        if (!std::isfinite(len) || !std::isfinite(mut)) st |= MOPE_ST_NOT_FINITE;
        if (len > c->max_coal) st |= MOPE_ST_TIME_TOO_LARGE;
        if (mut < c->min_mut) st |= MOPE_ST_MUT_TOO_SMALL;
        if (mut > c->max_mut) st |= MOPE_ST_MUT_TOO_LARGE;
    }
    return st;
}

// range checks of a rate key: times are mult*len, monotone in mult, so the extreme multipliers decide
int32_t status_rate(const mope_ctx* c, double min_mult, double max_mult, double len, double mut) {
    int32_t st = 0;
    const double tmax = max_mult * len, tmin = min_mult * len;
    if (!std::isfinite(tmax) || !std::isfinite(tmin) || !std::isfinite(mut)) st |= MOPE_ST_NOT_FINITE;
    if (tmax > c->max_coal) st |= MOPE_ST_TIME_TOO_LARGE;
    if (mut < c->min_mut) st |= MOPE_ST_MUT_TOO_SMALL;
    if (mut > c->max_mut) st |= MOPE_ST_MUT_TOO_LARGE;
    return st;
}

// generation-grid range [glo, ghi] a rate key needs; false when every row takes the identity short-cut
bool rate_range(const mope_ctx* c, double min_mult, double max_mult, double len, int& glo, int& ghi) {
    const double gmax = c->N * (max_mult * len), gmin = c->N * (min_mult * len);
    if (gmax < c->gens[0]) return false;
    ghi = std::min(lower_bound_idx(c->gens, gmax), c->n_gens - 1);
    glo = (gmin < c->gens[0]) ? 0 : std::max(std::min(lower_bound_idx(c->gens, gmin), c->n_gens - 1) - 1, 0);
    return true;
}

// one non-rate matrix with exact (len, mut); b = block index inside the chunk pool
void emit_fixed_job(const mope_ctx* c, bool bott, double len, double mut, size_t& b, PlanOut& out) {
    const size_t FF = (size_t)c->F * c->F;
    const size_t blk = blk_doubles(c->Fp);
    if (!bott && c->N * len < c->gens[0]) {
        // identity short-cut (transition_data_2d.py:184-187): nothing to interpolate, the kernel skips the GEMM
        out.mat_off.push_back(-1);
        return;
    }
    InterpJob jb{};
    jb.dst = (int64_t)(b * blk);
    out.mat_off.push_back(jb.dst);
    ++b;
    if (bott) {
        // transition_data_bottleneck.py:138-171
        const int nbn = c->n_nbs, nun = c->n_bot_us;
        int bi = std::min(lower_bound_idx(c->nbs, len), nbn - 1);
        const double u = mut / (2.0 * c->N);
        int ui = std::min(lower_bound_idx(c->bot_us, u), nun - 1);
        const int b1 = bi > 0 ? bi - 1 : nbn - 1, u1i = ui > 0 ? ui - 1 : nun - 1;  // python's index -1 wraps
        const double t = len, t1 = c->nbs[b1], t2 = c->nbs[bi], u1 = c->bot_us[u1i], u2 = c->bot_us[ui];
        const double denom = (t2 - t1) * (u2 - u1);
        jb.w[0] = (t2 - t) * (u2 - u) / denom;  // w11
        jb.w[2] = (t - t1) * (u2 - u) / denom;  // w21
        jb.w[1] = (t2 - t) * (u - u1) / denom;  // w12
        jb.w[3] = (t - t1) * (u - u1) / denom;  // w22
        jb.src[0] = c->bot_base + (int64_t)(((size_t)b1 * nun + u1i) * FF);
        jb.src[1] = c->bot_base + (int64_t)(((size_t)b1 * nun + ui) * FF);
        jb.src[2] = c->bot_base + (int64_t)(((size_t)bi * nun + u1i) * FF);
        jb.src[3] = c->bot_base + (int64_t)(((size_t)bi * nun + ui) * FF);
        jb.flags = 0;
    } else {
        // transition_data_2d.py:183-224
        const double t = c->N * len;
        {
            const int G = c->n_gens, U = c->n_us;
            const int gi = std::min(lower_bound_idx(c->gens, t), G - 1);
            const double xd = mut / (2.0 * c->N);
            const int xi = std::min(lower_bound_idx(c->us, xd), U - 1);
            const int g1 = gi > 0 ? gi - 1 : G - 1, x1i = xi > 0 ? xi - 1 : U - 1;
            const double t1 = c->gens[g1], t2 = c->gens[gi], x1 = c->us[x1i], x2 = c->us[xi];
            const double denom = (t2 - t1) * (x2 - x1);
            jb.w[0] = (t2 - t) * (x2 - xd) / denom;
            jb.w[2] = (t - t1) * (x2 - xd) / denom;
            jb.w[1] = (t2 - t) * (xd - x1) / denom;
            jb.w[3] = (t - t1) * (xd - x1) / denom;
            jb.src[0] = (int64_t)(((size_t)g1 * U + x1i) * FF);
            jb.src[1] = (int64_t)(((size_t)g1 * U + xi) * FF);
            jb.src[2] = (int64_t)(((size_t)gi * U + x1i) * FF);
            jb.src[3] = (int64_t)(((size_t)gi * U + xi) * FF);
            jb.flags = 1;
        }
    }
    out.jobs.push_back(jb);
}

// the mutation-axis blends A_g = bx*Q[g,x1] + cx*Q[g,x2] for every generation-grid index a rate key touches;
// the time-axis blend is per row and happens inside the tree kernel
void emit_rate_jobs(const mope_ctx* c, double min_mult, double max_mult, double len, double mut, size_t& b, PlanOut& out) {
    const size_t FF = (size_t)c->F * c->F;
    const size_t blk = blk_doubles(c->Fp);
    RateInfo ri{};
    ri.len = len;
    ri.off = (int64_t)(b * blk);
    ri.gbase = 0;
    ri.gcount = 0;
    int glo, ghi;
    if (rate_range(c, min_mult, max_mult, len, glo, ghi)) {
        const int U = c->n_us;
        const double xd = mut / (2.0 * c->N);
        const int xi = std::min(lower_bound_idx(c->us, xd), U - 1);
        const int x1i = xi > 0 ? xi - 1 : U - 1;
        const double x1 = c->us[x1i], x2 = c->us[xi];
        const double bx = (x2 - xd) / (x2 - x1), cx = (xd - x1) / (x2 - x1);
        ri.gbase = glo;
        ri.gcount = ghi - glo + 1;
        for (int g = glo; g <= ghi; ++g) {
            InterpJob jb{};
            jb.dst = (int64_t)(b * blk);
            ++b;
            jb.src[0] = (int64_t)(((size_t)g * U + x1i) * FF);
            jb.src[1] = (int64_t)(((size_t)g * U + xi) * FF);
            jb.src[2] = jb.src[0];
            jb.src[3] = jb.src[1];
            jb.w[0] = bx; jb.w[1] = cx; jb.w[2] = 0; jb.w[3] = 0;
            jb.flags = 4;  // two-term blend; the clamp/renormalisation is applied per row in the tree kernel
            out.jobs.push_back(jb);
        }
    }
    out.rate.push_back(ri);
}

// status + number of blocks for one walker (no jobs yet)
void plan_walker_size(const mope_ctx* c, const double* x, WalkerPlan& wp) {
    wp.status = 0;
    wp.n_blocks = 0;
    const int nv = c->n_var;
    for (const FixedKey& k : c->fixed_keys) {
        const double len = std::pow(10.0, x[k.var]);
        wp.status |= status_fixed(c, k.bott != 0, len, std::pow(10.0, x[nv + k.var]));
        if (k.bott || !(c->N * len < c->gens[0])) wp.n_blocks += 1;
    }
    for (const RateKey& k : c->rate_keys) {
        const double len = std::pow(10.0, x[k.var]);
        const int32_t st = status_rate(c, k.min_mult, k.max_mult, len, std::pow(10.0, x[nv + k.var]));
        wp.status |= st;
        int glo, ghi;
        if (!st && rate_range(c, k.min_mult, k.max_mult, len, glo, ghi)) wp.n_blocks += (size_t)(ghi - glo + 1);
    }
}

// emit jobs for one (valid) walker; block0 is the walker's first block inside the chunk pool
void plan_walker_jobs(const mope_ctx* c, const double* x, size_t block0, PlanOut& out) {
    const int nv = c->n_var;
    size_t b = block0;
    for (const FixedKey& k : c->fixed_keys)
        emit_fixed_job(c, k.bott != 0, std::pow(10.0, x[k.var]), std::pow(10.0, x[nv + k.var]), b, out);
    for (const RateKey& k : c->rate_keys)
        emit_rate_jobs(c, k.min_mult, k.max_mult, std::pow(10.0, x[k.var]), std::pow(10.0, x[nv + k.var]), b, out);
}

// fixed (walker-independent) part of the eval workspace
struct WsFixed {
    size_t scratch_doubles, out_ll, out_root, out_asc;
};

size_t per_walker_bytes(const mope_ctx* c, size_t n_blocks) {
    const size_t blk = blk_doubles(c->Fp) * sizeof(double);
    size_t b = n_blocks * (blk + sizeof(InterpJob));
    b += c->fixed_keys.size() * sizeof(int64_t) + c->rate_keys.size() * sizeof(RateInfo);
    b += (size_t)c->Fp * sizeof(double);                 // stat
    b += (size_t)c->max_tiles * sizeof(double);          // tile_ll
    b += (size_t)c->max_asc2 * sizeof(double);           // asc_dot
    b += (size_t)c->max_uroot * 2 * c->Fp * sizeof(double);   // family-independent ascertainment messages
    return b + 64;
}

size_t fixed_ws_bytes(const mope_ctx* c, int W) {
    size_t b = scratch_doubles(c) * sizeof(double);  // scratch
    b += (size_t)W * sizeof(double) * (1 + c->n_ped);                             // out_ll, out_root
    b += (size_t)W * c->asc_total * sizeof(double);                               // out_log_asc
    b += (size_t)W * sizeof(int32_t);                                             // walker index
    return b + 16 * 256 + 4096;
}

size_t max_blocks_per_walker(const mope_ctx* c) {
    return c->fixed_keys.size() + c->rate_keys.size() * (size_t)c->n_gens;
}

// the core: evaluate walkers whose params are on the host; stat may be a host or device pointer
int eval_impl(mope_ctx* c, const double* params, const double* stat, bool stat_on_device, int W, void* workspace,
              size_t ws_bytes, cudaStream_t st, double* out_ll, bool out_on_device, int32_t* out_status,
              double* out_root_ll, double* out_log_asc) {
    if (!c || !params || !stat || !out_ll || W < 0) return fail(MOPE_E_INVALID, "eval: null argument");
    if (W == 0) return 0;
    if (!workspace) return fail(MOPE_E_INVALID, "eval: null workspace");
    CUDA_TRY(cudaSetDevice(c->device));
    const int ndim = c->ndim, Fp = c->Fp, F = c->F;
    const int n_mat = (int)c->fixed_keys.size(), n_rate = (int)c->rate_keys.size();

    // host-side phase timing (MOPE_HOST_PROF=<ms>: print the phases of calls whose host part exceeds <ms>)
    static const double host_prof_ms = getenv("MOPE_HOST_PROF") ? atof(getenv("MOPE_HOST_PROF")) : -1.0;
    std::vector<std::pair<const char*, double>> hp_marks;
    const auto hp_t0 = std::chrono::steady_clock::now();
    auto hp_mark = [&](const char* name) {
        if (host_prof_ms >= 0.0) hp_marks.emplace_back(name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - hp_t0).count());
    };
    // 1. plan sizes / statuses
    std::vector<WalkerPlan> plans(W);
    std::vector<int> good;
    good.reserve(W);
    for (int w = 0; w < W; ++w) {
        plan_walker_size(c, params + (size_t)w * ndim, plans[w]);
        if (out_status) out_status[w] = plans[w].status;
        if (plans[w].status == 0) good.push_back(w);
    }
    const int Wg = (int)good.size();
    hp_mark("plan_sizes");

    // 2. carve the walker-independent part of the workspace
    Bump ws;
    ws.base = (char*)workspace;
    ws.cap = ws_bytes;
    double* scratch = ws.take<double>(scratch_doubles(c));
    double* d_out_ll = out_on_device ? out_ll : ws.take<double>(W);
    double* d_out_root = out_root_ll ? ws.take<double>((size_t)W * c->n_ped) : nullptr;
    double* d_out_asc = out_log_asc ? ws.take<double>((size_t)W * c->asc_total) : nullptr;
    if (!ws.ok()) return fail(MOPE_E_NOMEM, "workspace too small: %zu bytes, need more than %zu", ws_bytes, ws.used);
    const size_t fixed_used = ws.used;
    CUtensorMap tmapS;
    if (!c->use_v4)
        if (int rc = make_tmap(&tmapS, scratch, (uint64_t)c->num_sms * c->max_slots * TM, Fp, TM)) return rc;

    // outputs of invalid walkers: NaN (the host wrapper raises ValueError like the reference); valid walkers overwrite
    fill_nan_kernel<<<(W + 255) / 256, 256, 0, st>>>(d_out_ll, W);
    CUDA_TRY(cudaGetLastError());
    c->launches++;

    c->last_gemm_ops = 0;
    c->last_identity_ops = 0;
    CUDA_TRY(cudaMemsetAsync(c->work_ctr + 8, 0, 3 * sizeof(unsigned long long), st));   // executed row x unit counter + segment counters of this eval
    c->last_stream = st;
    hp_mark("fill_nan");
    // 3. chunk the valid walkers
    size_t gi0 = 0;
    // staging size upper bound for one chunk is computed per chunk
    while (gi0 < (size_t)Wg) {
        size_t avail = ws_bytes - align_up(fixed_used, 256);
        size_t used = 16 * 256;  // alignment slack of the per-chunk carve
        size_t n_blocks = 0;
        size_t gi1 = gi0;
        while (gi1 < (size_t)Wg) {
            const size_t nb = plans[good[gi1]].n_blocks;
            const size_t add = per_walker_bytes(c, nb);
            if (used + add > avail) break;
            used += add;
            n_blocks += nb;
            ++gi1;
        }
        if (gi1 == gi0) return fail(MOPE_E_NOMEM, "workspace too small for a single walker: %zu bytes available, %zu needed",
                                    avail, used + per_walker_bytes(c, plans[good[gi0]].n_blocks));
        const int Wc = (int)(gi1 - gi0);

        PlanOut po;
        po.jobs.reserve(n_blocks);
        po.mat_off.reserve((size_t)Wc * n_mat);
        po.rate.reserve((size_t)Wc * n_rate);
        size_t b0 = 0;
        std::vector<int32_t> widx(Wc);
        for (int i = 0; i < Wc; ++i) {
            const int w = good[gi0 + i];
            widx[i] = w;
            plan_walker_jobs(c, params + (size_t)w * ndim, b0, po);
            b0 += plans[w].n_blocks;
        }
        if (po.jobs.size() != n_blocks) return fail(MOPE_E_INVALID, "internal: plan size mismatch (%zu vs %zu)", po.jobs.size(), n_blocks);
        hp_mark("plan_jobs");

        // carve the chunk part
        Bump cw = ws;
        cw.used = fixed_used;
        const size_t blk = blk_doubles(Fp);
        InterpJob* d_jobs = cw.take<InterpJob>(std::max<size_t>(n_blocks, 1));
        int64_t* d_mat_off = cw.take<int64_t>(std::max<size_t>((size_t)Wc * n_mat, 1));
        RateInfo* d_rate = cw.take<RateInfo>(std::max<size_t>((size_t)Wc * n_rate, 1));
        int32_t* d_widx = cw.take<int32_t>(Wc);
        double* d_stat = cw.take<double>((size_t)Wc * Fp);
        double* d_pool = cw.take<double>(std::max<size_t>(n_blocks, 1) * blk);
        double* d_tile_ll = cw.take<double>((size_t)Wc * c->max_tiles);
        double* d_asc_dot = cw.take<double>((size_t)Wc * std::max(c->max_asc2, 2));
        double* d_uni = cw.take<double>(std::max<size_t>((size_t)Wc * c->max_uroot * 2 * Fp, 1));
        if (!cw.ok()) return fail(MOPE_E_NOMEM, "internal: chunk carve exceeded the workspace (%zu > %zu)", cw.used, cw.cap);

        // stage host -> device
        const size_t sz_jobs = n_blocks * sizeof(InterpJob), sz_mo = (size_t)Wc * n_mat * sizeof(int64_t),
                     sz_rate = (size_t)Wc * n_rate * sizeof(RateInfo), sz_widx = (size_t)Wc * sizeof(int32_t),
                     sz_stat = (size_t)Wc * Fp * sizeof(double);
        const size_t stage_bytes = align_up(sz_jobs, 64) + align_up(sz_mo, 64) + align_up(sz_rate, 64) + align_up(sz_widx, 64) +
                                   (stat_on_device ? 0 : align_up(sz_stat, 64));
        char* hp = nullptr;
        if (int rc = stage_alloc(c, stage_bytes, st, &hp)) return rc;
        hp_mark("stage_alloc");
        auto push = [&](void* dptr, const void* src, size_t bytes) -> int {
            if (bytes == 0) return 0;
            std::memcpy(hp, src, bytes);
            CUDA_TRY(cudaMemcpyAsync(dptr, hp, bytes, cudaMemcpyHostToDevice, st));
            hp += align_up(bytes, 64);
            return 0;
        };
        if (push(d_jobs, po.jobs.data(), sz_jobs)) return MOPE_E_CUDA;
        if (push(d_mat_off, po.mat_off.data(), sz_mo)) return MOPE_E_CUDA;
        if (push(d_rate, po.rate.data(), sz_rate)) return MOPE_E_CUDA;
        if (push(d_widx, widx.data(), sz_widx)) return MOPE_E_CUDA;
        if (!stat_on_device) {
            double* hs = reinterpret_cast<double*>(hp);
            for (int i = 0; i < Wc; ++i) {
                const double* src = stat + (size_t)widx[i] * F;
                std::memcpy(hs + (size_t)i * Fp, src, (size_t)F * sizeof(double));
                for (int f = F; f < Fp; ++f) hs[(size_t)i * Fp + f] = 0.0;
            }
            CUDA_TRY(cudaMemcpyAsync(d_stat, hs, sz_stat, cudaMemcpyHostToDevice, st));
        } else {
            // gather + pad rows on the device
            const long long total = (long long)Wc * Fp;
            gather_pad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(stat, d_widx, d_stat, Wc, F, Fp);
            CUDA_TRY(cudaGetLastError());
            c->launches++;
        }

        // interpolation
        if (n_blocks > 0) {
            ScopedTimer timer(c, st, &c->ev_interp);
            if (int rc = launch_interp(c, d_jobs, n_blocks, d_pool, st, c->use_v4)) return rc;
        }

        CUtensorMap tmapP;
        if (int rc = make_tmap(&tmapP, d_pool, (uint64_t)std::max<size_t>(n_blocks, 1) * (Fp + BLK_EXTRA), Fp, tree_prows(c))) return rc;

        // pruning + finalize per pedigree
        for (int k = 0; k < c->n_ped; ++k) {
            const Pedigree& P = c->peds[k];
            for (const TreeOp& op : P.ops_h) {
                if (op.kind != 0) { c->last_gemm_ops += Wc; continue; }
                for (int i = 0; i < Wc; ++i) {
                    if (po.mat_off[(size_t)i * n_mat + op.key] < 0) c->last_identity_ops++;
                    else c->last_gemm_ops++;
                }
            }
            hp_mark("copies+interp");
            const int tile_block_default = std::max(1, std::min(P.n_tiles, c->num_sms / 4));
            const long long items = (long long)Wc * P.n_tiles;
            const int grid = (int)std::min<long long>(items, c->num_sms);
            if (c->use_v4) {
                Tree4Params t4{};
                t4.tmapE = P.tmapE; t4.tmapP = tmapP;
                t4.E = P.E; t4.E_leaf_stride = (int64_t)P.R_pad * Fp; t4.emask = getenv("MOPE_NO_EMASK") ? nullptr : P.emask;
                t4.scratch = scratch;
                t4.ops = P.ops4; t4.n_ops = (int)P.ops4_h.size(); t4.n_slots = c->max_slots4;
                t4.R = P.R; t4.L = P.L; t4.R_pad = P.R_pad; t4.Fp = Fp; t4.F = F; t4.NT = c->NT; t4.n_tiles = P.n_tiles; t4.W = Wc;
                t4.tile_block = tile_block_default;
                if (const char* e = getenv("MOPE_TILE_BLOCK")) t4.tile_block = std::max(1, std::min(P.n_tiles, atoi(e)));
                t4.pool = d_pool; t4.mat_off = d_mat_off; t4.n_mat = n_mat; t4.n_rate = n_rate; t4.rate = d_rate;
                t4.mult = P.mult; t4.gens = c->gens_dev; t4.n_gens = c->n_gens; t4.Npop = c->N; t4.stat = d_stat;
                t4.tile_ll = d_tile_ll; t4.asc_dot = d_asc_dot; t4.Yout = nullptr;
                if (const char* e = getenv("MOPE_DEBUG_MODE")) t4.debug = atoi(e);
                // family-independent messages of the ascertainment pass: one canonical tile per walker first, when the
                // pedigree has enough pure pseudo-row tiles to pay for it
                const int pure_asc_tiles = (P.R_pad - (P.L + TM - 1) / TM * TM) / TM;
                if (P.n_uroot4 > 0 && pure_asc_tiles >= 8 && !getenv("MOPE_NO_ASC_FAST") && !getenv("MOPE_NO_ASC_CSE")) {
                    t4.uni_msg = d_uni; t4.n_uroot = P.n_uroot4;
                    Tree4Params tc = t4;
                    tc.canonical = 1;
                    if (launch_tree4(c, tc, std::min(Wc, c->num_sms), st)) return MOPE_E_CUDA;
                }
                if (launch_tree4(c, t4, grid, st)) return MOPE_E_CUDA;
            } else {
            TreeParams tp{};
            tp.tmapE = P.tmapE;
            tp.tmapS = tmapS;
            tp.tmapP = tmapP;
            tp.E = P.E; tp.emask = (getenv("MOPE_NO_EMASK") || c->Fp / KC > 64) ? nullptr : P.emask;
            tp.E_leaf_stride = (int64_t)P.R_pad * Fp;
            tp.scratch = scratch;
            tp.ops = P.ops;
            tp.n_ops = P.n_ops;
            tp.n_slots = c->max_slots;
            tp.R = P.R; tp.L = P.L; tp.R_pad = P.R_pad; tp.Fp = Fp; tp.F = F; tp.NT = c->NT; tp.n_tiles = P.n_tiles; tp.W = Wc;
            tp.tile_block = std::max(1, std::min(P.n_tiles, c->num_sms / 4));
            if (const char* e = getenv("MOPE_TILE_BLOCK")) tp.tile_block = std::max(1, std::min(P.n_tiles, atoi(e)));  // tuning knob
            tp.pool = d_pool;
            tp.mat_off = d_mat_off;
            tp.n_mat = n_mat; tp.n_rate = n_rate;
            tp.rate = d_rate;
            tp.mult = P.mult;
            tp.gens = c->gens_dev;
            tp.n_gens = c->n_gens;
            tp.Npop = c->N;
            tp.stat = d_stat;
            tp.tile_ll = d_tile_ll;
            tp.asc_dot = d_asc_dot;
            tp.Yout = nullptr;
            if (const char* e = getenv("MOPE_DEBUG_MODE")) tp.debug = atoi(e);  // measurement only (results are garbage)
            if (launch_tree(c, tp, grid, st)) return MOPE_E_CUDA;
            }

            FinalizeParams fp{};
            fp.tile_ll = d_tile_ll; fp.asc_dot = d_asc_dot;
            fp.fam_asc = P.fam_asc; fp.fam_count = P.fam_count; fp.fam_lgamma = P.fam_lgamma;
            fp.out_ll = d_out_ll; fp.out_root_ll = d_out_root; fp.out_log_asc = d_out_asc;
            fp.widx = d_widx;
            fp.n_tiles = P.n_tiles; fp.n_asc = P.n_asc; fp.n_fam = P.n_fam; fp.W = Wc;
            fp.ped_index = k; fp.ped_stride = c->n_ped; fp.asc_offset = P.asc_offset; fp.asc_stride = c->asc_total;
            fp.accumulate = (k > 0);
            fp.log_genome = std::log(c->genome);
            fp.penalty = c->penalty;
            finalize_kernel<<<Wc, 256, 0, st>>>(fp);
            CUDA_TRY(cudaGetLastError());
            c->launches++;
            hp_mark("tree+finalize");
        }
        gi0 = gi1;
    }
    if (host_prof_ms >= 0.0 && !hp_marks.empty() && hp_marks.back().second > host_prof_ms) {
        fprintf(stderr, "[mope_b200 host] eval W=%d:", W);
        for (auto& m : hp_marks) fprintf(stderr, " %s@%.2f", m.first, m.second);
        fprintf(stderr, "\n");
    }

    if (!out_on_device) {
        // device -> host
        size_t need = (size_t)W * sizeof(double) * (1 + (out_root_ll ? c->n_ped : 0) + (out_log_asc ? c->asc_total : 0));
        char* hp = nullptr;
        if (int rc = stage_alloc(c, need, st, &hp)) return rc;
        CUDA_TRY(cudaMemcpyAsync(hp, d_out_ll, (size_t)W * sizeof(double), cudaMemcpyDeviceToHost, st));
        char* hp_root = hp + (size_t)W * sizeof(double);
        char* hp_asc = hp_root + (out_root_ll ? (size_t)W * c->n_ped * sizeof(double) : 0);
        if (out_root_ll) CUDA_TRY(cudaMemcpyAsync(hp_root, d_out_root, (size_t)W * c->n_ped * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (out_log_asc) CUDA_TRY(cudaMemcpyAsync(hp_asc, d_out_asc, (size_t)W * c->asc_total * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        std::memcpy(out_ll, hp, (size_t)W * sizeof(double));
        if (out_root_ll) std::memcpy(out_root_ll, hp_root, (size_t)W * c->n_ped * sizeof(double));
        if (out_log_asc) std::memcpy(out_log_asc, hp_asc, (size_t)W * c->asc_total * sizeof(double));
        const double qnan = std::numeric_limits<double>::quiet_NaN();
        for (int w = 0; w < W; ++w) {
            if (plans[w].status) {
                out_ll[w] = qnan;
                if (out_root_ll) for (int k = 0; k < c->n_ped; ++k) out_root_ll[(size_t)w * c->n_ped + k] = qnan;
                if (out_log_asc) for (int u = 0; u < c->asc_total; ++u) out_log_asc[(size_t)w * c->asc_total + u] = qnan;
            }
        }
    }
    return 0;
}

// Device-resident evaluation (mope_b200_eval_resident): nothing on the host but launch parameters.
int eval_resident_impl(mope_ctx* c, const double* params_dev, const double* stat_dev, const double* lnprior_dev, int W,
                       void* workspace, size_t ws_bytes, cudaStream_t st, double* out_ll_dev, int32_t* out_status_dev) {
    if (!c || !params_dev || !out_ll_dev || !out_status_dev || W < 0) return fail(MOPE_E_INVALID, "eval_resident: null argument");
    if (W == 0) return 0;
    if (!workspace) return fail(MOPE_E_INVALID, "eval_resident: null workspace");
    if (!stat_dev && !c->has_breaks) return fail(MOPE_E_INVALID, "eval_resident: no root distributions given and the context has no `breaks` to compute them");
    CUDA_TRY(cudaSetDevice(c->device));
    const int Fp = c->Fp, F = c->F;
    const int n_mat = (int)c->fixed_keys.size(), n_rate = (int)c->rate_keys.size();
    const size_t max_blocks = std::max<size_t>(max_blocks_per_walker(c), 1);
    const size_t blk = blk_doubles(Fp);

    Bump ws;
    ws.base = (char*)workspace;
    ws.cap = ws_bytes;
    double* scratch = ws.take<double>(scratch_doubles(c));
    if (!ws.ok()) return fail(MOPE_E_NOMEM, "workspace too small: %zu bytes, need more than %zu", ws_bytes, ws.used);
    const size_t fixed_used = ws.used;
    CUtensorMap tmapS;
    if (!c->use_v4)
        if (int rc = make_tmap(&tmapS, scratch, (uint64_t)c->num_sms * c->max_slots * TM, Fp, TM)) return rc;
    const size_t pw = per_walker_bytes(c, max_blocks);
    const size_t avail = ws_bytes - align_up(fixed_used, 256);
    const size_t slack = 16 * 256;
    if (avail < slack + pw) return fail(MOPE_E_NOMEM, "workspace too small for a single walker: %zu bytes available, %zu needed", avail, slack + pw);
    const int Wchunk = (int)std::min<size_t>((size_t)W, (avail - slack) / pw);

    CUDA_TRY(cudaMemsetAsync(out_status_dev, 0, (size_t)W * sizeof(int32_t), st));
    CUDA_TRY(cudaMemsetAsync(c->work_ctr + 8, 0, 3 * sizeof(unsigned long long), st));
    c->last_stream = st;
    c->last_gemm_ops = -1;
    c->last_identity_ops = -1;

    for (int w0 = 0; w0 < W; w0 += Wchunk) {
        const int Wc = std::min(Wchunk, W - w0);
        const size_t n_blocks = (size_t)Wc * max_blocks;
        Bump cw = ws;
        cw.used = fixed_used;
        InterpJob* d_jobs = cw.take<InterpJob>(n_blocks);
        int64_t* d_mat_off = cw.take<int64_t>(std::max<size_t>((size_t)Wc * n_mat, 1));
        RateInfo* d_rate = cw.take<RateInfo>(std::max<size_t>((size_t)Wc * n_rate, 1));
        cw.take<int32_t>(Wc);   // (widx of the host-planned path; keeps the carve identical)
        double* d_stat = cw.take<double>((size_t)Wc * Fp);
        double* d_pool = cw.take<double>(n_blocks * blk);
        double* d_tile_ll = cw.take<double>((size_t)Wc * c->max_tiles);
        double* d_asc_dot = cw.take<double>((size_t)Wc * std::max(c->max_asc2, 2));
        double* d_uni = cw.take<double>(std::max<size_t>((size_t)Wc * c->max_uroot * 2 * Fp, 1));
        if (!cw.ok()) return fail(MOPE_E_NOMEM, "internal: chunk carve exceeded the workspace (%zu > %zu)", cw.used, cw.cap);
        int32_t* d_status = out_status_dev + w0;

        // plans
        PlanParams pp{};
        pp.params = params_dev + (size_t)w0 * c->ndim;
        pp.lnprior = lnprior_dev ? lnprior_dev + w0 : nullptr;
        pp.ndim = c->ndim; pp.n_var = c->n_var; pp.W = Wc;
        pp.gens = c->gens_dev; pp.us = c->us_dev; pp.nbs = c->nbs_dev; pp.bot_us = c->bot_us_dev;
        pp.n_gens = c->n_gens; pp.n_us = c->n_us; pp.n_nbs = c->n_nbs; pp.n_bot_us = c->n_bot_us; pp.has_bot = c->has_bot ? 1 : 0;
        pp.N = c->N; pp.max_coal = c->max_coal; pp.min_mut = c->min_mut; pp.max_mut = c->max_mut;
        pp.min_bmut = c->min_bmut; pp.max_bmut = c->max_bmut; pp.min_nb = c->min_nb; pp.max_nb = c->max_nb;
        pp.fkeys = c->fkeys_dev; pp.n_fixed = n_mat; pp.rkeys = c->rkeys_dev; pp.n_rate = n_rate;
        pp.bot_base = c->bot_base; pp.F = F; pp.Fp = Fp; pp.max_blocks = (int)max_blocks;
        pp.jobs = d_jobs; pp.mat_off = d_mat_off; pp.rate = d_rate; pp.status = d_status;
        const long long nthreads = (long long)Wc * std::max(n_mat + n_rate, 1);
        plan_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(pp);
        CUDA_TRY(cudaGetLastError());
        c->launches++;

        // root distributions
        if (stat_dev) {
            const long long total = (long long)Wc * Fp;
            gather_pad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(stat_dev, nullptr, d_stat, Wc, F, Fp, w0);
        } else {
            const int Nwf = (int)c->N;
            root_prior_kernel<<<Wc, 256, (size_t)(Nwf + 1) * sizeof(double), st>>>(params_dev + (size_t)w0 * c->ndim, c->ndim, Wc, c->breaks_dev,
                                                                                   F, Nwf, d_stat, Fp);
        }
        CUDA_TRY(cudaGetLastError());
        c->launches++;

        {
            ScopedTimer timer(c, st, &c->ev_interp);
            if (int rc = launch_interp(c, d_jobs, n_blocks, d_pool, st, c->use_v4)) return rc;
        }
        CUtensorMap tmapP;
        if (int rc = make_tmap(&tmapP, d_pool, (uint64_t)n_blocks * (Fp + BLK_EXTRA), Fp, tree_prows(c))) return rc;

        for (int k = 0; k < c->n_ped; ++k) {
            const Pedigree& P = c->peds[k];
            const int tile_block_default = std::max(1, std::min(P.n_tiles, c->num_sms / 4));
            const long long items = (long long)Wc * P.n_tiles;
            const int grid = (int)std::min<long long>(items, c->num_sms);
            if (c->use_v4) {
                Tree4Params t4{};
                t4.tmapE = P.tmapE; t4.tmapP = tmapP;
                t4.E = P.E; t4.E_leaf_stride = (int64_t)P.R_pad * Fp; t4.emask = getenv("MOPE_NO_EMASK") ? nullptr : P.emask;
                t4.scratch = scratch;
                t4.ops = P.ops4; t4.n_ops = (int)P.ops4_h.size(); t4.n_slots = c->max_slots4;
                t4.R = P.R; t4.L = P.L; t4.R_pad = P.R_pad; t4.Fp = Fp; t4.F = F; t4.NT = c->NT; t4.n_tiles = P.n_tiles; t4.W = Wc;
                t4.tile_block = tile_block_default;
                t4.pool = d_pool; t4.mat_off = d_mat_off; t4.n_mat = n_mat; t4.n_rate = n_rate; t4.rate = d_rate;
                t4.mult = P.mult; t4.gens = c->gens_dev; t4.n_gens = c->n_gens; t4.Npop = c->N; t4.stat = d_stat;
                t4.tile_ll = d_tile_ll; t4.asc_dot = d_asc_dot; t4.Yout = nullptr;
                t4.wstatus = d_status;
                const int pure_asc_tiles = (P.R_pad - (P.L + TM - 1) / TM * TM) / TM;
                if (P.n_uroot4 > 0 && pure_asc_tiles >= 8 && !getenv("MOPE_NO_ASC_FAST") && !getenv("MOPE_NO_ASC_CSE")) {
                    t4.uni_msg = d_uni; t4.n_uroot = P.n_uroot4;
                    Tree4Params tc = t4;
                    tc.canonical = 1;
                    if (launch_tree4(c, tc, std::min(Wc, c->num_sms), st)) return MOPE_E_CUDA;
                }
                if (launch_tree4(c, t4, grid, st)) return MOPE_E_CUDA;
            } else {
                TreeParams tp{};
                tp.tmapE = P.tmapE; tp.tmapS = tmapS; tp.tmapP = tmapP;
                tp.E = P.E; tp.E_leaf_stride = (int64_t)P.R_pad * Fp; tp.scratch = scratch;
                tp.emask = (getenv("MOPE_NO_EMASK") || c->Fp / KC > 64) ? nullptr : P.emask;
                tp.ops = P.ops; tp.n_ops = P.n_ops; tp.n_slots = c->max_slots;
                tp.R = P.R; tp.L = P.L; tp.R_pad = P.R_pad; tp.Fp = Fp; tp.F = F; tp.NT = c->NT; tp.n_tiles = P.n_tiles; tp.W = Wc;
                tp.tile_block = tile_block_default;
                tp.pool = d_pool; tp.mat_off = d_mat_off; tp.n_mat = n_mat; tp.n_rate = n_rate; tp.rate = d_rate;
                tp.mult = P.mult; tp.gens = c->gens_dev; tp.n_gens = c->n_gens; tp.Npop = c->N; tp.stat = d_stat;
                tp.tile_ll = d_tile_ll; tp.asc_dot = d_asc_dot; tp.Yout = nullptr;
                tp.wstatus = d_status;
                if (launch_tree(c, tp, grid, st)) return MOPE_E_CUDA;
            }
            FinalizeParams fp{};
            fp.tile_ll = d_tile_ll; fp.asc_dot = d_asc_dot;
            fp.fam_asc = P.fam_asc; fp.fam_count = P.fam_count; fp.fam_lgamma = P.fam_lgamma;
            fp.out_ll = out_ll_dev; fp.out_root_ll = nullptr; fp.out_log_asc = nullptr;
            fp.widx = nullptr; fp.w_base = w0; fp.wstatus = d_status;
            fp.n_tiles = P.n_tiles; fp.n_asc = P.n_asc; fp.n_fam = P.n_fam; fp.W = Wc;
            fp.ped_index = k; fp.ped_stride = c->n_ped; fp.asc_offset = P.asc_offset; fp.asc_stride = c->asc_total;
            fp.accumulate = (k > 0);
            fp.log_genome = std::log(c->genome);
            fp.penalty = c->penalty;
            finalize_kernel<<<Wc, 256, 0, st>>>(fp);
            CUDA_TRY(cudaGetLastError());
            c->launches++;
        }
    }
    return 0;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char* mope_b200_last_error(void) { return g_err.c_str(); }
const char* mope_b200_version(void) { return "mope_b200 0.1 (sm_100a, fp64 DMMA)"; }

int mope_b200_arena_bytes(const mope_tables_desc* tables, const mope_model_desc* model, size_t* out_bytes) {
    if (!out_bytes) return fail(MOPE_E_INVALID, "null out_bytes");
    if (int rc = validate(tables, model)) return rc;
    Bump b;
    b.dry = true;
    layout_arena(nullptr, tables, model, b, true);
    *out_bytes = align_up(b.used, 256) + 4096;
    return 0;
}

int mope_b200_create(const mope_tables_desc* t, const mope_model_desc* m, int device, void* arena, size_t arena_bytes,
                     void* stream, mope_ctx** out) {
    if (!out) return fail(MOPE_E_INVALID, "null out");
    *out = nullptr;
    if (int rc = validate(t, m)) return rc;
    if (!arena) return fail(MOPE_E_INVALID, "null arena");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device || device < 0)
        return fail(MOPE_E_NODEVICE, "CUDA device %d not available (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(MOPE_E_NODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    cudaStream_t st = (cudaStream_t)stream;

    mope_ctx* c = new mope_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    c->F = t->F;
    c->Fp = round_up_i(t->F, 16);
    c->NT = (t->F + 7) / 8;
    c->tree_variant = c->Fp <= 80 ? 5 : (c->Fp <= 128 ? 8 : 13);
    // F <= 208: warp-owned rows (tree4.cuh); larger grids need several N passes per branch and use tree_kernel
    c->use_v4 = (c->Fp <= 208) && !getenv("MOPE_FORCE_V3");
    c->ntt = c->NT <= 10 ? 10 : (c->NT <= 16 ? 16 : 26);
    c->n_gens = t->n_gens; c->n_us = t->n_us;
    c->has_bot = t->bot != nullptr;
    c->n_nbs = c->has_bot ? t->n_nbs : 0;
    c->n_bot_us = c->has_bot ? t->n_bot_us : 0;
    c->N = t->N;
    c->gens.assign(t->gens, t->gens + t->n_gens);
    c->us.assign(t->us, t->us + t->n_us);
    c->freqs.assign(t->freqs, t->freqs + t->F);
    if (c->has_bot) { c->nbs.assign(t->nbs, t->nbs + t->n_nbs); c->bot_us.assign(t->bot_us, t->bot_us + t->n_bot_us); }
    // transition_data_2d.py:105-108, transition_data_bottleneck.py:90-93
    c->min_coal = c->gens.front() / c->N; c->max_coal = c->gens.back() / c->N;
    c->min_mut = c->us.front() * 2 * c->N; c->max_mut = c->us.back() * 2 * c->N;
    if (c->has_bot) {
        c->min_nb = c->nbs.front(); c->max_nb = c->nbs.back();
        c->min_bmut = c->bot_us.front() * 2 * c->N; c->max_bmut = c->bot_us.back() * 2 * c->N;
    }
    c->n_var = m->n_var; c->ndim = 2 * m->n_var + 2; c->n_ped = m->n_ped;
    c->min_phred = m->min_phred_score; c->min_freq = m->min_freq; c->genome = m->genome_size; c->penalty = m->poisson_like_penalty;
    c->peds.resize(m->n_ped);

    // matrix keys shared by all pedigrees
    std::vector<std::vector<int>> fk(m->n_ped), rk(m->n_ped);
    for (int k = 0; k < m->n_ped; ++k) {
        const mope_pedigree_desc& pd = m->peds[k];
        fk[k].assign(pd.n_branches, -1);
        rk[k].assign(pd.n_branches, -1);
        for (int b = 0; b < pd.n_branches; ++b) {
            const int v = pd.br_var[b];
            if (pd.br_mult[b] < 0) {
                const int bott = pd.br_bottleneck[b] ? 1 : 0;
                int idx = -1;
                for (size_t i = 0; i < c->fixed_keys.size(); ++i) if (c->fixed_keys[i].var == v && c->fixed_keys[i].bott == bott) idx = (int)i;
                if (idx < 0) { c->fixed_keys.push_back({v, bott}); idx = (int)c->fixed_keys.size() - 1; }
                fk[k][b] = idx;
            } else {
                double mn = std::numeric_limits<double>::infinity(), mx = -mn;
                const int mc = pd.br_mult[b];
                for (int r = 0; r < pd.n_loci; ++r) { double v2 = pd.row_mult[(size_t)mc * pd.n_loci + r]; mn = std::min(mn, v2); mx = std::max(mx, v2); }
                for (int u = 0; u < pd.n_asc; ++u) { double v2 = pd.asc_mult[(size_t)mc * pd.n_asc + u]; mn = std::min(mn, v2); mx = std::max(mx, v2); }
                int idx = -1;
                for (size_t i = 0; i < c->rate_keys.size(); ++i) if (c->rate_keys[i].var == v) idx = (int)i;
                if (idx < 0) { c->rate_keys.push_back({v, mn, mx}); idx = (int)c->rate_keys.size() - 1; }
                else { c->rate_keys[idx].min_mult = std::min(c->rate_keys[idx].min_mult, mn); c->rate_keys[idx].max_mult = std::max(c->rate_keys[idx].max_mult, mx); }
                rk[k][b] = idx;
            }
        }
    }

    Bump bump;
    bump.base = (char*)arena;
    bump.cap = arena_bytes;
    layout_arena(c, t, m, bump, false);
    if (!bump.ok()) { delete c; return fail(MOPE_E_NOMEM, "arena too small: %zu bytes, need %zu", arena_bytes, bump.used); }

    auto cleanup_fail = [&](int rc) { delete c; return rc; };
#define CUDA_TRY_C(expr)                                                                                        \
    do {                                                                                                        \
        cudaError_t e__ = (expr);                                                                               \
        if (e__ != cudaSuccess)                                                                                 \
            return cleanup_fail(fail(MOPE_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__)); \
    } while (0)

    // tables
    const size_t FF = (size_t)t->F * t->F;
    CUDA_TRY_C(cudaMemcpyAsync(c->tables, t->drift, (size_t)t->n_gens * t->n_us * FF * sizeof(double), cudaMemcpyHostToDevice, st));
    if (c->has_bot)
        CUDA_TRY_C(cudaMemcpyAsync(c->tables + c->bot_base, t->bot, (size_t)t->n_nbs * t->n_bot_us * FF * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY_C(cudaMemcpyAsync(c->gens_dev, t->gens, (size_t)t->n_gens * sizeof(double), cudaMemcpyHostToDevice, st));
    if (t->breaks) {
        if (t->breaks[0] != 0) return cleanup_fail(fail(MOPE_E_INVALID, "breaks must start at state 0"));
        for (int f = 1; f < t->F; ++f)
            if (t->breaks[f] <= t->breaks[f - 1] || t->breaks[f] > (int)t->N)
                return cleanup_fail(fail(MOPE_E_INVALID, "breaks must be strictly ascending states in 0..N"));
        CUDA_TRY_C(cudaMemcpyAsync(c->breaks_dev, t->breaks, (size_t)t->F * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        c->has_breaks = true;
    }

    // per-frequency emission constants on the host with libm (mope/_binom.pyx:76-91)
    const int F = t->F, Fp = c->Fp;
    std::vector<double> logp(F), log1mp(F);
    std::vector<int8_t> pclass(F), e0(F), eN(F);
    const double perr = (m->min_phred_score != -1.0) ? std::pow(10.0, -1.0 * m->min_phred_score / 10.0) : 0.0;
    for (int f = 0; f < F; ++f) {
        const double fr = t->freqs[f];
        const double p = (1.0 - perr) * fr + (perr / 3.0) * (1.0 - fr);
        pclass[f] = (p == 0.0) ? 1 : ((p == 1.0) ? 2 : 0);
        logp[f] = (p > 0.0) ? std::log(p) : 0.0;
        log1mp[f] = (p < 1.0) ? std::log1p(-p) : 0.0;
        e0[f] = fr < m->min_freq;         // ascertainment.py:178-183
        eN[f] = fr > 1 - m->min_freq;
    }
    CUDA_TRY_C(cudaMemcpyAsync(c->asc_mask, e0.data(), F, cudaMemcpyHostToDevice, st));
    CUDA_TRY_C(cudaMemcpyAsync(c->asc_mask + F, eN.data(), F, cudaMemcpyHostToDevice, st));
    std::vector<unsigned> lane_masks_h(64, 0u);
    if (c->lane_masks) {
        for (int f = 0; f < F; ++f) {
            if (e0[f]) lane_masks_h[f & 31] |= 1u << (f >> 5);
            if (eN[f]) lane_masks_h[32 + (f & 31)] |= 1u << (f >> 5);
        }
        CUDA_TRY_C(cudaMemcpyAsync(c->lane_masks, lane_masks_h.data(), 64 * sizeof(unsigned), cudaMemcpyHostToDevice, st));
    }

    int asc_off = 0;
    for (int k = 0; k < m->n_ped; ++k) {
        const mope_pedigree_desc& pd = m->peds[k];
        Pedigree& P = c->peds[k];
        P.n_leaves = pd.n_leaves; P.n_branches = pd.n_branches; P.L = pd.n_loci; P.n_fam = pd.n_fam; P.n_mult = pd.n_mult; P.n_asc = pd.n_asc;
        P.R = pd.n_loci + 2 * pd.n_asc;
        P.R_pad = round_up_i(std::max(P.R, 1), TM);
        P.n_tiles = P.R_pad / TM;
        P.br_parent_h.assign(pd.br_parent, pd.br_parent + pd.n_branches);
        P.br_leaf_h.assign(pd.br_leaf, pd.br_leaf + pd.n_branches);
        P.br_mult_h.assign(pd.br_mult, pd.br_mult + pd.n_branches);
        if (pd.n_mult > 0 && pd.n_loci > 0) P.row_mult_h.assign(pd.row_mult, pd.row_mult + (size_t)pd.n_mult * pd.n_loci);
        P.asc_offset = asc_off;
        asc_off += pd.n_asc;
        if (int rc = compile_schedule(pd, fk[k], rk[k], P)) return cleanup_fail(rc);
        c->max_slots = std::max(c->max_slots, P.n_slots);
        if (c->use_v4) {
            if (int rc = compile_schedule4(pd, fk[k], rk[k], 256 / (c->ntt * 4) - 1, P)) return cleanup_fail(rc);
            c->max_slots4 = std::max(c->max_slots4, P.n_slots4);
        }
        c->max_tiles = std::max(c->max_tiles, P.n_tiles);
        c->max_asc2 = std::max(c->max_asc2, 2 * P.n_asc);
        c->max_uroot = std::max(c->max_uroot, P.n_uroot4);

        // multipliers per row (real rows then ascertainment pairs, np.repeat(..., 2): ascertainment.py:205-206)
        std::vector<double> mult((size_t)std::max(pd.n_mult, 1) * P.R_pad, 0.0);
        for (int mc = 0; mc < pd.n_mult; ++mc) {
            for (int r = 0; r < pd.n_loci; ++r) mult[(size_t)mc * P.R_pad + r] = pd.row_mult[(size_t)mc * pd.n_loci + r];
            for (int u = 0; u < pd.n_asc; ++u) {
                mult[(size_t)mc * P.R_pad + pd.n_loci + 2 * u] = pd.asc_mult[(size_t)mc * pd.n_asc + u];
                mult[(size_t)mc * P.R_pad + pd.n_loci + 2 * u + 1] = pd.asc_mult[(size_t)mc * pd.n_asc + u];
            }
        }
        CUDA_TRY_C(cudaMemcpyAsync(P.mult, mult.data(), mult.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        CUDA_TRY_C(cudaMemcpyAsync(P.fam_asc, pd.fam_asc, (size_t)pd.n_fam * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CUDA_TRY_C(cudaMemcpyAsync(P.fam_count, pd.fam_count, (size_t)pd.n_fam * sizeof(double), cudaMemcpyHostToDevice, st));
        CUDA_TRY_C(cudaMemcpyAsync(P.fam_lgamma, pd.fam_lgamma, (size_t)pd.n_fam * sizeof(double), cudaMemcpyHostToDevice, st));
        CUDA_TRY_C(cudaMemcpyAsync(P.ops, P.ops_h.data(), (size_t)P.n_ops * sizeof(TreeOp), cudaMemcpyHostToDevice, st));
        if (c->use_v4)
            CUDA_TRY_C(cudaMemcpyAsync(P.ops4, P.ops4_h.data(), P.ops4_h.size() * sizeof(Op4), cudaMemcpyHostToDevice, st));

        // emissions: transient inputs at the arena tail
        Bump tail = bump;
        const size_t cells = (size_t)pd.n_loci * pd.n_leaves;
        double* d_counts = tail.take<double>(std::max<size_t>(cells, 1));
        double* d_cov = tail.take<double>(std::max<size_t>(cells, 1));
        double* d_lnc = tail.take<double>(std::max<size_t>(cells, 1));
        double* d_logp = tail.take<double>(F);
        double* d_log1mp = tail.take<double>(F);
        int8_t* d_cls = tail.take<int8_t>(F);
        int8_t* d_e0 = tail.take<int8_t>(F);
        int8_t* d_eN = tail.take<int8_t>(F);
        if (!tail.ok()) return cleanup_fail(fail(MOPE_E_NOMEM, "arena too small for the emission inputs"));
        std::vector<double> lnc(cells);
        if (cells) {
            host_lnchoose(pd.counts, pd.coverage, cells, lnc.data());
            CUDA_TRY_C(cudaMemcpyAsync(d_counts, pd.counts, cells * sizeof(double), cudaMemcpyHostToDevice, st));
            CUDA_TRY_C(cudaMemcpyAsync(d_cov, pd.coverage, cells * sizeof(double), cudaMemcpyHostToDevice, st));
            CUDA_TRY_C(cudaMemcpyAsync(d_lnc, lnc.data(), cells * sizeof(double), cudaMemcpyHostToDevice, st));
        }
        CUDA_TRY_C(cudaMemcpyAsync(d_logp, logp.data(), F * sizeof(double), cudaMemcpyHostToDevice, st));
        CUDA_TRY_C(cudaMemcpyAsync(d_log1mp, log1mp.data(), F * sizeof(double), cudaMemcpyHostToDevice, st));
        CUDA_TRY_C(cudaMemcpyAsync(d_cls, pclass.data(), F, cudaMemcpyHostToDevice, st));
        CUDA_TRY_C(cudaMemcpyAsync(d_e0, e0.data(), F, cudaMemcpyHostToDevice, st));
        CUDA_TRY_C(cudaMemcpyAsync(d_eN, eN.data(), F, cudaMemcpyHostToDevice, st));
        EmissionParams ep{};
        ep.counts = d_counts; ep.coverage = d_cov; ep.lnchoose = d_lnc; ep.logp = d_logp; ep.log1mp = d_log1mp; ep.pclass = d_cls;
        ep.asc_e0 = d_e0; ep.asc_eN = d_eN; ep.E = P.E;
        ep.L = P.L; ep.S = P.n_leaves; ep.R = P.R; ep.R_pad = P.R_pad; ep.F = F; ep.Fp = Fp;
        ep.perm = c->use_v4 ? 1 : 0;
        const long long warps = (long long)P.n_leaves * P.R_pad;
        const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
        emission_kernel<<<grid, 256, 0, st>>>(ep);
        CUDA_TRY_C(cudaGetLastError());
        c->launches++;
        {
            const long long mwarps = (long long)P.n_leaves * (P.R_pad / 8);
            emission_mask_kernel<<<(unsigned)((mwarps * 32 + 255) / 256), 256, 0, st>>>(P.E, P.emask, P.n_leaves, P.R_pad, Fp);
            CUDA_TRY_C(cudaGetLastError());
            c->launches++;
        }
        CUDA_TRY_C(cudaStreamSynchronize(st));  // host vectors above go out of scope per pedigree
        if (int rc = make_tmap(&P.tmapE, P.E, (uint64_t)P.n_leaves * P.R_pad, Fp, TM)) return cleanup_fail(rc);
    }
    c->asc_total = asc_off;
    {   // constants of the device-side planner
        std::vector<FixedKeyDev> fk_h;
        for (const FixedKey& k : c->fixed_keys) fk_h.push_back({k.var, k.bott});
        std::vector<RateKeyDev> rk_h;
        for (const RateKey& k : c->rate_keys) rk_h.push_back({k.var, 0, k.min_mult, k.max_mult});
        Bump pb;
        pb.dry = true;
        auto carve = [&](Bump& b) {
            c->us_dev = b.take<double>(c->n_us);
            c->nbs_dev = b.take<double>(std::max(c->n_nbs, 1));
            c->bot_us_dev = b.take<double>(std::max(c->n_bot_us, 1));
            c->fkeys_dev = b.take<FixedKeyDev>(std::max<size_t>(fk_h.size(), 1));
            c->rkeys_dev = b.take<RateKeyDev>(std::max<size_t>(rk_h.size(), 1));
        };
        carve(pb);
        CUDA_TRY_C(cudaMalloc(&c->plan_consts, pb.used + 256));
        Bump rb;
        rb.base = c->plan_consts; rb.cap = pb.used + 256;
        carve(rb);
        CUDA_TRY_C(cudaMemcpyAsync(c->us_dev, c->us.data(), c->us.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        if (c->has_bot) {
            CUDA_TRY_C(cudaMemcpyAsync(c->nbs_dev, c->nbs.data(), c->nbs.size() * sizeof(double), cudaMemcpyHostToDevice, st));
            CUDA_TRY_C(cudaMemcpyAsync(c->bot_us_dev, c->bot_us.data(), c->bot_us.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        }
        if (!fk_h.empty()) CUDA_TRY_C(cudaMemcpyAsync(c->fkeys_dev, fk_h.data(), fk_h.size() * sizeof(FixedKeyDev), cudaMemcpyHostToDevice, st));
        if (!rk_h.empty()) CUDA_TRY_C(cudaMemcpyAsync(c->rkeys_dev, rk_h.data(), rk_h.size() * sizeof(RateKeyDev), cudaMemcpyHostToDevice, st));
        CUDA_TRY_C(cudaStreamSynchronize(st));
    }
    CUDA_TRY_C(cudaStreamSynchronize(st));
#undef CUDA_TRY_C
    *out = c;
    return 0;
}

int mope_b200_destroy(mope_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->pinned) cudaFreeHost(c->pinned);
    for (int i = 0; i < 2; ++i) if (c->pin_ev[i]) cudaEventDestroy(c->pin_ev[i]);
    if (c->opbuf) cudaFree(c->opbuf);
    if (c->plan_consts) cudaFree(c->plan_consts);
    for (auto& P : c->peds) if (P.sel.dev) cudaFree(P.sel.dev);
    for (auto& pr : c->ev_tree) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (auto& pr : c->ev_interp) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    delete c;
    return 0;
}

int mope_b200_workspace_bytes(const mope_ctx* c, int W, size_t* out_min, size_t* out_full) {
    if (!c || W < 1) return fail(MOPE_E_INVALID, "workspace_bytes: bad argument");
    const size_t fixed = fixed_ws_bytes(c, W);
    const size_t pw = per_walker_bytes(c, max_blocks_per_walker(c));
    if (out_min) *out_min = fixed + pw + 8192;
    if (out_full) *out_full = fixed + (size_t)W * pw + 8192;
    return 0;
}

int mope_b200_eval(mope_ctx* c, const double* params, const double* stat, int W, void* workspace, size_t ws_bytes,
                   void* stream, double* out_ll, int32_t* out_status, double* out_root_ll, double* out_log_asc) {
    return eval_impl(c, params, stat, false, W, workspace, ws_bytes, (cudaStream_t)stream, out_ll, false, out_status, out_root_ll, out_log_asc);
}

int mope_b200_eval_device(mope_ctx* c, const double* params_host, const double* stat_dev, int W, void* workspace,
                          size_t ws_bytes, void* stream, double* out_ll_dev, int32_t* out_status_host) {
    return eval_impl(c, params_host, stat_dev, true, W, workspace, ws_bytes, (cudaStream_t)stream, out_ll_dev, true, out_status_host, nullptr, nullptr);
}

int mope_b200_root_prior(mope_ctx* c, const double* params_dev, int W, void* stream, double* out_stat_dev) {
    if (!c || !params_dev || !out_stat_dev || W < 0) return fail(MOPE_E_INVALID, "root_prior: bad argument");
    if (!c->has_breaks) return fail(MOPE_E_INVALID, "root_prior: the context was created without mope_tables_desc.breaks");
    if (W == 0) return 0;
    CUDA_TRY(cudaSetDevice(c->device));
    const int Nwf = (int)c->N;
    root_prior_kernel<<<W, 256, (size_t)(Nwf + 1) * sizeof(double), (cudaStream_t)stream>>>(params_dev, c->ndim, W, c->breaks_dev, c->F, Nwf,
                                                                                             out_stat_dev, c->F);
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    return 0;
}

int mope_b200_eval_resident(mope_ctx* c, const double* params_dev, const double* stat_dev, const double* lnprior_dev, int W,
                            void* workspace, size_t ws_bytes, void* stream, double* out_ll_dev, int32_t* out_status_dev) {
    return eval_resident_impl(c, params_dev, stat_dev, lnprior_dev, W, workspace, ws_bytes, (cudaStream_t)stream, out_ll_dev, out_status_dev);
}

int mope_b200_mcmc_halfstep(mope_ctx* c, const mope_sampler_state* s, int which_half, void* workspace, size_t ws_bytes, void* stream) {
    if (!c || !s || !s->pos || !s->lnprob || !s->naccept || !s->q || !s->folded || !s->factor || !s->lnprior || !s->ll || !s->status ||
        !s->lower || !s->upper || !s->prior_slope)
        return fail(MOPE_E_INVALID, "mcmc_halfstep: null argument");
    if (s->n_walkers < 2 || (s->n_walkers & 1)) return fail(MOPE_E_INVALID, "mcmc_halfstep: the stretch move needs an even number of walkers");
    if (s->zu && !s->partner) return fail(MOPE_E_INVALID, "mcmc_halfstep: partner indices are required with a supplied z stream");
    if (!(s->a > 1.0)) return fail(MOPE_E_INVALID, "mcmc_halfstep: stretch scale a must exceed 1");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int half = s->n_walkers / 2;
    SamplerParams p{};
    p.pos = s->pos; p.lnprob = s->lnprob; p.naccept = reinterpret_cast<long long*>(s->naccept);
    p.q = s->q; p.folded = s->folded; p.factor = s->factor; p.lnprior = s->lnprior; p.ll = s->ll;
    p.lower = s->lower; p.upper = s->upper; p.prior_slope = s->prior_slope; p.prior_const = s->prior_const;
    p.zu = s->zu; p.au = s->au; p.partner = s->partner;
    p.seed = s->seed; p.step = s->step; p.a = s->a;
    p.ndim = c->ndim; p.n_var = c->n_var; p.half = half;
    p.first = which_half ? half : 0;
    p.other_first = which_half ? 0 : half;
    stretch_propose_kernel<<<(half + 127) / 128, 128, 0, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    if (int rc = eval_resident_impl(c, s->folded, nullptr, s->lnprior, half, workspace, ws_bytes, st, s->ll, s->status)) return rc;
    stretch_accept_kernel<<<(half + 127) / 128, 128, 0, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    return 0;
}

int64_t mope_b200_launch_count(const mope_ctx* c) { return c ? c->launches : 0; }

int mope_b200_schedule_stats(const mope_model_desc* m, int F, int ped, int32_t* out) {
    if (!m || !out || ped < 0 || ped >= m->n_ped || F < 2) return fail(MOPE_E_INVALID, "schedule_stats: bad argument");
    const int Fp = round_up_i(F, 16), NT = (F + 7) / 8;
    out[0] = out[1] = out[2] = out[3] = 0;
    if (Fp > 208) return 0;   // wide grids use tree_kernel, whose messages all go through its L2 scratch ring
    const int ntt = NT <= 10 ? 10 : (NT <= 16 ? 16 : 26);
    const mope_pedigree_desc& pd = m->peds[ped];
    std::vector<int> fk(pd.n_branches, 0), rk(pd.n_branches, 0);
    Pedigree P;
    if (int rc = compile_schedule4(pd, fk, rk, 256 / (ntt * 4) - 1, P)) return rc;
    out[0] = P.parks_chain4; out[1] = P.parks_tm4; out[2] = P.parks_l2_4; out[3] = P.n_slots4;
    return 0;
}

int mope_b200_last_eval_stats(const mope_ctx* c, int64_t* gemm_ops, int64_t* identity_ops) {
    if (!c) return fail(MOPE_E_INVALID, "null ctx");
    if (gemm_ops) *gemm_ops = c->last_gemm_ops;
    if (identity_ops) *identity_ops = c->last_identity_ops;
    return 0;
}

int mope_b200_last_eval_unit_rows(mope_ctx* c, int64_t* unit_rows) {
    if (!c || !unit_rows) return fail(MOPE_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->last_stream));
    unsigned long long v = 0;
    CUDA_TRY(cudaMemcpy(&v, c->work_ctr + 8, sizeof v, cudaMemcpyDeviceToHost));
    *unit_rows = (int64_t)((double)v / (double)c->F + 0.5);   // the kernels count rows x k columns; a full unit has F of them
    return 0;
}

int mope_b200_last_eval_segment_stats(mope_ctx* c, int64_t* warp_segments, int64_t* active_warp_segments) {
    if (!c || !warp_segments || !active_warp_segments) return fail(MOPE_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->last_stream));
    unsigned long long v[3] = {0, 0, 0};
    CUDA_TRY(cudaMemcpy(v, c->work_ctr + 8, sizeof v, cudaMemcpyDeviceToHost));
    *warp_segments = (int64_t)v[1];
    *active_warp_segments = (int64_t)v[2];
    return 0;
}

int mope_b200_set_profiling(mope_ctx* c, int on) {
    if (!c) return fail(MOPE_E_INVALID, "null ctx");
    c->profiling = on != 0;
    return 0;
}

int mope_b200_kernel_times(mope_ctx* c, double* tree_ms, int64_t* tree_launches, double* interp_ms, int64_t* interp_launches) {
    if (!c) return fail(MOPE_E_INVALID, "null ctx");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaDeviceSynchronize());
    auto drain = [&](std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& v, double* ms, int64_t* n) {
        double tot = 0;
        for (auto& pr : v) {
            float m = 0;
            cudaEventElapsedTime(&m, pr.first, pr.second);
            tot += m;
            c->ev_pool.push_back(pr.first);
            c->ev_pool.push_back(pr.second);
        }
        if (ms) *ms = tot;
        if (n) *n = (int64_t)v.size();
        v.clear();
    };
    drain(c->ev_tree, tree_ms, tree_launches);
    drain(c->ev_interp, interp_ms, interp_launches);
    return 0;
}

int mope_b200_fp64_peak(mope_ctx* c, double millis, void* stream, double* out_tflops) {
    if (!c || !out_tflops) return fail(MOPE_E_INVALID, "fp64_peak: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (ensure_opbuf(c, 4096)) return MOPE_E_CUDA;
    double* out = reinterpret_cast<double*>(c->opbuf);
    const int grid = c->num_sms * 2, threads = 512, iters = 20000;
    const double flop_per_launch = 2.0 * 256 * 8 * (double)iters * ((double)grid * threads / 32);
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    dmma_peak_kernel<<<grid, threads, 0, st>>>(out, iters, 1.0000001, 1e-9);  // warm-up
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(st));
    int reps = 0;
    double total_ms = 0;
    CUDA_TRY(cudaEventRecord(e0, st));
    do {
        dmma_peak_kernel<<<grid, threads, 0, st>>>(out, iters, 1.0000001, 1e-9);
        ++reps;
        c->launches++;
        CUDA_TRY(cudaEventRecord(e1, st));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        total_ms = ms;
    } while (total_ms < millis && reps < 10000);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *out_tflops = flop_per_launch * reps / (total_ms * 1e-3) * 1e-12;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// operator-level entry points
// ------------------------------------------------------------------------------------------------
int mope_b200_transition_matrix(mope_ctx* c, int bottleneck, double x, double scaled_mut, void* stream, double* out_P,
                                int32_t* out_status) {
    if (!c || !out_P || !out_status) return fail(MOPE_E_INVALID, "transition_matrix: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    WalkerPlan wp;
    wp.status = status_fixed(c, bottleneck != 0, x, scaled_mut);
    *out_status = wp.status;
    if (wp.status) return 0;
    PlanOut po;
    size_t b0 = 0;
    emit_fixed_job(c, bottleneck != 0, x, scaled_mut, b0, po);
    const int F = c->F, Fp = c->Fp;
    if (po.jobs.empty()) {  // identity short-cut (transition_data_2d.py:184-187)
        for (int i = 0; i < F; ++i) for (int j = 0; j < F; ++j) out_P[(size_t)i * F + j] = (i == j) ? 1.0 : 0.0;
        return 0;
    }
    const size_t blk = blk_doubles(Fp);
    const size_t need = align_up(sizeof(InterpJob), 256) + blk * sizeof(double) + (size_t)F * F * sizeof(double) + 1024;
    if (ensure_opbuf(c, need)) return MOPE_E_CUDA;
    InterpJob* d_job = reinterpret_cast<InterpJob*>(c->opbuf);
    double* d_blk = reinterpret_cast<double*>(c->opbuf + align_up(sizeof(InterpJob), 256));
    double* d_out = d_blk + blk;
    CUDA_TRY(cudaMemcpyAsync(d_job, po.jobs.data(), sizeof(InterpJob), cudaMemcpyHostToDevice, st));
    if (int rc = launch_interp(c, d_job, 1, d_blk, st, false)) return rc;
    const long long total = (long long)F * F;
    unpad_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_blk, d_out, F, F, Fp);
    CUDA_TRY(cudaGetLastError());
    c->launches += 1;
    CUDA_TRY(cudaMemcpyAsync(out_P, d_out, (size_t)F * F * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int mope_b200_leaf_likelihoods(mope_ctx* c, const double* counts, const double* coverage, int L, int S, const double* freqs,
                               int F, double min_phred, void* stream, double* out) {
    if (!c || !counts || !coverage || !freqs || !out || L < 0 || S < 1 || F < 1) return fail(MOPE_E_INVALID, "leaf_likelihoods: bad argument");
    if (L == 0) return 0;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int Fp = round_up_i(F, 16);
    const int R_pad = round_up_i(L, TM);
    std::vector<double> logp(F), log1mp(F);
    std::vector<int8_t> pclass(F), zeros(F, 0);
    const double perr = (min_phred != -1.0) ? std::pow(10.0, -1.0 * min_phred / 10.0) : 0.0;
    for (int f = 0; f < F; ++f) {
        const double p = (1.0 - perr) * freqs[f] + (perr / 3.0) * (1.0 - freqs[f]);
        pclass[f] = (p == 0.0) ? 1 : ((p == 1.0) ? 2 : 0);
        logp[f] = (p > 0.0) ? std::log(p) : 0.0;
        log1mp[f] = (p < 1.0) ? std::log1p(-p) : 0.0;
    }
    Bump b;
    b.dry = true;
    double* dlc = nullptr;
    auto carve = [&](Bump& bb, double*& dc, double*& dn, double*& dl, double*& dl1, int8_t*& cls, int8_t*& z, double*& E, double*& o) {
        dc = bb.take<double>((size_t)L * S); dn = bb.take<double>((size_t)L * S); dlc = bb.take<double>((size_t)L * S);
        dl = bb.take<double>(F); dl1 = bb.take<double>(F); cls = bb.take<int8_t>(F); z = bb.take<int8_t>(F);
        E = bb.take<double>((size_t)S * R_pad * Fp); o = bb.take<double>((size_t)L * S * F);
    };
    double *dc, *dn, *dl, *dl1, *E, *o; int8_t *cls, *z;
    carve(b, dc, dn, dl, dl1, cls, z, E, o);
    if (ensure_opbuf(c, b.used + 256)) return MOPE_E_CUDA;
    Bump r;
    r.base = c->opbuf; r.cap = c->opbuf_cap;
    carve(r, dc, dn, dl, dl1, cls, z, E, o);
    CUDA_TRY(cudaMemcpyAsync(dc, counts, (size_t)L * S * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dn, coverage, (size_t)L * S * sizeof(double), cudaMemcpyHostToDevice, st));
    std::vector<double> lnc((size_t)L * S);
    host_lnchoose(counts, coverage, (size_t)L * S, lnc.data());
    CUDA_TRY(cudaMemcpyAsync(dlc, lnc.data(), (size_t)L * S * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dl, logp.data(), F * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dl1, log1mp.data(), F * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(cls, pclass.data(), F, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(z, zeros.data(), F, cudaMemcpyHostToDevice, st));
    EmissionParams ep{};
    ep.counts = dc; ep.coverage = dn; ep.lnchoose = dlc; ep.logp = dl; ep.log1mp = dl1; ep.pclass = cls; ep.asc_e0 = z; ep.asc_eN = z; ep.E = E;
    ep.L = L; ep.S = S; ep.R = L; ep.R_pad = R_pad; ep.F = F; ep.Fp = Fp;
    ep.perm = 0;
    const long long warps = (long long)S * R_pad;
    emission_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(ep);
    CUDA_TRY(cudaGetLastError());
    // [S][R_pad][Fp] -> [L][S][F]
    const long long total = (long long)L * S * F;
    emission_to_lsf_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(E, o, L, S, F, Fp, R_pad);
    CUDA_TRY(cudaGetLastError());
    c->launches += 2;
    CUDA_TRY(cudaMemcpyAsync(out, o, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int mope_b200_branch_update(mope_ctx* c, int kind, const double* child, double* parent, int n_rows, const double* lengths,
                            double scaled_mut, void* stream, int32_t* out_status) {
    if (!c || !child || !parent || !lengths || !out_status || n_rows < 0 || kind < 0 || kind > 2)
        return fail(MOPE_E_INVALID, "branch_update: bad argument");
    *out_status = 0;
    if (n_rows == 0) return 0;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int F = c->F, Fp = c->Fp;
    const int R_pad = round_up_i(n_rows, TM);
    const size_t blk = blk_doubles(Fp);

    // plan
    PlanOut po;
    RateInfo ri{};
    int32_t status = 0;
    size_t b0 = 0;
    if (kind == 1) {
        double mn = lengths[0], mx = lengths[0];
        for (int r = 1; r < n_rows; ++r) { mn = std::min(mn, lengths[r]); mx = std::max(mx, lengths[r]); }
        // lengths enter as multipliers of a unit branch length (mult * 1.0 is exact)
        status = status_rate(c, mn, mx, 1.0, scaled_mut);
        if (!status) { emit_rate_jobs(c, mn, mx, 1.0, scaled_mut, b0, po); ri = po.rate[0]; }
    } else {
        status = status_fixed(c, kind == 2, lengths[0], scaled_mut);
        if (!status) emit_fixed_job(c, kind == 2, lengths[0], scaled_mut, b0, po);
    }
    *out_status = status;
    if (status) return 0;
    const size_t n_blocks = po.jobs.size();

    Bump b;
    b.dry = true;
    struct Ptrs { InterpJob* jobs; int64_t* mo; RateInfo* rate; double *pool, *E, *Y, *par, *mult, *raw; TreeOp* op; } P{};
    auto carve = [&](Bump& bb) {
        P.jobs = bb.take<InterpJob>(std::max<size_t>(n_blocks, 1));
        P.mo = bb.take<int64_t>(1);
        P.rate = bb.take<RateInfo>(1);
        P.op = bb.take<TreeOp>(1);
        P.pool = bb.take<double>(std::max<size_t>(n_blocks, 1) * blk);
        P.E = bb.take<double>((size_t)R_pad * Fp);
        P.Y = bb.take<double>((size_t)R_pad * Fp);
        P.par = bb.take<double>((size_t)n_rows * F);
        P.raw = bb.take<double>((size_t)n_rows * F);
        P.mult = bb.take<double>(R_pad);
    };
    carve(b);
    if (ensure_opbuf(c, b.used + 256)) return MOPE_E_CUDA;
    Bump r;
    r.base = c->opbuf; r.cap = c->opbuf_cap;
    carve(r);

    if (n_blocks) CUDA_TRY(cudaMemcpyAsync(P.jobs, po.jobs.data(), n_blocks * sizeof(InterpJob), cudaMemcpyHostToDevice, st));
    const int64_t mat_off0 = po.mat_off.empty() ? 0 : po.mat_off[0];  // -1: identity short-cut
    CUDA_TRY(cudaMemcpyAsync(P.mo, &mat_off0, sizeof(int64_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(P.rate, &ri, sizeof(RateInfo), cudaMemcpyHostToDevice, st));
    TreeOp op{};
    op.nsrc = 1; op.src_kind[0] = 0; op.src[0] = 0; op.src_kind[1] = 0; op.src[1] = 0; op.dst = -1; op.first = 1;
    op.kind = (kind == 1) ? 1 : 0; op.key = 0; op.mult = 0; op.last = 0; op.sib[0] = op.sib[1] = -1; op.dep_prev = 0;
    CUDA_TRY(cudaMemcpyAsync(P.op, &op, sizeof(TreeOp), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(P.raw, child, (size_t)n_rows * F * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(P.par, parent, (size_t)n_rows * F * sizeof(double), cudaMemcpyHostToDevice, st));
    {
        std::vector<double> mu(R_pad, 0.0);
        if (kind == 1) for (int i = 0; i < n_rows; ++i) mu[i] = lengths[i];
        CUDA_TRY(cudaMemcpyAsync(P.mult, mu.data(), (size_t)R_pad * sizeof(double), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    const long long tot = (long long)R_pad * Fp;
    pad_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(P.raw, P.E, n_rows, F, Fp, R_pad, c->use_v4 ? 1 : 0);
    CUDA_TRY(cudaGetLastError());
    if (n_blocks) {
        if (int rc = launch_interp(c, P.jobs, n_blocks, P.pool, st, c->use_v4)) return rc;
    }
    if (c->use_v4) {
        Op4 o4{};
        o4.leaf = 0; o4.kind = (kind == 1) ? 1 : 0; o4.key = 0; o4.mult = 0; o4.mul_slot = -1; o4.store_slot = -1; o4.store_first = 0; o4.last = 0; o4.home = HOME_L2;
        static_assert(sizeof(Op4) <= sizeof(TreeOp), "the single-op staging buffer is sized for TreeOp");
        CUDA_TRY(cudaMemcpyAsync(P.op, &o4, sizeof(Op4), cudaMemcpyHostToDevice, st));
        Tree4Params t4{};
        if (int rc = make_tmap(&t4.tmapE, P.E, (uint64_t)R_pad, Fp, TM)) return rc;
        if (int rc = make_tmap(&t4.tmapP, P.pool, (uint64_t)std::max<size_t>(n_blocks, 1) * (Fp + BLK_EXTRA), Fp, tree_prows(c))) return rc;
        t4.E = P.E; t4.E_leaf_stride = (int64_t)R_pad * Fp; t4.scratch = P.Y /*unused*/; t4.ops = reinterpret_cast<const Op4*>(P.op);
        t4.n_ops = 1; t4.n_slots = 1;
        t4.R = n_rows; t4.L = n_rows; t4.R_pad = R_pad; t4.Fp = Fp; t4.F = F; t4.NT = c->NT; t4.n_tiles = R_pad / TM; t4.W = 1;
        t4.tile_block = t4.n_tiles;
        t4.pool = P.pool; t4.mat_off = P.mo; t4.n_mat = 1; t4.n_rate = 1; t4.rate = P.rate; t4.mult = P.mult; t4.gens = c->gens_dev;
        t4.n_gens = c->n_gens; t4.Npop = c->N; t4.stat = nullptr; t4.tile_ll = nullptr; t4.asc_dot = nullptr; t4.Yout = P.Y;
        if (launch_tree4(c, t4, std::min(t4.n_tiles, c->num_sms), st)) return MOPE_E_CUDA;
    } else {
    TreeParams tp{};
    if (int rc = make_tmap(&tp.tmapE, P.E, (uint64_t)R_pad, Fp, TM)) return rc;
    if (int rc = make_tmap(&tp.tmapS, P.Y, (uint64_t)R_pad, Fp, TM)) return rc;   // no scratch ring in single-op mode
    if (int rc = make_tmap(&tp.tmapP, P.pool, (uint64_t)std::max<size_t>(n_blocks, 1) * (Fp + BLK_EXTRA), Fp, tree_prows(c))) return rc;
    tp.E = P.E; tp.E_leaf_stride = (int64_t)R_pad * Fp; tp.scratch = P.Y /*unused*/; tp.ops = P.op; tp.n_ops = 1; tp.n_slots = 0;
    tp.R = n_rows; tp.L = n_rows; tp.R_pad = R_pad; tp.Fp = Fp; tp.F = F; tp.NT = c->NT; tp.n_tiles = R_pad / TM; tp.W = 1;
    tp.tile_block = tp.n_tiles;
    tp.pool = P.pool; tp.mat_off = P.mo; tp.n_mat = 1; tp.n_rate = 1; tp.rate = P.rate; tp.mult = P.mult; tp.gens = c->gens_dev;
    tp.n_gens = c->n_gens; tp.Npop = c->N; tp.stat = nullptr; tp.tile_ll = nullptr; tp.asc_dot = nullptr; tp.Yout = P.Y;
    if (launch_tree(c, tp, std::min(tp.n_tiles, c->num_sms), st)) return MOPE_E_CUDA;
    }
    const long long tot2 = (long long)n_rows * F;
    mul_unpad_kernel<<<(unsigned)((tot2 + 255) / 256), 256, 0, st>>>(P.Y, P.par, n_rows, F, Fp);
    CUDA_TRY(cudaGetLastError());
    c->launches += 3;
    CUDA_TRY(cudaMemcpyAsync(parent, P.par, (size_t)n_rows * F * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int mope_b200_non_asc_probs(mope_ctx* c, const double* qualities, const int64_t* qual_off, int n_sets, const double* freqs, int F,
                            double min_freq, void* stream, double* out) {
    if (!c || !qualities || !qual_off || !freqs || !out || n_sets < 0 || F < 1) return fail(MOPE_E_INVALID, "non_asc_probs: bad argument");
    if (n_sets == 0) return 0;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nq = qual_off[n_sets];
    // sets whose K = ceil(min_freq * nreads) exceeds one register block need a carry scratch of 2 x nreads x F doubles
    std::vector<int64_t> coff(n_sets, 0);
    size_t carry_doubles = 0;
    for (int s = 0; s < n_sets; ++s) {
        if (qual_off[s + 1] < qual_off[s]) return fail(MOPE_E_INVALID, "non_asc_probs: offsets not ascending");
        const int64_t n = qual_off[s + 1] - qual_off[s];
        const int K = (int)std::ceil(min_freq * (double)n);
        if (K > PB_BLK) { coff[s] = (int64_t)carry_doubles; carry_doubles += (size_t)2 * n * F; }
    }
    if (carry_doubles * sizeof(double) > ((size_t)8 << 30))
        return fail(MOPE_E_NOMEM, "non_asc_probs: the read sets of this call need %zu bytes of carry scratch; pass fewer sets per call", carry_doubles * sizeof(double));
    Bump b;
    b.dry = true;
    double *dq, *df, *dout, *dcarry; int64_t *doff, *dcoff;
    auto carve = [&](Bump& bb) {
        dq = bb.take<double>(std::max<int64_t>(nq, 1)); doff = bb.take<int64_t>(n_sets + 1); dcoff = bb.take<int64_t>(n_sets);
        df = bb.take<double>(F); dout = bb.take<double>((size_t)n_sets * F); dcarry = bb.take<double>(std::max<size_t>(carry_doubles, 1));
    };
    carve(b);
    if (ensure_opbuf(c, b.used + 256)) return MOPE_E_CUDA;
    Bump r;
    r.base = c->opbuf; r.cap = c->opbuf_cap;
    carve(r);
    if (nq) CUDA_TRY(cudaMemcpyAsync(dq, qualities, (size_t)nq * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(doff, qual_off, (size_t)(n_sets + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dcoff, coff.data(), (size_t)n_sets * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(df, freqs, (size_t)F * sizeof(double), cudaMemcpyHostToDevice, st));
    pb_kernel<<<n_sets, 128, 0, st>>>(dq, doff, df, F, min_freq, dout, dcarry, dcoff);
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    CUDA_TRY(cudaMemcpyAsync(out, dout, (size_t)n_sets * F * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int mope_b200_beta_cdf_table(void* betainc_fn, const double* ab, int n_ab, const double* edges, int n_edges, double* out, int n_threads) {
    if (!betainc_fn || !ab || !edges || !out || n_ab < 0 || n_edges < 0) return fail(MOPE_E_INVALID, "beta_cdf_table: bad argument");
    typedef double (*fn_t)(double, double, double, int);
    const fn_t fn = reinterpret_cast<fn_t>(betainc_fn);
    auto rows = [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i) {
            const double a = ab[i];
            double* o = out + (size_t)i * n_edges;
            for (int j = 0; j < n_edges; ++j) o[j] = fn(a, a, edges[j], 1);
        }
    };
    const int nt = std::max(1, std::min(n_threads, n_ab / 4));
    if (nt <= 1) { rows(0, n_ab); return 0; }
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (int t = 1; t < nt; ++t) th.emplace_back(rows, (int)((long long)n_ab * t / nt), (int)((long long)n_ab * (t + 1) / nt));
    rows(0, (int)((long long)n_ab / nt));
    for (auto& x : th) x.join();
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// transition-table generator (device pointers; no context: the tables are made before any model exists)
// ------------------------------------------------------------------------------------------------
namespace {
// 2-D fp64 tensor [rows][cols] with row pitch ld doubles, box = 16 columns x box_rows rows, 128-byte swizzle;
// out-of-range parts of a box read as zero
int make_tmap_ld(CUtensorMap* out, const double* base, uint64_t rows, uint64_t cols, int ld, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(MOPE_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(double)};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MOPE_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows %llu, cols %llu, ld %d)", (int)r,
                                       (unsigned long long)rows, (unsigned long long)cols, ld);
    return 0;
}

template <int NTW>
int launch_dgemm_tn(const double* A, const double* Bt, double* C, int M, int N, int K, int lda, int ldbt, int ldc, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(dgemm_tn_kernel<NTW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmCfg<NTW>::SMEM_BYTES));
        attr_set = true;
    }
    GemmParams gp{};
    if (int rc = make_tmap_ld(&gp.tmapA, A, (uint64_t)M, (uint64_t)K, lda, TM)) return rc;
    if (int rc = make_tmap_ld(&gp.tmapB, Bt, (uint64_t)N, (uint64_t)K, ldbt, 16 * NTW)) return rc;
    gp.C = C; gp.M = M; gp.N = N; gp.K = K; gp.ldc = ldc;
    gp.tiles_n = (N + 16 * NTW - 1) / (16 * NTW);
    const int tiles_m = (M + TM - 1) / TM;
    dgemm_tn_kernel<NTW><<<tiles_m * gp.tiles_n, NTHREADS, GemmCfg<NTW>::SMEM_BYTES, st>>>(gp);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
}  // namespace

extern "C" {

int mope_b200_gen_binom_matrix(double* P_dev, int ld, int rows, int cols, int rows_valid, int n_trials, const double* p_host,
                               const double* lnchoose_host, void* stream) {
    if (!P_dev || !p_host || rows <= 0 || cols <= 0 || ld < cols || rows_valid < 0 || rows_valid > rows || n_trials < 0)
        return fail(MOPE_E_INVALID, "gen_binom_matrix: bad argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int nj = n_trials + 1;
    std::vector<double> buf((size_t)2 * rows_valid + nj);
    std::vector<int8_t> cls(std::max(rows_valid, 1));
    double* logp = buf.data(); double* log1mp = logp + rows_valid; double* lnc = log1mp + rows_valid;
    for (int i = 0; i < rows_valid; ++i) {
        const double p = p_host[i];
        if (!(p >= 0.0 && p <= 1.0)) return fail(MOPE_E_INVALID, "gen_binom_matrix: probability %g of row %d outside [0, 1]", p, i);
        cls[i] = p == 0.0 ? 1 : (p == 1.0 ? 2 : 0);
        logp[i] = cls[i] ? 0.0 : std::log(p);
        log1mp[i] = cls[i] ? 0.0 : std::log1p(-p);
    }
    if (lnchoose_host) {
        std::memcpy(lnc, lnchoose_host, (size_t)nj * sizeof(double));
    } else {
        const double lg = std::lgamma((double)n_trials + 1.0);
        for (int j = 0; j < nj; ++j) lnc[j] = lg - (std::lgamma((double)j + 1.0) + std::lgamma((double)(n_trials - j) + 1.0));
    }
    double* d_buf = nullptr; int8_t* d_cls = nullptr;
    CUDA_TRY(cudaMallocAsync(&d_buf, buf.size() * sizeof(double), st));
    CUDA_TRY(cudaMallocAsync(&d_cls, cls.size(), st));
    CUDA_TRY(cudaMemcpyAsync(d_buf, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_cls, cls.data(), cls.size(), cudaMemcpyHostToDevice, st));
    dim3 grid((ld + 255) / 256, rows);
    binom_matrix_kernel<<<grid, 256, 0, st>>>(P_dev, ld, rows, cols, rows_valid, n_trials, d_buf, d_buf + rows_valid, d_buf + 2 * rows_valid, d_cls);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaFreeAsync(d_buf, st));
    CUDA_TRY(cudaFreeAsync(d_cls, st));
    return 0;
}

int mope_b200_gen_transpose(const double* A_dev, double* At_dev, int rows, int cols, int lda, int ldt, void* stream) {
    if (!A_dev || !At_dev || rows <= 0 || cols <= 0 || lda < cols || ldt < rows) return fail(MOPE_E_INVALID, "gen_transpose: bad argument");
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(A_dev, At_dev, rows, cols, lda, ldt);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int mope_b200_gen_dgemm_tn(const double* A_dev, const double* Bt_dev, double* C_dev, int M, int N, int K, int lda, int ldbt, int ldc, void* stream) {
    if (!A_dev || !Bt_dev || !C_dev || M <= 0 || N <= 0 || K <= 0 || lda < K || ldbt < K || ldc < N)
        return fail(MOPE_E_INVALID, "gen_dgemm_tn: bad argument");
    if ((lda & 1) || (ldbt & 1) || (ldc & 1) || ((uintptr_t)A_dev & 15) || ((uintptr_t)Bt_dev & 15) || ((uintptr_t)C_dev & 15))
        return fail(MOPE_E_INVALID, "gen_dgemm_tn: operands must be 16-byte aligned with even leading dimensions (TMA row pitch)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // tile width: the one whose tile count wastes the fewest SM slots (1001 x 1001: 16 x 9 = 144 tiles of 64 x 112)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int tiles_m = (M + TM - 1) / TM;
    auto cost = [&](int ntw) {   // waves x tile work
        const long long tiles = (long long)tiles_m * ((N + 16 * ntw - 1) / (16 * ntw));
        return (double)((tiles + sms - 1) / sms) * ntw;
    };
    const double c7 = cost(7), c13 = cost(13);
    if (c13 < c7) return launch_dgemm_tn<13>(A_dev, Bt_dev, C_dev, M, N, K, lda, ldbt, ldc, st);
    return launch_dgemm_tn<7>(A_dev, Bt_dev, C_dev, M, N, K, lda, ldbt, ldc, st);
}

int mope_b200_gen_bin_matrix(const double* P_dev, int ld, int n_states, const int32_t* breaks_dev, int F, double* out_dev, void* stream) {
    if (!P_dev || !breaks_dev || !out_dev || n_states <= 0 || F <= 0 || F > n_states || ld < n_states) return fail(MOPE_E_INVALID, "gen_bin_matrix: bad argument");
    bin_matrix_kernel<<<(F * F + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P_dev, ld, n_states, breaks_dev, F, out_dev);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// selection model / masked-matrix variants (selection.cuh)
// ------------------------------------------------------------------------------------------------
namespace {
struct SelPlan {
    PlanOut po;                       // jobs + mat_off per distinct (time, alpha) pair
    std::vector<int32_t> row_off, rows;
    size_t n_blocks = 0;
    int32_t status = 0;
};

// distinct (time, alpha) pairs of the rows -> interpolation jobs + row lists
void plan_sel_rows(const mope_ctx* c, bool bott, int n_rows, const double* times, int time_stride, const double* alphas, SelPlan& pl) {
    std::vector<int32_t> order(n_rows);
    for (int i = 0; i < n_rows; ++i) order[i] = i;
    auto key_less = [&](int a, int b) {
        const double ta = times[(size_t)a * time_stride], tb = times[(size_t)b * time_stride];
        if (ta != tb) return ta < tb;
        if (alphas[a] != alphas[b]) return alphas[a] < alphas[b];
        return a < b;
    };
    std::sort(order.begin(), order.end(), key_less);
    pl.rows = order;
    pl.row_off.clear();
    size_t b = 0;
    for (int k = 0; k < n_rows; ++k) {
        const int r = order[k];
        const double t = times[(size_t)r * time_stride], al = alphas[r];
        const bool fresh = k == 0 || !(t == times[(size_t)order[k - 1] * time_stride] && al == alphas[order[k - 1]]);
        if (!fresh) continue;
        pl.row_off.push_back(k);
        const int32_t st = status_fixed(c, bott, t, al);
        pl.status |= st;
        if (st) { pl.po.mat_off.push_back(-1); continue; }
        emit_fixed_job(c, bott, t, al, b, pl.po);
    }
    pl.row_off.push_back(n_rows);
    pl.n_blocks = b;
}

// device side of one branch: interpolate the distinct matrices into `pool`, apply them (plan arrays staged in `stage`)
int run_sel_branch(mope_ctx* c, const SelPlan& pl, char* stage, size_t stage_bytes, double* pool, const double* V, double* parent, double* trans,
                   int mask, cudaStream_t st) {
    const int n_jobs = (int)pl.po.mat_off.size();
    if (n_jobs == 0) return 0;
    Bump sb;
    sb.base = stage; sb.cap = stage_bytes;
    InterpJob* d_jobs = sb.take<InterpJob>(std::max<size_t>(pl.po.jobs.size(), 1));
    int64_t* d_off = sb.take<int64_t>(n_jobs);
    int32_t* d_roff = sb.take<int32_t>(n_jobs + 1);
    int32_t* d_rows = sb.take<int32_t>(std::max<size_t>(pl.rows.size(), 1));
    if (!sb.ok()) return fail(MOPE_E_NOMEM, "selection: staging area too small");
    if (!pl.po.jobs.empty()) {
        CUDA_TRY(cudaMemcpyAsync(d_jobs, pl.po.jobs.data(), pl.po.jobs.size() * sizeof(InterpJob), cudaMemcpyHostToDevice, st));
        if (int rc = launch_interp(c, d_jobs, pl.po.jobs.size(), pool, st, false)) return rc;
    }
    // (the host-planned operator path keeps the two-kernel form: its job list holds no entries for identity groups)
    CUDA_TRY(cudaMemcpyAsync(d_off, pl.po.mat_off.data(), (size_t)n_jobs * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_roff, pl.row_off.data(), (size_t)(n_jobs + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_rows, pl.rows.data(), pl.rows.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(launch_sel_matvec(n_jobs, st, pool, d_off, d_roff, d_rows, V, parent, trans, c->F, c->Fp, mask));
    c->launches++;
    // the host vectors of the plan are pageable: the copies above were staged synchronously by the runtime
    return 0;
}
// Row groups of the selection model: on branch b the rows (locus i and its two pseudo-rows L + i, 2L + i) with the same
// (multiplier value, position) share their transition matrix for every parameter vector.  Built once per pedigree and
// position index; everything the per-evaluation planning kernel needs stays on the device.
int build_sel_cache(mope_ctx* c, Pedigree& P, const int32_t* position_idx, int n_pos, cudaStream_t st) {
    SelCache& S = P.sel;
    const int L = P.L, B = P.n_branches, R3 = 3 * L;
    if (S.valid && S.n_pos == n_pos && (int)S.pos_idx.size() == L && std::memcmp(S.pos_idx.data(), position_idx, (size_t)L * sizeof(int32_t)) == 0)
        return 0;
    for (int i = 0; i < L; ++i)
        if (position_idx[i] < 0 || position_idx[i] >= n_pos) return fail(MOPE_E_INVALID, "sel_loglike: position index %d of row %d out of range", position_idx[i], i);
    if (S.dev) { CUDA_TRY(cudaStreamSynchronize(st)); cudaFree(S.dev); S.dev = nullptr; }
    S.valid = false;
    S.pos_idx.assign(position_idx, position_idx + L);
    S.n_pos = n_pos;
    S.br_first.assign(B, 0);
    S.br_groups.assign(B, 0);
    std::vector<double> g_mult;
    std::vector<int32_t> g_pos, g_branch, g_local, g_row_off, g_rows((size_t)B * R3);
    std::vector<int32_t> order(L);
    for (int b = 0; b < B; ++b) {
        const int mc = P.br_mult_h[b];
        const double* mult = mc >= 0 ? P.row_mult_h.data() + (size_t)mc * L : nullptr;
        for (int i = 0; i < L; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int a, int bq) {
            const double ma = mult ? mult[a] : 1.0, mb = mult ? mult[bq] : 1.0;
            if (ma != mb) return ma < mb;                  // NaN multipliers were rejected at create
            if (position_idx[a] != position_idx[bq]) return position_idx[a] < position_idx[bq];
            return a < bq;
        });
        S.br_first[b] = (int)g_mult.size();
        int32_t* rows = g_rows.data() + (size_t)b * R3;
        int nrow = 0, local = 0;
        for (int k = 0; k < L; ++k) {
            const int i = order[k];
            const double m = mult ? mult[i] : 1.0;
            const bool fresh = k == 0 || m != (mult ? mult[order[k - 1]] : 1.0) || position_idx[i] != position_idx[order[k - 1]];
            if (fresh) {
                g_mult.push_back(m); g_pos.push_back(position_idx[i]); g_branch.push_back(b); g_local.push_back(local++);
                g_row_off.push_back(nrow);
            }
            rows[nrow++] = i; rows[nrow++] = L + i; rows[nrow++] = 2 * L + i;   // np.tile(..., 3): likelihood, all-zero, all-one rows
        }
        g_row_off.push_back(nrow);
        S.br_groups[b] = local;
    }
    S.total_groups = (int)g_mult.size();
    const size_t TG = g_mult.size();
    const size_t bytes = align_up(TG * 8, 256) + 3 * align_up(TG * 4, 256) + align_up(g_row_off.size() * 4, 256) + align_up(g_rows.size() * 4, 256) + 1024;
    CUDA_TRY(cudaMalloc(&S.dev, bytes));
    Bump bb;
    bb.base = S.dev; bb.cap = bytes;
    S.g_mult = bb.take<double>(TG); S.g_pos = bb.take<int32_t>(TG); S.g_branch = bb.take<int32_t>(TG); S.g_local = bb.take<int32_t>(TG);
    S.g_row_off = bb.take<int32_t>(g_row_off.size()); S.g_rows = bb.take<int32_t>(g_rows.size());
    if (!bb.ok()) return fail(MOPE_E_NOMEM, "sel_loglike: group cache layout");
    CUDA_TRY(cudaMemcpy(S.g_mult, g_mult.data(), TG * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(S.g_pos, g_pos.data(), TG * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(S.g_branch, g_branch.data(), TG * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(S.g_local, g_local.data(), TG * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(S.g_row_off, g_row_off.data(), g_row_off.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(S.g_rows, g_rows.data(), g_rows.size() * 4, cudaMemcpyHostToDevice));
    S.valid = true;
    return 0;
}

size_t sel_stage_bytes(int n_rows) {
    return align_up((size_t)n_rows * sizeof(InterpJob), 256) + align_up((size_t)n_rows * 8, 256) + align_up((size_t)(n_rows + 1) * 4, 256) +
           align_up((size_t)n_rows * 4, 256) + 1024;
}
}  // namespace

extern "C" {

int mope_b200_sel_branch_update(mope_ctx* c, int mask, const double* child, double* parent, int n_rows, const double* times, int per_row_times,
                                const double* alphas, double* out_trans, void* stream, int32_t* out_status) {
    if (!c || !child || !parent || !times || !alphas || !out_status || n_rows <= 0 || mask < 0 || mask > 7)
        return fail(MOPE_E_INVALID, "sel_branch_update: bad argument");
    const bool bott = (mask & MOPE_MASK_BOTTLENECK) != 0;
    mask &= 3;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int F = c->F, Fp = c->Fp;
    SelPlan pl;
    plan_sel_rows(c, bott, n_rows, times, per_row_times ? 1 : 0, alphas, pl);
    *out_status = pl.status;
    if (pl.status) return 0;   // the reference raises at the first matrix fetch: nothing is modified
    const size_t blk = blk_doubles(Fp);
    const size_t stage_b = sel_stage_bytes(n_rows);
    const size_t raw_b = align_up((size_t)n_rows * F * sizeof(double), 256), pad_b = align_up((size_t)n_rows * Fp * sizeof(double), 256);
    const size_t need = stage_b + align_up(std::max<size_t>(pl.n_blocks, 1) * blk * sizeof(double), 256) + 2 * raw_b + 3 * pad_b + 1024;
    if (ensure_opbuf(c, need)) return MOPE_E_CUDA;
    Bump bb;
    bb.base = c->opbuf; bb.cap = c->opbuf_cap;
    char* stage = bb.take<char>(stage_b);
    double* pool = bb.take<double>(std::max<size_t>(pl.n_blocks, 1) * blk);
    double* raw_c = bb.take<double>((size_t)n_rows * F);
    double* raw_p = bb.take<double>((size_t)n_rows * F);
    double* V = bb.take<double>((size_t)n_rows * Fp);
    double* Pn = bb.take<double>((size_t)n_rows * Fp);
    double* T = bb.take<double>((size_t)n_rows * Fp);
    if (!bb.ok()) return fail(MOPE_E_NOMEM, "sel_branch_update: scratch layout");
    CUDA_TRY(cudaMemcpyAsync(raw_c, child, (size_t)n_rows * F * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(raw_p, parent, (size_t)n_rows * F * sizeof(double), cudaMemcpyHostToDevice, st));
    const long long tot = (long long)n_rows * Fp;
    pad_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(raw_c, V, n_rows, F, Fp, n_rows, 0);
    pad_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(raw_p, Pn, n_rows, F, Fp, n_rows, 0);
    CUDA_TRY(cudaGetLastError());
    c->launches += 2;
    if (int rc = run_sel_branch(c, pl, stage, stage_b, pool, V, Pn, out_trans ? T : nullptr, mask, st)) return rc;
    const long long totu = (long long)n_rows * F;
    unpad_rows_kernel<<<(unsigned)((totu + 255) / 256), 256, 0, st>>>(Pn, raw_p, n_rows, F, Fp);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(parent, raw_p, (size_t)n_rows * F * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (out_trans) {
        unpad_rows_kernel<<<(unsigned)((totu + 255) / 256), 256, 0, st>>>(T, raw_c, n_rows, F, Fp);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(out_trans, raw_c, (size_t)n_rows * F * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    c->launches += 2;
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int mope_b200_sel_loglike(mope_ctx* c, int ped, const double* branch_lengths, const double* locus_alphas, int n_pos, const int32_t* position_idx,
                          const double* stat, void* stream, double* out_loglikes, double* out_log_asc, int32_t* out_status) {
    if (!c || ped < 0 || ped >= c->n_ped || !branch_lengths || !locus_alphas || !position_idx || !stat || !out_loglikes || !out_log_asc ||
        !out_status || n_pos <= 0)
        return fail(MOPE_E_INVALID, "sel_loglike: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    Pedigree& P = c->peds[ped];
    const int F = c->F, Fp = c->Fp, L = P.L, B = P.n_branches, R3 = 3 * L;
    if (L <= 0) return fail(MOPE_E_INVALID, "sel_loglike: pedigree without loci");
    if (int rc = build_sel_cache(c, P, position_idx, n_pos, st)) return rc;
    const SelCache& S = P.sel;
    const int TG = S.total_groups;
    // buffers: cond[node] for internal nodes and the root, one leaf buffer, trans, asc, the matrix pool (a branch's groups go
    // through it in chunks of pool_blocks)
    std::vector<int> node_buf(B + 1, -1);
    int n_internal = 0;
    for (int b = 0; b < B; ++b) if (P.br_leaf_h[b] < 0) node_buf[b] = n_internal++;
    node_buf[B] = n_internal++;
    const size_t blk = blk_doubles(Fp), nodeN = (size_t)R3 * Fp;
    int max_groups = 1;
    for (int b = 0; b < B; ++b) max_groups = std::max(max_groups, S.br_groups[b]);
    const size_t pool_cap_bytes = (size_t)1 << 30;   // at most ~1 GB of interpolated matrices in flight
    const bool fused = Fp <= 32 * APPLY_MAXJ && !getenv("MOPE_SEL_UNFUSED");
    const int pool_blocks = fused ? max_groups   // nothing is materialised: one launch per branch
                                  : (int)std::max<size_t>(1, std::min<size_t>((size_t)max_groups, pool_cap_bytes / (blk * sizeof(double))));
    const size_t need = align_up((size_t)TG * sizeof(InterpJob), 256) + align_up((size_t)TG * 8, 256) + align_up((size_t)B * 8, 256) +
                        align_up((size_t)B * n_pos * 8, 256) + align_up((fused ? (size_t)1 : (size_t)pool_blocks) * blk * 8, 256) +
                        (size_t)(n_internal + 3) * align_up(nodeN * 8, 256) + align_up((size_t)Fp * 8, 256) + align_up((size_t)2 * L * 8, 256) +
                        align_up(2 * (size_t)F, 256) + 8192;
    if (ensure_opbuf(c, need)) return MOPE_E_CUDA;
    Bump bb;
    bb.base = c->opbuf; bb.cap = c->opbuf_cap;
    InterpJob* d_jobs = bb.take<InterpJob>(TG);
    int64_t* d_moff = bb.take<int64_t>(TG);
    double* d_len = bb.take<double>(B);
    double* d_alpha = bb.take<double>((size_t)B * n_pos);
    int32_t* d_status = bb.take<int32_t>(64);
    double* pool = bb.take<double>((fused ? (size_t)1 : (size_t)pool_blocks) * blk);
    double* cond = bb.take<double>((size_t)n_internal * nodeN);
    double* leafbuf = bb.take<double>(nodeN);
    double* trans = bb.take<double>(nodeN);
    double* asc = bb.take<double>((size_t)L * Fp);
    double* d_stat = bb.take<double>(Fp);
    double* d_out = bb.take<double>((size_t)2 * L);
    int8_t* d_e = bb.take<int8_t>(2 * (size_t)F);
    if (!bb.ok()) return fail(MOPE_E_NOMEM, "sel_loglike: scratch layout");
    // ---- plan on the device: (time, alpha) of every group, range checks, interpolation jobs ----
    CUDA_TRY(cudaMemcpyAsync(d_len, branch_lengths, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_alpha, locus_alphas, (size_t)B * n_pos * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t), st));
    {
        SelPlanParams sp{};
        sp.g_mult = S.g_mult; sp.g_pos = S.g_pos; sp.g_branch = S.g_branch; sp.g_local = S.g_local;
        sp.branch_len = d_len; sp.alphas = d_alpha; sp.n = TG; sp.n_pos = n_pos; sp.pool_blocks = pool_blocks;
        sp.gens = c->gens_dev; sp.us = c->us_dev; sp.n_gens = c->n_gens; sp.n_us = c->n_us;
        sp.N = c->N; sp.max_coal = c->max_coal; sp.min_mut = c->min_mut; sp.max_mut = c->max_mut;
        sp.F = F; sp.Fp = Fp; sp.jobs = d_jobs; sp.mat_off = d_moff; sp.status = d_status;
        sel_plan_kernel<<<(TG + 255) / 256, 256, 0, st>>>(sp);
        CUDA_TRY(cudaGetLastError());
        c->launches++;
    }
    int32_t status = 0;
    CUDA_TRY(cudaMemcpyAsync(&status, d_status, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *out_status = status;
    if (status) return 0;   // a parameter outside the tables fails the whole evaluation like the reference's ValueError
    // ---- the pass ----
    std::vector<int8_t> e(2 * (size_t)F);
    for (int f = 0; f < F; ++f) {   // likelihoods.py:425-430: inclusive comparisons
        e[f] = c->freqs[f] <= c->min_freq ? 1 : 0;
        e[F + f] = c->freqs[f] >= 1.0 - c->min_freq ? 1 : 0;
    }
    CUDA_TRY(cudaMemcpyAsync(d_e, e.data(), e.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_stat, stat, (size_t)F * sizeof(double), cudaMemcpyHostToDevice, st));
    {
        const long long n1 = (long long)n_internal * nodeN, n2 = (long long)L * Fp;
        sel_fill_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(cond, n1, 1.0);   // reset_likes_ones
        sel_fill_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(asc, n2, 1.0);
        CUDA_TRY(cudaGetLastError());
        c->launches += 2;
    }
    for (int b = 0; b < B; ++b) {
        const double* V;
        if (P.br_leaf_h[b] >= 0) {
            const long long n = (long long)nodeN;
            sel_leaf_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P.E + (size_t)P.br_leaf_h[b] * P.R_pad * Fp, c->use_v4 ? 1 : 0, L, F, Fp, d_e, d_e + F, leafbuf);
            CUDA_TRY(cudaGetLastError());
            c->launches++;
            V = leafbuf;
        } else {
            V = cond + (size_t)node_buf[b] * nodeN;
        }
        const int parent = P.br_parent_h[b];
        const bool do_asc = parent == B;
        double* par = cond + (size_t)node_buf[parent] * nodeN;
        const int g0 = S.br_first[b], ng = S.br_groups[b];
        const int32_t* roff = S.g_row_off + g0 + b;            // this branch's ng + 1 offsets
        const int32_t* rows = S.g_rows + (size_t)b * R3;
        for (int k0 = 0; k0 < ng; k0 += pool_blocks) {         // groups in chunks that fit the pool (stream order protects the reuse)
            const int nk = std::min(pool_blocks, ng - k0);
            if (fused) {
                // one pass: interpolate, check and apply -- the matrices never leave the SM
                ApplyArgs ap{roff + k0, rows, V, par, do_asc ? trans : nullptr, 0};
                if (int rc = launch_interp_apply(c, d_jobs + g0 + k0, (size_t)nk, st, ap)) return rc;
                continue;
            }
            if (int rc = launch_interp(c, d_jobs + g0 + k0, (size_t)nk, pool, st, false)) return rc;
            CUDA_TRY(launch_sel_matvec(nk, st, pool, d_moff + g0 + k0, roff + k0, rows, V, par, do_asc ? trans : nullptr, F, Fp, 0));
            c->launches++;
        }
        if (do_asc) {
            const long long n = (long long)L * Fp;
            sel_asc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(trans, asc, L, Fp);
            CUDA_TRY(cudaGetLastError());
            c->launches++;
        }
    }
    sel_root_kernel<<<(unsigned)(((long long)L * 32 + 255) / 256), 256, 0, st>>>(cond + (size_t)node_buf[B] * nodeN, asc, d_stat, L, F, Fp, d_out, d_out + L);
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    CUDA_TRY(cudaMemcpyAsync(out_loglikes, d_out, (size_t)L * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(out_log_asc, d_out + L, (size_t)L * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
