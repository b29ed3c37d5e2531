// mfb_internal.h -- context and helpers shared by the translation units of libmetafem_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/metafem_b200.h"
#include "mfb_skeleton.cuh"

#define MFB_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (call);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                       ":" + std::to_string(__LINE__) + ")";                                \
            return MFB_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define MFB_REQUIRE(cond, code, msg)  \
    do {                              \
        if (!(cond)) {                \
            ctx->err = (msg);         \
            return (code);            \
        }                             \
    } while (0)

#define MFB_TRY(expr)                 \
    do {                              \
        int _s = (expr);              \
        if (_s < 0) return _s;        \
    } while (0)

// Owning device buffer: released on scope exit (also on the early returns of MFB_TRY / MFB_CUDA), move-only.
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t count) {
        if (count == n && p) return cudaSuccess;
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void**)&p, count * sizeof(T));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

// The TMA-ring SpMV copies whole stages of the value / column streams: the last stage of the last rows reads up to this many
// node-pair entries past the end of the arrays, so nodecol and every matrix value array carry that much slack.
constexpr int MFB_STREAM_PAD = 1024;

struct BlockKernels {
    int kind = 0, bg_ID = 0;
    cudaKernel_t lin = nullptr, nonlin = nullptr, eval = nullptr;
    std::vector<std::string> cp_vars, globals, qp_in, qp_out;
    int tpb = 128, smem = 0, has_nonlinear_K = 0;
};

struct BoundaryGroup {
    DevBuf<int> elem, face;  // 0-based host element and local face of each facet of the group, ordered by colour
    int64_t n = 0;
    // facets of one colour share no node through their host elements: one launch per colour makes the boundary scatter
    // conflict-free, hence bit-reproducible
    std::vector<int64_t> color_ptr;     // [n_colors + 1] into elem / face
    DevBuf<int> bemap;                  // [n][n_a*n_a] ENTRY (never a side slot) of every node pair of the facet's host element
};

struct Comm;       // mfb_dist.cu
struct MeshBuild;  // mfb_meshbuild.cu

// CUDA-event timers on the context's stream (bench.py roofline numbers); ids = MFB_T_*
enum { MFB_T_SPMV = 0, MFB_T_ASM_NONLINEAR = 1, MFB_T_ASM_LINEAR = 2, MFB_T_SOLVE = 3, MFB_T_ELEM_KERNEL = 4, MFB_T_HALO = 5,
       MFB_T_REDUCE = 6, MFB_T_PRECOND = 7, MFB_T_COUNT = 8 };
struct ProfEvents {
    std::vector<cudaEvent_t> start, stop;
};
// fine-grained timers (one pair of events per reduction group / interface exchange) perturb the stream they measure:
// they are recorded only at profile level 2
inline bool mfb_prof_detail(int id) { return id == 5 || id == 6; }

struct mfb_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err, compile_log;
    int64_t launches = 0;
    int sm_count = 148;

    // ---- mesh ----
    int n_a = 0, n_q = 0, n_faces = 0, n_qb = 0;
    int64_t n_el = 0, N = 0, n_facets = 0;
    DevBuf<int> conn_ref;     // [n_el][n_a] 0-based reference node ids
    DevBuf<int> conn;         // [n_el][n_a] internal node ids, elements in internal (Morton) order
    DevBuf<int> elem_order;   // elem_order[internal element] = reference element (0-based)
    DevBuf<int> elem_rank;    // inverse
    DevBuf<double> xyz;       // [N][3] internal order
    DevBuf<double> ref, wq;   // domain tables [4][n_q][n_a], [n_q]
    DevBuf<double> bref, bwq, btan;  // boundary tables per face
    DevBuf<int> facet_elem, facet_face;  // [n_facets] 0-based
    std::map<int, BoundaryGroup> groups;

    // ---- numbering / pattern ----
    int n_var = 0, L1 = 1, n_blocks = 0;
    std::vector<int> sparse_mapping;  // [n_blocks][2]
    std::vector<int> block_of;        // [n_var*n_var] -> block number or -1
    DevBuf<int> perm;       // perm[ref node] = internal node
    DevBuf<int> iperm;      // iperm[internal] = ref node
    DevBuf<int> nodeptr;    // [N+1] node graph CSR (internal)
    DevBuf<int> nodecol;    // [U]
    DevBuf<int> emap;       // [n_el][n_a*n_a] scatter target of every element node pair: an entry (< U) or a side slot (U + s)
    int64_t U = 0;          // sparse_unitsize
    // ---- deterministic scatter (pairwise accumulators) ----
    // An accumulator that starts at zero and receives at most TWO atomic adds holds the same bits whatever their order
    // (a + b == b + a). The m contributions an entry receives from the domain elements are therefore dealt, by their rank in
    // the (fixed) element order, to ceil(m/2) accumulators: ranks 0,1 -> the entry itself, ranks 2,3 -> side slot 0, ... ;
    // a combine pass then folds the side slots into the entry in slot order. Same for the residual (per node).
    bool deterministic = true;
    int64_t S = 0;                  // side slots of the matrix (blocks of n_var^2 values behind the U entries)
    DevBuf<int> slot_entry;         // [S] entry each side slot belongs to (slots of one entry are consecutive)
    DevBuf<int> ext_entry, ext_first, ext_count;   // entries that own side slots: entry, first slot, number of slots
    int64_t n_ext = 0;
    DevBuf<int> rmap;               // [n_el][n_a] residual scatter target of every element node: a node (< N) or a side slot (N + s)
    int64_t SR = 0;                 // side slots of the residual (n_var values each)
    DevBuf<int> rext_node, rext_first, rext_count;
    int64_t n_rext = 0;
    DevBuf<int> lin_entries;        // unique entries the boundary-group elements touch (K_total += K_linear there), lazily built
    int64_t n_lin_entries = -1;
    DevBuf<int> ref_pos;    // [U] position of entry inside its row, ranked by reference column id (lazy)

    // ---- state (internal layout) ----
    DevBuf<double> x, dx, x_star;  // [L1][N][n_var]
    DevBuf<double> residue;        // [N][n_var]
    DevBuf<double> K_linear, K_total;  // [U][n_var*n_var]
    DevBuf<double> delta;          // last solve result [N][n_var]
    bool have_delta = false;
    std::map<std::string, DevBuf<double>> fields;  // internal order [N]
    std::map<std::string, double> globals;
    std::map<std::string, DevBuf<double>> qp;      // INTEGRATION_POINT_VAR arrays [n_q, n_el], reference element order
    struct J2State { std::string e[6], ep[6]; };
    std::map<std::string, J2State> j2;
    DevBuf<unsigned long long> qp_counter;         // yielded-point counter of the J2 return map

    // ---- kernels ----
    cudaLibrary_t lib = nullptr;
    std::vector<BlockKernels> blocks;
    bool any_nonlinear_K = false;

    // ---- krylov workspace ----
    std::vector<DevBuf<double>> work;
    DevBuf<double> jac, scal;   // Jacobi vector, device scalars
    DevBuf<double> ksc;         // device-resident Krylov recurrence scalars
    DevBuf<unsigned> red_counter;        // ticket counter of the fused reductions (last-block detection)
    DevBuf<unsigned char> kprog;         // fused vector programs of the running solve (device copy)
    DevBuf<double> Ks;          // Jacobi-scaled copy of K_total for the solve (kept across solves)
    cudaEvent_t lag_ev[2] = {nullptr, nullptr};   // convergence read-back events (the host lags one outer iteration behind)
    double* h_scal = nullptr;   // pinned host mirror of scal

    // ---- staging ----
    DevBuf<unsigned char> stage;

    Comm* comm = nullptr;
    MeshBuild* meshbuild = nullptr;   // tables of the last mfb_mesh_build_second_order
    DevBuf<unsigned char> owned;   // [N] internal order: 1 if this rank owns the node (all 1 on a single GPU)
    double n_global_nodes = 0;     // number of distinct nodes over all ranks (0 = single GPU: use N)
    DevBuf<long long> gid;         // [N] internal order: global reference node id, 0-based (seeds the shadow vectors)

    // ---- profiling ----
    bool profile = false;
    int profile_level = 1;
    std::vector<cudaEvent_t> event_pool;   // recycled by mfb_profile_get
    ProfEvents prof[MFB_T_COUNT];
    double prof_ms[MFB_T_COUNT] = {0};
    int64_t prof_n[MFB_T_COUNT] = {0};
};

struct ProfScope {
    mfb_ctx* c;
    int id;
    bool on;
    static cudaEvent_t get(mfb_ctx* c) {
        cudaEvent_t e;
        if (!c->event_pool.empty()) { e = c->event_pool.back(); c->event_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    }
    ProfScope(mfb_ctx* ctx, int id_) : c(ctx), id(id_), on(ctx->profile && (ctx->profile_level >= 2 || !mfb_prof_detail(id_))) {
        if (!on) return;
        cudaEvent_t e = get(c);
        cudaEventRecord(e, c->stream);
        c->prof[id].start.push_back(e);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEvent_t e = get(c);
        cudaEventRecord(e, c->stream);
        c->prof[id].stop.push_back(e);
    }
};

// staging: returns a device-readable pointer for `src` (device/unified/pinned pass through; pageable host is copied)
int mfb_stage_in(mfb_ctx* ctx, const void* src, size_t bytes, void* dev_dst);
int mfb_stage_out(mfb_ctx* ctx, const void* dev_src, size_t bytes, void* dst);

// mfb_pattern.cu
int mfb_build_permutation(mfb_ctx* ctx, const double* x1, const double* x2, const double* x3);
int mfb_build_pattern(mfb_ctx* ctx);
int mfb_to_internal(mfb_ctx* ctx, const double* ref_vec, double* int_vec, int levels);    // device ptrs
int mfb_to_reference(mfb_ctx* ctx, const double* int_vec, double* ref_vec, int levels);  // device ptrs
int mfb_field_to_internal(mfb_ctx* ctx, const double* ref_field, double* int_field);
int mfb_export_matrix(mfb_ctx* ctx, const double* Kint, double* Kref_dev);
int mfb_export_pattern(mfb_ctx* ctx, int* K_I, int* K_J, int* K_J_ptr, int* K_val_ids);  // device ptrs (nullable)
int mfb_export_sparse_ids(mfb_ctx* ctx, int block, int* out_dev);
int mfb_entry_map(mfb_ctx* ctx, const int* item_elem, int64_t n_items, int* out);   // [n_items][n_a^2] entries of the items' elements
int mfb_combine_matrix(mfb_ctx* ctx, double* K);      // fold the side slots of K into its entries
int mfb_combine_residue(mfb_ctx* ctx, double* res);

// ---- peer-memory mailboxes of the scalar-batch allreduce (mfb_dist.cu owns them; the reduction epilogues of mfb_krylov.cu
// publish into / read from them) ----
constexpr int P2P_MAXV = 24;
constexpr int P2P_MAXR = 16;
struct P2PSlot {
    double v[P2P_MAXV];
    unsigned long long seq;
    unsigned long long pad[7];
};
static_assert(sizeof(P2PSlot) == 256, "mailbox slot is 256 bytes");
struct P2PPeers {
    P2PSlot* box[P2P_MAXR];       // mailbox base of every rank: [2 parities][n_ranks slots]
};
struct P2PInfo {
    P2PPeers peers;
    int rank = 0, n_ranks = 1;
    unsigned long long seq = 0;   // sequence number of THIS round (already advanced)
    int* err = nullptr;
};
// true when the peer-memory allreduce is up: fills `out` and advances the round counter (one call per reduction launch)
bool mfb_p2p_next(mfb_ctx* ctx, P2PInfo* out);

// mfb_dist.cu
int mfb_node_ids_init(mfb_ctx* ctx, const unsigned char* owned_ref_dev, const long long* gid_ref_dev);
int mfb_halo_add(mfb_ctx* ctx, double* v, int nv);
int mfb_allreduce_sum(mfb_ctx* ctx, double* dev, int n);
// fused fold of per-block partials + allreduce over peer memory; returns 1 when unavailable (single GPU / NCCL fallback)
int mfb_reduce_allreduce(mfb_ctx* ctx, const double* partials, int nb, int k, double* out);
int mfb_p2p_check(mfb_ctx* ctx);
bool mfb_is_distributed(mfb_ctx* ctx);
void mfb_comm_free(mfb_ctx* ctx);

// mfb_meshbuild.cu
void mfb_meshbuild_free(mfb_ctx* ctx);
// mfb_totalmesh.cu
void mfb_totalmesh_free(mfb_ctx* ctx);
// mfb_ilu.cu
int mfb_ilu_factor(mfb_ctx* ctx, const double* A, int* n_levels);   // A: [U][n_var^2] block values (internal layout)
int mfb_ilu_apply(mfb_ctx* ctx, double* v);                         // v <- U^-1 L^-1 v, internal layout
void mfb_ilu_free(mfb_ctx* ctx);

// mfb_qp.cu
int mfb_qp_lookup(mfb_ctx* ctx, const std::string& name, double** p);   // creates the array (zeroed) on first use

// mfb_krylov.cu
int mfb_spmv_internal(mfb_ctx* ctx, const double* K, const double* x, double* y);
int mfb_spmv_kind(mfb_ctx* ctx);   // 0 = one warp per row (k_spmv_bsr), 1 = multi-row streams (k_spmv_mr)
bool mfb_sweep_level_mr(mfb_ctx* ctx, bool upper, const int* ptr, const int* col, const void* val, bool f32, const int* rowid,
                        const double* dinv, double* v, int w0, int w1);
int mfb_spmv_t_internal(mfb_ctx* ctx, const double* K, const double* x, double* y);   // y = K' x
