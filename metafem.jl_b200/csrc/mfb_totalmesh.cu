// mfb_totalmesh.cu -- first-order geometry tables of a mesh on the device.
//
// Reference: construct_TotalMesh_3D (src/mesh/ref_geometry/002_Initialization.jl:113-217) builds, from vertices and block
// (element) connectivity, the segment and face tables and the block -> segment / face incidences through its GPU hash table
// FEM_Dict (src/misc/06_GPU_Dict.jl:2-226); get_BoundaryMesh (:285-290) picks the faces referenced by one block, and the
// facet allocation of mesh_Classical needs, per boundary face, its host block and local face number (specify_eindex,
// src/mesh/unstructured_mesh/3_InitializeMesh.jl:165-178). Two numbering modes:
//   MFB_NUMBERING_SORTED     segment / face IDs = rank of their key among all keys (one radix sort + unique): deterministic and
//                            locality-preserving -- what a B200-native set-up wants;
//   MFB_NUMBERING_REFERENCE  the reference's order: IDs handed out pass by pass "in ascending hash-slot order" (:153-158,
//                            :181-186). The hash table is reproduced with its probing / chaining / growth policy and keys are
//                            inserted SEQUENTIALLY in array order by a one-thread kernel -- one legal outcome of the reference's
//                            racing atomic_cas insertion and the one the oracle (oracle/femdict.py) defines as "the reference's
//                            numbering" (SURVEY Appendix E). Meant for parity work: it is serial (~ 1 us per key).
// Outputs follow the reference's tables: 1-based Int32 IDs, column-major.
#include <thrust/copy.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/scan.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#include "mfb_internal.h"

namespace {
constexpr int TPB = 256;
inline unsigned nblk(int64_t n) { return (unsigned)((n + TPB - 1) / TPB); }
typedef unsigned long long u64;

// 002_Initialization.jl:1-8 (1-based in the reference; 0-based here)
struct Topo3 {
    int vpb, nsb, nfb, vpf;
    int bsv[12][2];
    int bfs[6][4];
};
Topo3 make_topo(int vpb) {
    Topo3 T;
    memset(&T, 0, sizeof(T));
    if (vpb == 4) {
        const int sv[6][2] = {{1, 2}, {2, 3}, {3, 1}, {1, 4}, {2, 4}, {3, 4}};
        const int fs[4][3] = {{1, 2, 3}, {1, 5, 4}, {2, 6, 5}, {3, 4, 6}};
        T.vpb = 4; T.nsb = 6; T.nfb = 4; T.vpf = 3;
        for (int i = 0; i < 6; ++i) for (int k = 0; k < 2; ++k) T.bsv[i][k] = sv[i][k] - 1;
        for (int i = 0; i < 4; ++i) for (int k = 0; k < 3; ++k) T.bfs[i][k] = fs[i][k] - 1;
    } else {
        const int sv[12][2] = {{1, 2}, {2, 3}, {3, 4}, {4, 1}, {1, 5}, {2, 6}, {3, 7}, {4, 8}, {5, 6}, {6, 7}, {7, 8}, {8, 5}};
        const int fs[6][4] = {{1, 2, 3, 4}, {1, 6, 9, 5}, {2, 7, 10, 6}, {3, 8, 11, 7}, {4, 8, 12, 5}, {9, 10, 11, 12}};
        T.vpb = 8; T.nsb = 12; T.nfb = 6; T.vpf = 4;
        for (int i = 0; i < 12; ++i) for (int k = 0; k < 2; ++k) T.bsv[i][k] = sv[i][k] - 1;
        for (int i = 0; i < 6; ++i) for (int k = 0; k < 4; ++k) T.bfs[i][k] = fs[i][k] - 1;
    }
    return T;
}

// I4I30I30_To_UI64 (06_GPU_Dict.jl:227-238)
__host__ __device__ inline u64 pack_key(u64 x, u64 y, u64 z) { return ((x + 1) << 60) + ((y & 0x3fffffffull) << 30) + (z & 0x3fffffffull); }

// segment key of (block, local segment): (max vertex, the other vertex) (:140-150)
__global__ void k_segment_keys(const int* conn, int64_t nb, Topo3 T, int pos, u64* keys, int* vmax, int* vnext) {
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const int v0 = conn[b * T.vpb + T.bsv[pos][0]], v1 = conn[b * T.vpb + T.bsv[pos][1]];
    const int mx = v1 > v0 ? v1 : v0, nx = v1 > v0 ? v0 : v1;          // findmax takes the FIRST maximum
    keys[b] = pack_key(0, (u64)mx, (u64)nx);
    vmax[b] = mx; vnext[b] = nx;
}
// face key of (block, local face): (max segment, its neighbour in the direction of the larger neighbour) (:167-180)
__global__ void k_face_keys3(const int* bseg /*[nsb][nb]*/, int64_t nb, Topo3 T, int pos, u64* keys, int* maxpos, unsigned char* forward) {
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= nb) return;
    int s[4];
    for (int k = 0; k < T.vpf; ++k) s[k] = bseg[(int64_t)T.bfs[pos][k] * nb + b];
    int mp = 0;
    for (int k = 1; k < T.vpf; ++k) if (s[k] > s[mp]) mp = k;
    const int pp = (mp + T.vpf - 1) % T.vpf, np = (mp + 1) % T.vpf;
    const bool fw = s[np] >= s[pp];
    keys[b] = pack_key(0, (u64)s[mp], (u64)s[fw ? np : pp]);
    maxpos[b] = mp; forward[b] = fw ? 1 : 0;
}
__global__ void k_rank_keys(const u64* keys, int64_t n, const u64* uniq, int64_t n_uniq, int* ids) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const u64 key = keys[t];
    int64_t lo = 0, hi = n_uniq - 1;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (uniq[mid] < key) lo = mid + 1; else hi = mid; }
    ids[t] = (int)lo + 1;
}
__global__ void k_fill_segments(const int* ids, const int* vmax, const int* vnext, int64_t nb, int* seg_v /*[2][ns] col-major*/) {
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const int64_t s = ids[b] - 1;
    seg_v[2 * s + 0] = vmax[b];            // every block that holds the segment writes the same pair
    seg_v[2 * s + 1] = vnext[b];
}
// face_segment_IDs / face_vertex_IDs of the faces met in this pass: walk the face's segments from the largest one in the
// direction of its larger neighbour; vertex i is the end of segment i that segment i+1 shares (:188-213)
__global__ void k_fill_faces(const int* bseg, const int* seg_v, const int* ids, const int* maxpos, const unsigned char* forward,
                             int64_t nb, Topo3 T, int pos, int* f_v, int* f_s) {
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= nb) return;
    int s[4];
    for (int k = 0; k < T.vpf; ++k) s[k] = bseg[(int64_t)T.bfs[pos][k] * nb + b];
    const int64_t f = ids[b] - 1;
    const int step = forward[b] ? 1 : T.vpf - 1;
    int p = maxpos[b];
    for (int i = 0; i < T.vpf; ++i) {
        const int cur = s[p], nxt = s[(p + step) % T.vpf];
        f_s[T.vpf * f + i] = cur;
        const int a = seg_v[2 * (int64_t)(cur - 1)], a2 = seg_v[2 * (int64_t)(cur - 1) + 1];
        const bool is_first = (a == seg_v[2 * (int64_t)(nxt - 1)]) || (a == seg_v[2 * (int64_t)(nxt - 1) + 1]);
        f_v[T.vpf * f + i] = is_first ? a : a2;
        p = (p + step) % T.vpf;
    }
}
__global__ void k_count_refs(const int* bface, int64_t n, int* cnt) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) atomicAdd(cnt + bface[t] - 1, 1);
}
struct IsOne {
    const int* cnt;
    __device__ bool operator()(int f) const { return cnt[f] == 1; }
};
__global__ void k_boundary_hosts(const int* bface /*[nfb][nb]*/, const int* cnt, const int* brank /*exclusive scan of (cnt == 1)*/,
                                 int64_t nb, int nfb, int* b_el, int* b_eidx) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nb * nfb) return;
    const int pos = (int)(t / nb);
    const int64_t b = t - (int64_t)pos * nb;
    const int f = bface[t] - 1;
    if (cnt[f] == 1) { b_el[brank[f]] = (int)b + 1; b_eidx[brank[f]] = pos + 1; }
}

// ---- FEM_Dict, sequential insertion (06_GPU_Dict.jl) ------------------------------------------------------------------------
__device__ __forceinline__ u64 wang64(u64 a) {                       // GPU_hash_64_64 (:2-11)
    a = ~a + (a << 21);
    a = a ^ (a >> 24);
    a = a + (a << 3) + (a << 8);
    a = a ^ (a >> 14);
    a = a + (a << 2) + (a << 4);
    a = a ^ (a >> 28);
    a = a + (a << 31);
    return a;
}
__device__ __forceinline__ int trunc_id(u64 prev, long long size) { return (int)(((unsigned)prev) & (unsigned)(size - 1)) + 1; }   // update_TruncID (:124)
long long dict_size(long long x) {                                   // _DictSize (:13)
    if ((double)x < 16.0 * 2.0 / 3.0) return 16;
    const long long need = (3 * x + 1) / 2;
    long long s = 1;
    while (s < need) s <<= 1;
    return s;
}
struct DictArrays {
    u64 *keys, *hashs;
    int *hinit, *hprev, *hnext, *vals;
    long long size;
};
// dict_SetID! (:125-162) for new_keys[0..m) IN ORDER: one thread
__global__ void k_dict_set_sequential(DictArrays D, long long m, const u64* new_keys, int* new_ids) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (long long t = 0; t < m; ++t) {
        const u64 key = new_keys[t];
        const u64 h = wang64(key);
        const int start = trunc_id(h, D.size);
        int cur = D.hinit[start - 1] == 0 ? start : D.hinit[start - 1];
        int last_front = 0;
        for (;;) {
            const u64 local = D.keys[cur - 1];
            if (local == 0) {
                D.keys[cur - 1] = key;
                D.hashs[cur - 1] = h;
                if (last_front != 0) { D.hprev[cur - 1] = last_front; D.hnext[last_front - 1] = cur; }
                else D.hinit[start - 1] = cur;
                break;
            } else if (local == key) {
                break;
            } else if (trunc_id(wang64(local), D.size) == start) {
                if (D.hnext[cur - 1] == 0) { last_front = cur; cur = trunc_id((u64)cur, D.size); }
                else cur = D.hnext[cur - 1];
            } else {
                cur = trunc_id((u64)cur, D.size);
            }
        }
        new_ids[t] = cur;
    }
}
struct Occupied {
    const u64* keys;
    __device__ bool operator()(int slot) const { return keys[slot] != 0; }
};
__global__ void k_gather_slots(const int* slots, long long n, const u64* keys, const int* vals, u64* okeys, int* ovals) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < n) { okeys[t] = keys[slots[t]]; ovals[t] = vals[slots[t]]; }
}
__global__ void k_scatter_vals(const int* mapped /*1-based slots*/, const int* vals, long long n, int* dvals) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < n) dvals[mapped[t] - 1] = vals[t];
}
// new IDs for the stored keys that have none yet, in ascending slot order (:153-158)
__global__ void k_flag_new(const u64* keys, const int* vals, long long size, int* flag) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < size) flag[t] = (keys[t] != 0 && vals[t] == 0) ? 1 : 0;
}
__global__ void k_assign_new(const int* flag, const int* rank, long long size, int base, int* vals) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < size && flag[t]) vals[t] = base + rank[t] + 1;
}
__global__ void k_vals_of_slots(const int* slots, const int* vals, long long n, int* out) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t < n) out[t] = vals[slots[t] - 1];
}

struct DeviceDict {                        // FEM_Dict with Int32 values
    DevBuf<u64> keys, hashs;
    DevBuf<int> hinit, hprev, hnext, vals;
    long long size = 0, count = 0;
};
}  // namespace

#define LAUNCH(kernel, grid, block, ...)                          \
    do {                                                          \
        kernel<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__); \
        ctx->launches++;                                          \
    } while (0)

namespace {
int dict_alloc(mfb_ctx* ctx, DeviceDict& d, long long n) {
    d.keys.release(); d.hashs.release(); d.hinit.release(); d.hprev.release(); d.hnext.release(); d.vals.release();
    MFB_CUDA(d.keys.alloc(n)); MFB_CUDA(d.hashs.alloc(n)); MFB_CUDA(d.hinit.alloc(n)); MFB_CUDA(d.hprev.alloc(n));
    MFB_CUDA(d.hnext.alloc(n)); MFB_CUDA(d.vals.alloc(n));
    MFB_CUDA(cudaMemsetAsync(d.keys.p, 0, n * sizeof(u64), ctx->stream));
    MFB_CUDA(cudaMemsetAsync(d.hashs.p, 0, n * sizeof(u64), ctx->stream));
    MFB_CUDA(cudaMemsetAsync(d.hinit.p, 0, n * sizeof(int), ctx->stream));
    MFB_CUDA(cudaMemsetAsync(d.hprev.p, 0, n * sizeof(int), ctx->stream));
    MFB_CUDA(cudaMemsetAsync(d.hnext.p, 0, n * sizeof(int), ctx->stream));
    MFB_CUDA(cudaMemsetAsync(d.vals.p, 0, n * sizeof(int), ctx->stream));
    d.size = n;
    return MFB_OK;
}
DictArrays arrays(DeviceDict& d) { return DictArrays{d.keys.p, d.hashs.p, d.hinit.p, d.hprev.p, d.hnext.p, d.vals.p, d.size}; }

// FEM_Dict_SetID! (:45-93): grow to _DictSize(new + stored) when that differs from the current size (stored keys re-inserted in
// ascending slot order, values moved along), then insert the new keys; slots_out[t] = slot (1-based) of new_keys[t].
// Afterwards the stored keys without a value get IDs base+1.. in ascending slot order and ids_out[t] = value of new_keys[t].
int dict_set_and_number(mfb_ctx* ctx, DeviceDict& d, const u64* new_keys, long long m, int* slots_tmp, int* ids_out, int* n_ids) {
    auto pol = thrust::cuda::par.on(ctx->stream);
    const long long est = dict_size(m + d.count);
    if (est != d.size) {
        DevBuf<int> occ, ovals, mapped;
        DevBuf<u64> okeys;
        const long long nc = d.count;
        if (nc > 0) {
            MFB_CUDA(occ.alloc(nc)); MFB_CUDA(ovals.alloc(nc)); MFB_CUDA(okeys.alloc(nc)); MFB_CUDA(mapped.alloc(nc));
            thrust::device_ptr<int> op(occ.p);
            Occupied pred{d.keys.p};
            thrust::copy_if(pol, thrust::counting_iterator<int>(0), thrust::counting_iterator<int>((int)d.size), op, pred);
            LAUNCH(k_gather_slots, nblk(nc), TPB, occ.p, nc, d.keys.p, d.vals.p, okeys.p, ovals.p);
            MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        MFB_TRY(dict_alloc(ctx, d, est));
        if (nc > 0) {
            LAUNCH(k_dict_set_sequential, 1, 1, arrays(d), nc, okeys.p, mapped.p);
            LAUNCH(k_scatter_vals, nblk(nc), TPB, mapped.p, ovals.p, nc, d.vals.p);
            MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
    LAUNCH(k_dict_set_sequential, 1, 1, arrays(d), m, new_keys, slots_tmp);
    DevBuf<int> flag, rank;
    MFB_CUDA(flag.alloc(d.size)); MFB_CUDA(rank.alloc(d.size));
    LAUNCH(k_flag_new, nblk(d.size), TPB, d.keys.p, d.vals.p, d.size, flag.p);
    thrust::device_ptr<int> fp(flag.p), rp(rank.p);
    thrust::exclusive_scan(pol, fp, fp + d.size, rp);
    int last_rank = 0, last_flag = 0;
    MFB_CUDA(cudaMemcpyAsync(&last_rank, rank.p + d.size - 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MFB_CUDA(cudaMemcpyAsync(&last_flag, flag.p + d.size - 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    const int n_new = last_rank + last_flag;
    LAUNCH(k_assign_new, nblk(d.size), TPB, flag.p, rank.p, d.size, *n_ids, d.vals.p);
    LAUNCH(k_vals_of_slots, nblk(m), TPB, slots_tmp, d.vals.p, m, ids_out);
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_ids += n_new;
    d.count += n_new;
    return MFB_OK;
}
}  // namespace

struct TotalMesh {
    int vpb = 0, nsb = 0, nfb = 0, vpf = 0;
    int64_t nb = 0, ns = 0, nf = 0, nbf = 0;
    DevBuf<int> bseg, bface;          // [nsb][nb], [nfb][nb]  (the reference's [nsb, nb] column-major is [nb][nsb]: transposed on export)
    DevBuf<int> seg_v, f_v, f_s;      // [ns][2], [nf][vpf], [nf][vpf]  == column-major [2, ns], [vpf, nf]
    DevBuf<int> bf_id, bf_el, bf_eidx;
};
static TotalMesh*& total_mesh_slot(mfb_ctx* ctx) {
    static std::map<mfb_ctx*, TotalMesh*> table;
    return table[ctx];
}
void mfb_totalmesh_free(mfb_ctx* ctx) {
    TotalMesh*& t = total_mesh_slot(ctx);
    delete t;
    t = nullptr;
}

namespace {
__global__ void k_export_block_table(const int* src /*[rows][nb]*/, int rows, int64_t nb, int* dst /*[nb][rows] = col-major [rows, nb]*/) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nb * rows) return;
    const int r = (int)(t % rows);
    const int64_t b = t / rows;
    dst[t] = src[(int64_t)r * nb + b];
}
}  // namespace

extern "C" int mfb_total_mesh_build(mfb_ctx* ctx, int64_t n_vert, int vpb, int64_t n_blocks, const int32_t* connections, int numbering,
                                    int64_t* n_segments, int64_t* n_faces, int64_t* n_boundary_faces) {
    if (!ctx) return MFB_ERR_ARG;
    MFB_REQUIRE(n_vert > 0 && n_blocks > 0 && connections && (vpb == 4 || vpb == 8), MFB_ERR_ARG, "mfb_total_mesh_build: tets (4) or hexes (8)");
    MFB_REQUIRE(numbering == MFB_NUMBERING_SORTED || numbering == MFB_NUMBERING_REFERENCE, MFB_ERR_ARG, "unknown numbering mode");
    MFB_REQUIRE(n_vert < (1ll << 30) && n_blocks * 12 < (1ll << 30), MFB_ERR_ARG, "ids exceed the 30-bit key fields of I4I30I30");
    MFB_CUDA(cudaSetDevice(ctx->device));
    mfb_totalmesh_free(ctx);
    TotalMesh* M = total_mesh_slot(ctx) = new TotalMesh();
    const Topo3 T = make_topo(vpb);
    const int64_t nb = n_blocks;
    M->vpb = vpb; M->nsb = T.nsb; M->nfb = T.nfb; M->vpf = T.vpf; M->nb = nb;
    auto pol = thrust::cuda::par.on(ctx->stream);
    DevBuf<int> conn, vmax, vnext, ids, slots, maxpos;
    DevBuf<unsigned char> fwd;
    DevBuf<u64> keys;
    MFB_CUDA(conn.alloc(nb * vpb));
    MFB_TRY(mfb_stage_in(ctx, connections, nb * vpb * sizeof(int), conn.p));        // [vpb, nb] column-major == [nb][vpb]
    MFB_CUDA(M->bseg.alloc((int64_t)T.nsb * nb)); MFB_CUDA(M->bface.alloc((int64_t)T.nfb * nb));
    const int passes = T.nsb > T.nfb ? T.nsb : T.nfb;
    MFB_CUDA(keys.alloc((int64_t)passes * nb)); MFB_CUDA(vmax.alloc((int64_t)T.nsb * nb)); MFB_CUDA(vnext.alloc((int64_t)T.nsb * nb));
    MFB_CUDA(maxpos.alloc((int64_t)T.nfb * nb)); MFB_CUDA(fwd.alloc((int64_t)T.nfb * nb)); MFB_CUDA(slots.alloc(nb));
    // ---------------- segments (:137-165) ----------------
    for (int p = 0; p < T.nsb; ++p)
        LAUNCH(k_segment_keys, nblk(nb), TPB, conn.p, nb, T, p, keys.p + (int64_t)p * nb, vmax.p + (int64_t)p * nb, vnext.p + (int64_t)p * nb);
    int n_seg = 0;
    if (numbering == MFB_NUMBERING_SORTED) {
        DevBuf<u64> uniq;
        const int64_t tot = (int64_t)T.nsb * nb;
        MFB_CUDA(uniq.alloc(tot));
        MFB_CUDA(cudaMemcpyAsync(uniq.p, keys.p, tot * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
        thrust::device_ptr<u64> up(uniq.p);
        thrust::sort(pol, up, up + tot);
        n_seg = (int)(thrust::unique(pol, up, up + tot) - up);
        LAUNCH(k_rank_keys, nblk(tot), TPB, keys.p, tot, uniq.p, (int64_t)n_seg, M->bseg.p);
    } else {
        DeviceDict d;
        MFB_TRY(dict_alloc(ctx, d, dict_size(16)));
        for (int p = 0; p < T.nsb; ++p)
            MFB_TRY(dict_set_and_number(ctx, d, keys.p + (int64_t)p * nb, nb, slots.p, M->bseg.p + (int64_t)p * nb, &n_seg));
    }
    M->ns = n_seg;
    MFB_CUDA(M->seg_v.alloc(2 * (int64_t)n_seg));
    for (int p = 0; p < T.nsb; ++p)
        LAUNCH(k_fill_segments, nblk(nb), TPB, M->bseg.p + (int64_t)p * nb, vmax.p + (int64_t)p * nb, vnext.p + (int64_t)p * nb, nb, M->seg_v.p);
    // ---------------- faces (:167-215) ----------------
    for (int p = 0; p < T.nfb; ++p)
        LAUNCH(k_face_keys3, nblk(nb), TPB, M->bseg.p, nb, T, p, keys.p + (int64_t)p * nb, maxpos.p + (int64_t)p * nb, fwd.p + (int64_t)p * nb);
    int n_face = 0;
    if (numbering == MFB_NUMBERING_SORTED) {
        DevBuf<u64> uniq;
        const int64_t tot = (int64_t)T.nfb * nb;
        MFB_CUDA(uniq.alloc(tot));
        MFB_CUDA(cudaMemcpyAsync(uniq.p, keys.p, tot * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
        thrust::device_ptr<u64> up(uniq.p);
        thrust::sort(pol, up, up + tot);
        n_face = (int)(thrust::unique(pol, up, up + tot) - up);
        LAUNCH(k_rank_keys, nblk(tot), TPB, keys.p, tot, uniq.p, (int64_t)n_face, M->bface.p);
    } else {
        DeviceDict d;
        MFB_TRY(dict_alloc(ctx, d, dict_size(16)));
        for (int p = 0; p < T.nfb; ++p)
            MFB_TRY(dict_set_and_number(ctx, d, keys.p + (int64_t)p * nb, nb, slots.p, M->bface.p + (int64_t)p * nb, &n_face));
    }
    M->nf = n_face;
    MFB_CUDA(M->f_v.alloc((int64_t)T.vpf * n_face)); MFB_CUDA(M->f_s.alloc((int64_t)T.vpf * n_face));
    for (int p = 0; p < T.nfb; ++p)
        LAUNCH(k_fill_faces, nblk(nb), TPB, M->bseg.p, M->seg_v.p, M->bface.p + (int64_t)p * nb, maxpos.p + (int64_t)p * nb,
               fwd.p + (int64_t)p * nb, nb, T, p, M->f_v.p, M->f_s.p);
    // ---------------- boundary faces and their hosts (get_BoundaryMesh :285-290, specify_eindex) ----------------
    DevBuf<int> cnt, flag, rank;
    MFB_CUDA(cnt.alloc(n_face)); MFB_CUDA(flag.alloc(n_face)); MFB_CUDA(rank.alloc(n_face));
    MFB_CUDA(cudaMemsetAsync(cnt.p, 0, n_face * sizeof(int), ctx->stream));
    LAUNCH(k_count_refs, nblk((int64_t)T.nfb * nb), TPB, M->bface.p, (int64_t)T.nfb * nb, cnt.p);
    MFB_CUDA(M->bf_id.alloc(n_face));
    thrust::device_ptr<int> bp(M->bf_id.p);
    IsOne one{cnt.p};
    M->nbf = thrust::copy_if(pol, thrust::counting_iterator<int>(0), thrust::counting_iterator<int>(n_face), bp, one) - bp;
    {
        // rank of each boundary face among the boundary faces = exclusive scan of (cnt == 1)
        thrust::device_ptr<int> cp(cnt.p), rp(rank.p), fp(flag.p);
        thrust::transform(pol, cp, cp + n_face, fp, [] __device__(int c) { return c == 1 ? 1 : 0; });
        thrust::exclusive_scan(pol, fp, fp + n_face, rp);
    }
    if (M->nbf > 0) {
        MFB_CUDA(M->bf_el.alloc(M->nbf)); MFB_CUDA(M->bf_eidx.alloc(M->nbf));
        LAUNCH(k_boundary_hosts, nblk((int64_t)T.nfb * nb), TPB, M->bface.p, cnt.p, rank.p, nb, T.nfb, M->bf_el.p, M->bf_eidx.p);
    }
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    MFB_CUDA(cudaGetLastError());
    if (n_segments) *n_segments = M->ns;
    if (n_faces) *n_faces = M->nf;
    if (n_boundary_faces) *n_boundary_faces = M->nbf;
    return MFB_OK;
}

extern "C" int mfb_total_mesh_get(mfb_ctx* ctx, int32_t* segment_vertex_IDs, int32_t* block_segment_IDs, int32_t* face_vertex_IDs,
                                  int32_t* face_segment_IDs, int32_t* block_face_IDs, int32_t* boundary_face_IDs,
                                  int32_t* boundary_face_block, int32_t* boundary_face_eindex) {
    if (!ctx) return MFB_ERR_ARG;
    TotalMesh* M = total_mesh_slot(ctx);
    MFB_REQUIRE(M != nullptr, MFB_ERR_STATE, "mfb_total_mesh_get: call mfb_total_mesh_build first");
    MFB_CUDA(cudaSetDevice(ctx->device));
    if (segment_vertex_IDs) MFB_TRY(mfb_stage_out(ctx, M->seg_v.p, 2 * M->ns * sizeof(int), segment_vertex_IDs));
    if (face_vertex_IDs) MFB_TRY(mfb_stage_out(ctx, M->f_v.p, M->vpf * M->nf * sizeof(int), face_vertex_IDs));
    if (face_segment_IDs) MFB_TRY(mfb_stage_out(ctx, M->f_s.p, M->vpf * M->nf * sizeof(int), face_segment_IDs));
    DevBuf<int> tmp;
    if (block_segment_IDs) {
        MFB_CUDA(tmp.alloc((int64_t)M->nsb * M->nb));
        LAUNCH(k_export_block_table, nblk((int64_t)M->nsb * M->nb), TPB, M->bseg.p, M->nsb, M->nb, tmp.p);
        MFB_TRY(mfb_stage_out(ctx, tmp.p, (int64_t)M->nsb * M->nb * sizeof(int), block_segment_IDs));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (block_face_IDs) {
        MFB_CUDA(tmp.alloc((int64_t)M->nfb * M->nb));
        LAUNCH(k_export_block_table, nblk((int64_t)M->nfb * M->nb), TPB, M->bface.p, M->nfb, M->nb, tmp.p);
        MFB_TRY(mfb_stage_out(ctx, tmp.p, (int64_t)M->nfb * M->nb * sizeof(int), block_face_IDs));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (M->nbf > 0) {
        if (boundary_face_IDs) {
            DevBuf<int> one;
            MFB_CUDA(one.alloc(M->nbf));
            auto pol = thrust::cuda::par.on(ctx->stream);
            thrust::device_ptr<int> s(M->bf_id.p), d(one.p);
            thrust::transform(pol, s, s + M->nbf, d, [] __device__(int f) { return f + 1; });      // 1-based, ascending
            MFB_TRY(mfb_stage_out(ctx, one.p, M->nbf * sizeof(int), boundary_face_IDs));
            MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        if (boundary_face_block) MFB_TRY(mfb_stage_out(ctx, M->bf_el.p, M->nbf * sizeof(int), boundary_face_block));
        if (boundary_face_eindex) MFB_TRY(mfb_stage_out(ctx, M->bf_eidx.p, M->nbf * sizeof(int), boundary_face_eindex));
    }
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MFB_OK;
}
