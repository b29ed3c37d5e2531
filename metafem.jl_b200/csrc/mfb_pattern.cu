// mfb_pattern.cu -- DOF numbering, locality permutation and CSR sparsity pattern on the device.
//
// Replaces assemble_Global_Variables! / assemble_SparseID! / assemble_KIJ!
// (reference src/solver/03_GlobalAssembly.jl:6-37,77-168) and sort_CUSPARSE_COO! / generate_J_ptr
// (src/misc/04_GPU_Utils.jl:87-118). The reference hashes all n_a^2*n_el (node,node) keys into an
// open-addressing table and sorts the resulting COO with cuSPARSE; here the node graph is built by
// one radix sort + unique of the packed keys, and the matrix is held as ONE node graph with
// n_var x n_var value blocks (all variable blocks of the reference share that graph:
// K_I = cp_i + dual_pos*N, K_J = cp_j + base_pos*N, 03_GlobalAssembly.jl:148-167).
// Reference-layout CSR arrays are produced on request for parity checks.
#include <thrust/binary_search.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/extrema.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/scan.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mfb_internal.h"

namespace {

constexpr int TPB = 256;
inline unsigned nblk(int64_t n) { return (unsigned)((n + TPB - 1) / TPB); }

__device__ __forceinline__ unsigned long long spread21(unsigned long long v) {   // 21 bits -> every third bit
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

// Morton key of the element centroid (coordinates in reference node order), quantised to 2^bits cells per axis
__global__ void k_elem_morton(const int* conn_ref, int n_a, int64_t n_el, const double* x1, const double* x2, const double* x3,
                              double lo0, double lo1, double lo2, double inv0, double inv1, double inv2,
                              unsigned long long* key, int* order) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n_el) return;
    double c0 = 0, c1 = 0, c2 = 0;
    for (int a = 0; a < n_a; ++a) {
        int g = conn_ref[e * n_a + a];
        c0 += x1[g]; c1 += x2[g]; c2 += x3[g];
    }
    const double s = 1.0 / n_a, top = 2097151.0;
    unsigned long long q0 = (unsigned long long)fmin(fmax((c0 * s - lo0) * inv0, 0.0), top);
    unsigned long long q1 = (unsigned long long)fmin(fmax((c1 * s - lo1) * inv1, 0.0), top);
    unsigned long long q2 = (unsigned long long)fmin(fmax((c2 * s - lo2) * inv2, 0.0), top);
    key[e] = spread21(q0) | (spread21(q1) << 1) | (spread21(q2) << 2);
    order[e] = (int)e;
}

// first touch in the NEW element order: key[node] = min over (new element index, local node)
__global__ void k_first_touch_ordered(const int* conn_ref, const int* elem_order, int n_a, int64_t n_el, unsigned long long* key) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_el * n_a) return;
    int64_t ne = i / n_a;
    int a = (int)(i % n_a);
    atomicMin(key + conn_ref[(int64_t)elem_order[ne] * n_a + a], (unsigned long long)i);
}

__global__ void k_renumber_ordered(const int* conn_ref, const int* elem_order, const int* perm, int n_a, int64_t n_el, int* conn) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_el * n_a) return;
    int64_t ne = i / n_a;
    int a = (int)(i % n_a);
    conn[i] = perm[conn_ref[(int64_t)elem_order[ne] * n_a + a]];
}

__global__ void k_invert(const int* iperm, int64_t N, int* perm) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < N) perm[iperm[i]] = (int)i;
}

__global__ void k_pair_keys(const int* conn, int n_a, int64_t n_el, unsigned long long* keys) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = n_el * n_a * n_a;
    if (i >= total) return;
    int64_t e = i / (n_a * n_a);
    int p = (int)(i % (n_a * n_a));
    unsigned long long a = (unsigned)conn[e * n_a + p / n_a], b = (unsigned)conn[e * n_a + p % n_a];
    keys[i] = (a << 32) | b;
}

// entry of node pair p of the element item_elem[item] (binary search in the element's first node's row)
__global__ void k_emap_items(const int* conn, const int* nodeptr, const int* nodecol, int n_a, const int* item_elem, int64_t n_items, int* out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = n_items * n_a * n_a;
    if (i >= total) return;
    int64_t e = item_elem[i / (n_a * n_a)];
    int p = (int)(i % (n_a * n_a));
    int a = conn[e * n_a + p / n_a], b = conn[e * n_a + p % n_a];
    int lo = nodeptr[a], hi = nodeptr[a + 1] - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (nodecol[mid] < b) lo = mid + 1; else hi = mid;
    }
    out[i] = lo;
}

// ---- deterministic scatter maps (see mfb_internal.h) ---------------------------------------------------------------------------
__global__ void k_iota_u32(unsigned* p, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (unsigned)i;
}
__global__ void k_heads(const unsigned long long* keys, int64_t n, int* head) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
// seg_start[entry] = first sorted position of the entry (+ the sentinel seg_start[U] = n)
__global__ void k_seg_start(const int* head, const int* entry_of, int64_t n, int* seg_start, int64_t U) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && head[i]) seg_start[entry_of[i]] = (int)i;
    if (i == 0) seg_start[U] = (int)n;
}
__global__ void k_extra(const int* seg_start, int64_t U, int deterministic, int* extra) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < U) extra[e] = deterministic ? (seg_start[e + 1] - seg_start[e] + 1) / 2 - 1 : 0;
}
// target of contribution vals[i] (its rank among the contributions to the entry = i - seg_start): entry or side slot
__global__ void k_targets(const unsigned* vals, const int* entry_of, const int* seg_start, const int* side_base, int64_t n, int64_t U,
                          int deterministic, int* target) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int e = entry_of[i];
    const int r = (int)i - seg_start[e];
    target[vals[i]] = (!deterministic || r < 2) ? e : (int)(U + side_base[e] + r / 2 - 1);
}
__global__ void k_slot_owner(const int* extra, const int* side_base, int64_t U, int* slot_entry, int* is_ext) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= U) return;
    for (int s = 0; s < extra[e]; ++s) slot_entry[side_base[e] + s] = (int)e;
    is_ext[e] = extra[e] > 0 ? 1 : 0;
}
__global__ void k_ext_list(const int* is_ext, const int* ext_rank, const int* extra, const int* side_base, int64_t U, int* ext_entry,
                           int* ext_first, int* ext_count) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= U || !is_ext[e]) return;
    const int j = ext_rank[e];
    ext_entry[j] = (int)e; ext_first[j] = side_base[e]; ext_count[j] = extra[e];
}
__global__ void k_keys_lo(const unsigned long long* keys, const int* head, const int* entry_of, int64_t n, int* col) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && head[i]) col[entry_of[i]] = (int)(keys[i] & 0xffffffffull);
}
__global__ void k_row_counts(const unsigned long long* keys, const int* head, int64_t n, int* cnt) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && head[i]) atomicAdd(cnt + (int)(keys[i] >> 32), 1);
}
__global__ void k_node_keys(const int* conn, int64_t n, unsigned long long* keys) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) keys[i] = (unsigned long long)(unsigned)conn[i];
}
// residual plan: ranks among the REFERENCED nodes -> node ids (a mesh may carry control points no element references)
__global__ void k_rank_to_node(int* v, int64_t n, const int* node_of_rank, int64_t n_ranks, int64_t N) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = v[i];
    v[i] = t < n_ranks ? node_of_rank[t] : (int)(N + (t - n_ranks));
}
// K[entry] += its side slots in slot order, one thread per (entry with slots, block component)
__global__ void k_combine(const int* ext_entry, const int* ext_first, const int* ext_count, int64_t n_ext, int64_t U, int BB, double* K) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_ext * BB) return;
    const int64_t j = t / BB;
    const int c = (int)(t - j * BB);
    double acc = K[(size_t)ext_entry[j] * BB + c];
    const size_t first = (size_t)(U + ext_first[j]) * BB + c;
    for (int s = 0; s < ext_count[j]; ++s) acc += K[first + (size_t)s * BB];
    K[(size_t)ext_entry[j] * BB + c] = acc;
}

// layout conversions ------------------------------------------------------------------------
__global__ void k_to_internal(const double* ref, double* in, const int* iperm, int64_t N, int nv, int levels) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = (int64_t)levels * N * nv;
    if (i >= total) return;
    int v = (int)(i % nv);
    int64_t g = (i / nv) % N;
    int64_t l = i / (nv * N);
    in[i] = ref[iperm[g] + (int64_t)v * N + l * N * nv];
}

__global__ void k_to_reference(const double* in, double* ref, const int* iperm, int64_t N, int nv, int levels) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = (int64_t)levels * N * nv;
    if (i >= total) return;
    int v = (int)(i % nv);
    int64_t g = (i / nv) % N;
    int64_t l = i / (nv * N);
    ref[iperm[g] + (int64_t)v * N + l * N * nv] = in[i];
}

__global__ void k_field_to_internal(const double* ref, double* in, const int* iperm, int64_t N) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < N) in[i] = ref[iperm[i]];
}

// reference CSR export ------------------------------------------------------------------------
// rank of each entry inside its row when the row is ordered by REFERENCE column id
__global__ void k_ref_pos(const int* nodeptr, const int* nodecol, const int* iperm, int64_t N, int* ref_pos) {
    int64_t a = blockIdx.x;  // one block per row
    if (a >= N) return;
    int s = nodeptr[a], t = nodeptr[a + 1];
    for (int k = s + threadIdx.x; k < t; k += blockDim.x) {
        int me = iperm[nodecol[k]], r = 0;
        for (int j = s; j < t; ++j) r += (iperm[nodecol[j]] < me);
        ref_pos[k] = r;
    }
}

// rowlen_ref[g_ref + i*N] = deg(g) * (#blocks with dual_pos == i)
__global__ void k_ref_rowlen(const int* nodeptr, const int* perm, int64_t N, int nv, const int* nk_of_row, int* rowlen) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= N * nv) return;
    int i = (int)(r / N);
    int64_t g = r % N;
    int a = perm[g];
    rowlen[r] = (nodeptr[a + 1] - nodeptr[a]) * nk_of_row[i];
}

// one thread per (entry, i, k)
__global__ void k_export(const int* nodeptr, const int* nodecol, const int* iperm, const int* ref_pos,
                         const int* rowptr_ref /*0-based*/, const int* kslot /*[nv*nv] index of k among populated, or -1*/,
                         int64_t N, int64_t U, int nv, const double* Kint, double* Kref, int* K_I, int* K_J,
                         const int* row_of_entry) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = U * nv * nv;
    if (t >= total) return;
    int64_t ent = t / (nv * nv);
    int ik = (int)(t % (nv * nv));
    int i = ik / nv, k = ik % nv;
    int ks = kslot[ik];
    if (ks < 0) return;
    int a = row_of_entry[ent];
    int deg = nodeptr[a + 1] - nodeptr[a];
    int64_t row = iperm[a] + (int64_t)i * N;
    int64_t pos = (int64_t)rowptr_ref[row] + (int64_t)ks * deg + ref_pos[ent];
    if (Kref) Kref[pos] = Kint[t];
    if (K_I) K_I[pos] = (int)row + 1;
    if (K_J) K_J[pos] = iperm[nodecol[ent]] + k * (int)N + 1;
}

__global__ void k_row_of_entry(const int* nodeptr, int64_t N, int* row_of_entry) {
    int64_t a = blockIdx.x;
    if (a >= N) return;
    for (int k = nodeptr[a] + threadIdx.x; k < nodeptr[a + 1]; k += blockDim.x) row_of_entry[k] = (int)a;
}

// sparse_IDs_by_el[a, b, e] (column-major, e = REFERENCE element): 1-based position, in the exported reference-layout CSR,
// of the entry that pair (a, b) of element e feeds in variable block (i, k)
__global__ void k_sparse_ids(const int* emap, const int* slot_entry, int64_t U, const int* elem_rank, const int* nodeptr, const int* iperm,
                             const int* ref_pos, const int* rowptr_ref, const int* row_of_entry, int n_a, int64_t n_el, int64_t N, int i,
                             int ks, int* out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t per = (int64_t)n_a * n_a;
    if (t >= per * n_el) return;
    const int64_t e = t / per;
    const int p = (int)(t - e * per);
    const int a = p % n_a, b = p / n_a;                     // output index a + n_a * (b + n_a * e)
    int ent = emap[(int64_t)elem_rank[e] * per + a * n_a + b];
    if (ent >= U) ent = slot_entry[ent - U];              // a side slot of the deterministic scatter: the entry it is folded into
    const int row_node = row_of_entry[ent];
    const int deg = nodeptr[row_node + 1] - nodeptr[row_node];
    const int64_t row = iperm[row_node] + (int64_t)i * N;
    out[t] = rowptr_ref[row] + ks * deg + ref_pos[ent] + 1;
}

__global__ void k_iota1(int* p, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int)i + 1;
}
__global__ void k_add1(int* p, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] += 1;
}

}  // namespace

#define LAUNCH(kernel, grid, block, ...)                      \
    do {                                                      \
        kernel<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__); \
        ctx->launches++;                                      \
    } while (0)

// Internal numbering: elements are ordered along a Morton curve of their centroids and nodes by first touch in that
// element order, so a row's neighbours and consecutive rows' neighbours sit close together in x (L1-resident gathers in
// the SpMV) and neighbouring elements scatter into nearby matrix entries (L2-resident atomics in the assembly).
// x1..x3: device pointers in REFERENCE node order. MFB_NO_PERMUTE=1 keeps the reference order (experiments).
int mfb_build_permutation(mfb_ctx* ctx, const double* x1, const double* x2, const double* x3) {
    const int64_t N = ctx->N, n_el = ctx->n_el, nconn = n_el * ctx->n_a;
    auto pol = thrust::cuda::par.on(ctx->stream);
    MFB_CUDA(ctx->perm.alloc(N));
    MFB_CUDA(ctx->iperm.alloc(N));
    MFB_CUDA(ctx->conn.alloc(nconn));
    MFB_CUDA(ctx->elem_order.alloc(n_el));
    MFB_CUDA(ctx->elem_rank.alloc(n_el));
    thrust::device_ptr<int> ip(ctx->iperm.p), eo(ctx->elem_order.p);
    thrust::sequence(pol, ip, ip + N);
    thrust::sequence(pol, eo, eo + n_el);
    const char* nop = getenv("MFB_NO_PERMUTE");
    if (!(nop && nop[0] == '1')) {
        double lo[3], hi[3];
        const double* xs[3] = {x1, x2, x3};
        for (int d = 0; d < 3; ++d) {
            thrust::device_ptr<const double> p(xs[d]);
            auto mm = thrust::minmax_element(pol, p, p + N);
            MFB_CUDA(cudaMemcpyAsync(&lo[d], thrust::raw_pointer_cast(mm.first), sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            MFB_CUDA(cudaMemcpyAsync(&hi[d], thrust::raw_pointer_cast(mm.second), sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        }
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        double ext = 0.0;
        for (int d = 0; d < 3; ++d) ext = ext > hi[d] - lo[d] ? ext : hi[d] - lo[d];
        // same cell size on all axes, about one cell per element edge: 2^bits cells across the largest extent
        int bits = 1;
        while ((1ll << (3 * bits)) < n_el * 8 && bits < 21) ++bits;
        const double inv = ext > 0 ? (double)(1ll << bits) / ext : 0.0;
        DevBuf<unsigned long long> ekey;
        MFB_CUDA(ekey.alloc(n_el));
        LAUNCH(k_elem_morton, nblk(n_el), TPB, ctx->conn_ref.p, ctx->n_a, n_el, x1, x2, x3, lo[0], lo[1], lo[2], inv, inv, inv,
               ekey.p, ctx->elem_order.p);
        thrust::device_ptr<unsigned long long> ek(ekey.p);
        thrust::stable_sort_by_key(pol, ek, ek + n_el, eo);
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        ekey.release();
        DevBuf<unsigned long long> key;
        MFB_CUDA(key.alloc(N));
        MFB_CUDA(cudaMemsetAsync(key.p, 0xff, N * sizeof(unsigned long long), ctx->stream));
        LAUNCH(k_first_touch_ordered, nblk(nconn), TPB, ctx->conn_ref.p, ctx->elem_order.p, ctx->n_a, n_el, key.p);
        thrust::device_ptr<unsigned long long> kp(key.p);
        thrust::sort_by_key(pol, kp, kp + N, ip);
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        key.release();
    }
    LAUNCH(k_invert, nblk(N), TPB, ctx->iperm.p, N, ctx->perm.p);
    LAUNCH(k_invert, nblk(n_el), TPB, ctx->elem_order.p, n_el, ctx->elem_rank.p);
    LAUNCH(k_renumber_ordered, nblk(nconn), TPB, ctx->conn_ref.p, ctx->elem_order.p, ctx->perm.p, ctx->n_a, n_el, ctx->conn.p);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

namespace {
// Sorts (key, contribution index) pairs, numbers the distinct keys (`entry_of` per sorted position, count U_out) and writes the
// scatter target of every contribution (`target[contribution]`): the key's index, or -- deterministic mode, contribution of
// rank >= 2 -- one of the key's side slots U_out + s. Leaves seg_start[U_out + 1], extra / side_base [U_out] for the caller.
struct SlotPlan {
    DevBuf<int> head, entry_of, seg_start, extra, side_base;
    int64_t U = 0, S = 0;
};
int plan_slots(mfb_ctx* ctx, DevBuf<unsigned long long>& keys, int64_t total, int* target, SlotPlan& P) {
    auto pol = thrust::cuda::par.on(ctx->stream);
    DevBuf<unsigned> vals;
    MFB_CUDA(vals.alloc(total));
    LAUNCH(k_iota_u32, nblk(total), TPB, vals.p, total);
    thrust::device_ptr<unsigned long long> kp(keys.p);
    thrust::device_ptr<unsigned> vp(vals.p);
    thrust::stable_sort_by_key(pol, kp, kp + total, vp);             // radix sort: contributions of one key stay in element order
    MFB_CUDA(P.head.alloc(total)); MFB_CUDA(P.entry_of.alloc(total));
    LAUNCH(k_heads, nblk(total), TPB, keys.p, total, P.head.p);
    thrust::device_ptr<int> hp(P.head.p), ep(P.entry_of.p);
    thrust::inclusive_scan(pol, hp, hp + total, ep);
    thrust::transform(pol, ep, ep + total, ep, [] __device__(int v) { return v - 1; });
    int last = 0;
    MFB_CUDA(cudaMemcpyAsync(&last, P.entry_of.p + total - 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    P.U = (int64_t)last + 1;
    MFB_CUDA(P.seg_start.alloc(P.U + 1)); MFB_CUDA(P.extra.alloc(P.U)); MFB_CUDA(P.side_base.alloc(P.U));
    LAUNCH(k_seg_start, nblk(total), TPB, P.head.p, P.entry_of.p, total, P.seg_start.p, P.U);
    LAUNCH(k_extra, nblk(P.U), TPB, P.seg_start.p, P.U, ctx->deterministic ? 1 : 0, P.extra.p);
    thrust::device_ptr<int> xp(P.extra.p), bp(P.side_base.p);
    thrust::exclusive_scan(pol, xp, xp + P.U, bp);
    int lb = 0, lx = 0;
    MFB_CUDA(cudaMemcpyAsync(&lb, P.side_base.p + P.U - 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MFB_CUDA(cudaMemcpyAsync(&lx, P.extra.p + P.U - 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    P.S = (int64_t)lb + lx;
    LAUNCH(k_targets, nblk(total), TPB, vals.p, P.entry_of.p, P.seg_start.p, P.side_base.p, total, P.U, ctx->deterministic ? 1 : 0, target);
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MFB_OK;
}
int ext_lists(mfb_ctx* ctx, SlotPlan& P, DevBuf<int>* slot_entry, DevBuf<int>& ext_entry, DevBuf<int>& ext_first, DevBuf<int>& ext_count,
              int64_t* n_ext) {
    auto pol = thrust::cuda::par.on(ctx->stream);
    DevBuf<int> is_ext, ext_rank, owner_tmp;
    MFB_CUDA(is_ext.alloc(P.U)); MFB_CUDA(ext_rank.alloc(P.U));
    DevBuf<int>& owner = slot_entry ? *slot_entry : owner_tmp;
    MFB_CUDA(owner.alloc(P.S > 0 ? P.S : 1));
    LAUNCH(k_slot_owner, nblk(P.U), TPB, P.extra.p, P.side_base.p, P.U, owner.p, is_ext.p);
    thrust::device_ptr<int> ip(is_ext.p), rp(ext_rank.p);
    thrust::exclusive_scan(pol, ip, ip + P.U, rp);
    int lr = 0, lf = 0;
    MFB_CUDA(cudaMemcpyAsync(&lr, ext_rank.p + P.U - 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MFB_CUDA(cudaMemcpyAsync(&lf, is_ext.p + P.U - 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_ext = (int64_t)lr + lf;
    const int64_t na = *n_ext > 0 ? *n_ext : 1;
    MFB_CUDA(ext_entry.alloc(na)); MFB_CUDA(ext_first.alloc(na)); MFB_CUDA(ext_count.alloc(na));
    LAUNCH(k_ext_list, nblk(P.U), TPB, is_ext.p, ext_rank.p, P.extra.p, P.side_base.p, P.U, ext_entry.p, ext_first.p, ext_count.p);
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MFB_OK;
}
}  // namespace

int mfb_build_pattern(mfb_ctx* ctx) {
    const int64_t N = ctx->N;
    const int n_a = ctx->n_a;
    const int64_t total = ctx->n_el * n_a * n_a;
    MFB_REQUIRE(total < 4294967295LL, MFB_ERR_ARG, "more than 2^32 element node pairs");
    {
        const char* e = getenv("MFB_DETERMINISTIC");
        ctx->deterministic = !(e && e[0] == '0');
    }
    auto pol = thrust::cuda::par.on(ctx->stream);
    // ---- node graph + scatter map of the matrix ----
    {
        DevBuf<unsigned long long> keys;
        MFB_CUDA(keys.alloc(total));
        LAUNCH(k_pair_keys, nblk(total), TPB, ctx->conn.p, n_a, ctx->n_el, keys.p);
        MFB_CUDA(ctx->emap.alloc(total));
        SlotPlan P;
        MFB_TRY(plan_slots(ctx, keys, total, ctx->emap.p, P));
        ctx->U = P.U;
        ctx->S = P.S;
        MFB_REQUIRE(ctx->U + ctx->S < (int64_t)2147483647, MFB_ERR_ARG, "node graph exceeds int32 indexing");
        MFB_CUDA(ctx->nodecol.alloc(ctx->U + MFB_STREAM_PAD));
        MFB_CUDA(cudaMemsetAsync(ctx->nodecol.p + ctx->U, 0, MFB_STREAM_PAD * sizeof(int), ctx->stream));
        MFB_CUDA(ctx->nodeptr.alloc(N + 1));
        LAUNCH(k_keys_lo, nblk(total), TPB, keys.p, P.head.p, P.entry_of.p, total, ctx->nodecol.p);
        // row pointer: entries per row (keys are sorted by row, then column), exclusive scan
        MFB_CUDA(cudaMemsetAsync(ctx->nodeptr.p, 0, (N + 1) * sizeof(int), ctx->stream));
        LAUNCH(k_row_counts, nblk(total), TPB, keys.p, P.head.p, total, ctx->nodeptr.p);
        thrust::device_ptr<int> np(ctx->nodeptr.p);
        thrust::exclusive_scan(pol, np, np + (N + 1), np);
        MFB_TRY(ext_lists(ctx, P, &ctx->slot_entry, ctx->ext_entry, ctx->ext_first, ctx->ext_count, &ctx->n_ext));
    }
    // ---- scatter map of the residual ----
    {
        const int64_t nconn = ctx->n_el * n_a;
        DevBuf<unsigned long long> keys;
        MFB_CUDA(keys.alloc(nconn));
        LAUNCH(k_node_keys, nblk(nconn), TPB, ctx->conn.p, nconn, keys.p);
        MFB_CUDA(ctx->rmap.alloc(nconn));
        SlotPlan P;
        MFB_TRY(plan_slots(ctx, keys, nconn, ctx->rmap.p, P));
        ctx->SR = P.S;
        MFB_TRY(ext_lists(ctx, P, nullptr, ctx->rext_node, ctx->rext_first, ctx->rext_count, &ctx->n_rext));
        // the plan numbers the REFERENCED nodes (a mesh may carry control points no element touches): back to node ids
        DevBuf<int> node_of_rank;
        MFB_CUDA(node_of_rank.alloc(P.U));
        LAUNCH(k_keys_lo, nblk(nconn), TPB, keys.p, P.head.p, P.entry_of.p, nconn, node_of_rank.p);
        LAUNCH(k_rank_to_node, nblk(nconn), TPB, ctx->rmap.p, nconn, node_of_rank.p, P.U, N);
        if (ctx->n_rext > 0) LAUNCH(k_rank_to_node, nblk(ctx->n_rext), TPB, ctx->rext_node.p, ctx->n_rext, node_of_rank.p, P.U, N);
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    MFB_CUDA(cudaGetLastError());
    ctx->ref_pos.release();
    ctx->n_lin_entries = -1;
    if (const char* dump = getenv("MFB_DUMP_PATTERN")) {   // experiments (profiles/exp): int64 N, int64 U, nodeptr, nodecol
        std::vector<int> hp(N + 1), hc(ctx->U);
        MFB_CUDA(cudaMemcpy(hp.data(), ctx->nodeptr.p, (N + 1) * sizeof(int), cudaMemcpyDeviceToHost));
        MFB_CUDA(cudaMemcpy(hc.data(), ctx->nodecol.p, ctx->U * sizeof(int), cudaMemcpyDeviceToHost));
        if (FILE* f = fopen(dump, "wb")) {
            int64_t hdr[2] = {N, ctx->U};
            fwrite(hdr, sizeof(int64_t), 2, f);
            fwrite(hp.data(), sizeof(int), hp.size(), f);
            fwrite(hc.data(), sizeof(int), hc.size(), f);
            fclose(f);
        }
    }
    return MFB_OK;
}

// entries (never side slots) of all node pairs of the given elements: the scatter map of the boundary kernels
int mfb_entry_map(mfb_ctx* ctx, const int* item_elem, int64_t n_items, int* out) {
    LAUNCH(k_emap_items, nblk(n_items * ctx->n_a * ctx->n_a), TPB, ctx->conn.p, ctx->nodeptr.p, ctx->nodecol.p, ctx->n_a, item_elem, n_items, out);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

int mfb_combine_matrix(mfb_ctx* ctx, double* K) {
    if (ctx->n_ext <= 0) return MFB_OK;
    const int BB = ctx->n_var * ctx->n_var;
    LAUNCH(k_combine, nblk(ctx->n_ext * BB), TPB, ctx->ext_entry.p, ctx->ext_first.p, ctx->ext_count.p, ctx->n_ext, ctx->U, BB, K);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}
int mfb_combine_residue(mfb_ctx* ctx, double* res) {
    if (ctx->n_rext <= 0) return MFB_OK;
    LAUNCH(k_combine, nblk(ctx->n_rext * ctx->n_var), TPB, ctx->rext_node.p, ctx->rext_first.p, ctx->rext_count.p, ctx->n_rext, ctx->N,
           ctx->n_var, res);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

int mfb_to_internal(mfb_ctx* ctx, const double* ref_vec, double* int_vec, int levels) {
    int64_t total = (int64_t)levels * ctx->N * ctx->n_var;
    LAUNCH(k_to_internal, nblk(total), TPB, ref_vec, int_vec, ctx->iperm.p, ctx->N, ctx->n_var, levels);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

int mfb_to_reference(mfb_ctx* ctx, const double* int_vec, double* ref_vec, int levels) {
    int64_t total = (int64_t)levels * ctx->N * ctx->n_var;
    LAUNCH(k_to_reference, nblk(total), TPB, int_vec, ref_vec, ctx->iperm.p, ctx->N, ctx->n_var, levels);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

int mfb_field_to_internal(mfb_ctx* ctx, const double* ref_field, double* int_field) {
    LAUNCH(k_field_to_internal, nblk(ctx->N), TPB, ref_field, int_field, ctx->iperm.p, ctx->N);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

namespace {
struct RefLayout {
    DevBuf<int> rowptr, row_of_entry, kslot, nk;
};
int build_ref_layout(mfb_ctx* ctx, RefLayout& R) {
    const int64_t N = ctx->N, U = ctx->U;
    const int nv = ctx->n_var;
    auto pol = thrust::cuda::par.on(ctx->stream);
    if (!ctx->ref_pos.p) {
        MFB_CUDA(ctx->ref_pos.alloc(U));
        LAUNCH(k_ref_pos, (unsigned)N, 64, ctx->nodeptr.p, ctx->nodecol.p, ctx->iperm.p, N, ctx->ref_pos.p);
    }
    std::vector<int> kslot(nv * nv, -1), nk(nv, 0);
    for (int i = 0; i < nv; ++i)
        for (int k = 0; k < nv; ++k)
            if (ctx->block_of[i * nv + k] >= 0) kslot[i * nv + k] = nk[i]++;
    MFB_CUDA(R.kslot.alloc(nv * nv));
    MFB_CUDA(R.nk.alloc(nv));
    MFB_CUDA(cudaMemcpyAsync(R.kslot.p, kslot.data(), nv * nv * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    MFB_CUDA(cudaMemcpyAsync(R.nk.p, nk.data(), nv * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    MFB_CUDA(R.rowptr.alloc(N * nv + 1));
    MFB_CUDA(cudaMemsetAsync(R.rowptr.p, 0, (N * nv + 1) * sizeof(int), ctx->stream));
    LAUNCH(k_ref_rowlen, nblk(N * nv), TPB, ctx->nodeptr.p, ctx->perm.p, N, nv, R.nk.p, R.rowptr.p);
    thrust::device_ptr<int> rp(R.rowptr.p);
    thrust::exclusive_scan(pol, rp, rp + (N * nv + 1), rp);
    MFB_CUDA(R.row_of_entry.alloc(U));
    LAUNCH(k_row_of_entry, (unsigned)N, 64, ctx->nodeptr.p, N, R.row_of_entry.p);
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MFB_OK;
}
}  // namespace

int mfb_export_matrix(mfb_ctx* ctx, const double* Kint, double* Kref_dev) {
    RefLayout R;
    MFB_TRY(build_ref_layout(ctx, R));
    const int nv = ctx->n_var;
    int64_t total = ctx->U * nv * nv;
    LAUNCH(k_export, nblk(total), TPB, ctx->nodeptr.p, ctx->nodecol.p, ctx->iperm.p, ctx->ref_pos.p, R.rowptr.p,
           R.kslot.p, ctx->N, ctx->U, nv, Kint, Kref_dev, (int*)nullptr, (int*)nullptr, R.row_of_entry.p);
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MFB_OK;
}

int mfb_export_pattern(mfb_ctx* ctx, int* K_I, int* K_J, int* K_J_ptr, int* K_val_ids) {
    RefLayout R;
    MFB_TRY(build_ref_layout(ctx, R));
    const int nv = ctx->n_var;
    int64_t total = ctx->U * nv * nv;
    int64_t nnz = ctx->U * ctx->n_blocks;
    if (K_I || K_J)
        LAUNCH(k_export, nblk(total), TPB, ctx->nodeptr.p, ctx->nodecol.p, ctx->iperm.p, ctx->ref_pos.p, R.rowptr.p,
               R.kslot.p, ctx->N, ctx->U, nv, (const double*)nullptr, (double*)nullptr, K_I, K_J, R.row_of_entry.p);
    if (K_J_ptr) {
        MFB_CUDA(cudaMemcpyAsync(K_J_ptr, R.rowptr.p, (ctx->N * nv + 1) * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
        LAUNCH(k_add1, nblk(ctx->N * nv + 1), TPB, K_J_ptr, ctx->N * nv + 1);
    }
    if (K_val_ids) LAUNCH(k_iota1, nblk(nnz), TPB, K_val_ids, nnz);
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MFB_OK;
}

// sparse_IDs_by_el of one variable block in the reference's table layout (03_GlobalAssembly.jl:111-118): device pointer out
int mfb_export_sparse_ids(mfb_ctx* ctx, int block, int* out_dev) {
    RefLayout R;
    MFB_TRY(build_ref_layout(ctx, R));
    const int nv = ctx->n_var;
    const int i = ctx->sparse_mapping[2 * block], k = ctx->sparse_mapping[2 * block + 1];
    int ks = 0;
    for (int kk = 0; kk < k; ++kk) ks += ctx->block_of[i * nv + kk] >= 0;
    const int64_t total = ctx->n_el * ctx->n_a * ctx->n_a;
    LAUNCH(k_sparse_ids, nblk(total), TPB, ctx->emap.p, ctx->slot_entry.p, ctx->U, ctx->elem_rank.p, ctx->nodeptr.p, ctx->iperm.p, ctx->ref_pos.p, R.rowptr.p,
           R.row_of_entry.p, ctx->n_a, ctx->n_el, ctx->N, i, ks, out_dev);
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MFB_OK;
}
