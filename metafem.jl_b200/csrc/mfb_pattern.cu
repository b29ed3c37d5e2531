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

__global__ void k_first_touch(const int* conn_ref, int64_t n, unsigned long long* key) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) atomicMin(key + conn_ref[i], (unsigned long long)i);
}

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

__global__ void k_split(const unsigned long long* keys, int64_t U, int* col) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < U) col[i] = (int)(keys[i] & 0xffffffffull);
}

struct RowKey {
    __host__ __device__ unsigned long long operator()(int64_t r) const { return ((unsigned long long)r) << 32; }
};

__global__ void k_emap(const int* conn, const int* nodeptr, const int* nodecol, int n_a, int64_t n_el, int* emap) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = n_el * n_a * n_a;
    if (i >= total) return;
    int64_t e = i / (n_a * n_a);
    int p = (int)(i % (n_a * n_a));
    int a = conn[e * n_a + p / n_a], b = conn[e * n_a + p % n_a];
    int lo = nodeptr[a], hi = nodeptr[a + 1] - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (nodecol[mid] < b) lo = mid + 1; else hi = mid;
    }
    emap[i] = lo;
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
__global__ void k_sparse_ids(const int* emap, const int* elem_rank, const int* nodeptr, const int* iperm, const int* ref_pos,
                             const int* rowptr_ref, const int* row_of_entry, int n_a, int64_t n_el, int64_t N, int i, int ks, int* out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t per = (int64_t)n_a * n_a;
    if (t >= per * n_el) return;
    const int64_t e = t / per;
    const int p = (int)(t - e * per);
    const int a = p % n_a, b = p / n_a;                     // output index a + n_a * (b + n_a * e)
    const int ent = emap[(int64_t)elem_rank[e] * per + a * n_a + b];
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

int mfb_build_pattern(mfb_ctx* ctx) {
    const int64_t N = ctx->N;
    const int n_a = ctx->n_a;
    const int64_t total = ctx->n_el * n_a * n_a;
    auto pol = thrust::cuda::par.on(ctx->stream);
    DevBuf<unsigned long long> keys;
    MFB_CUDA(keys.alloc(total));
    LAUNCH(k_pair_keys, nblk(total), TPB, ctx->conn.p, n_a, ctx->n_el, keys.p);
    thrust::device_ptr<unsigned long long> kp(keys.p);
    thrust::sort(pol, kp, kp + total);
    auto end = thrust::unique(pol, kp, kp + total);
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->U = end - kp;
    MFB_REQUIRE(ctx->U < (int64_t)2147483647, MFB_ERR_ARG, "node graph exceeds int32 indexing");
    MFB_CUDA(ctx->nodecol.alloc(ctx->U + MFB_STREAM_PAD));
    MFB_CUDA(cudaMemsetAsync(ctx->nodecol.p + ctx->U, 0, MFB_STREAM_PAD * sizeof(int), ctx->stream));
    MFB_CUDA(ctx->nodeptr.alloc(N + 1));
    LAUNCH(k_split, nblk(ctx->U), TPB, keys.p, ctx->U, ctx->nodecol.p);
    thrust::counting_iterator<int64_t> c0(0);
    auto rows = thrust::make_transform_iterator(c0, RowKey());
    thrust::lower_bound(pol, kp, kp + ctx->U, rows, rows + (N + 1), thrust::device_ptr<int>(ctx->nodeptr.p));
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    keys.release();
    MFB_CUDA(ctx->emap.alloc(total));
    LAUNCH(k_emap, nblk(total), TPB, ctx->conn.p, ctx->nodeptr.p, ctx->nodecol.p, n_a, ctx->n_el, ctx->emap.p);
    MFB_CUDA(cudaGetLastError());
    ctx->ref_pos.release();
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
    LAUNCH(k_sparse_ids, nblk(total), TPB, ctx->emap.p, ctx->elem_rank.p, ctx->nodeptr.p, ctx->iperm.p, ctx->ref_pos.p, R.rowptr.p,
           R.row_of_entry.p, ctx->n_a, ctx->n_el, ctx->N, i, ks, out_dev);
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MFB_OK;
}
