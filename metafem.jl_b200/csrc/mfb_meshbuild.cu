// mfb_meshbuild.cu -- second-order mesh tables from a first-order mesh, on the device.
//
// Reference: construct_TotalMesh_3D builds vertex/segment/face/block tables through its GPU hash (FEM_Dict,
// src/mesh/ref_geometry/002_Initialization.jl:113-217, src/misc/06_GPU_Dict.jl), get_BoundaryMesh picks the faces that
// belong to one block, and mesh_Classical / allocate_Basic_WP_Mesh_3D place one control point per vertex and one per
// segment and fill elements.controlpoint_IDs (src/mesh/unstructured_mesh/3_InitializeMesh.jl:70-178) -- what the hot path
// consumes. Here the same tables come out of radix sorts instead of a hash:
//   * segments: key = max(v_a, v_b) * (n_vert + 1) + min(v_a, v_b) over all (element, local segment); sort + unique; the
//     mid-edge control point of a segment is numbered n_vert + rank(key) + 1. Vertex control points keep the input
//     order, exactly as in the reference; mid-edge IDs are in sorted-key order -- a deterministic, locality-preserving
//     stand-in for the reference's hash-slot order (which is racy on the GPU). The library renumbers
//     internally along a Morton curve anyway (mfb_build_permutation).
//   * boundary facets: key = the three smallest vertex ids of the face; a face whose key occurs once is a boundary
//     face. They are listed in (local face, element) order with their centroid, so that the script's geometric
//     selection of boundary groups (e.g. static_Neo_Hookean.jl:19-34) runs on the returned centroids.
// The element-type tables (local segment/face vertices, control-point slots: 101_Structures.jl:129-196,224-247) are
// inputs, so hex8 -> hex20 and tet4 -> tet10 share the code.
#include <thrust/binary_search.h>
#include <thrust/copy.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#include "mfb_internal.h"

namespace {
constexpr int TPB = 256;
inline unsigned nblk(int64_t n) { return (unsigned)((n + TPB - 1) / TPB); }
typedef unsigned long long u64;

struct Topo {
    int vpb, n_seg, n_faces, vpf;
    int seg[32][2];       // local vertex ids (0-based) of each segment
    int vcp[16];          // control-point slot (0-based) of each vertex
    int scp[32];          // control-point slot (0-based) of each segment
    int face[8][4];       // local vertex ids (0-based) of each face
};

__global__ void k_seg_keys(const int* conn, int64_t n_el, Topo T, u64 nv1, u64* keys) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_el * T.n_seg) return;
    const int64_t e = t / T.n_seg;
    const int j = (int)(t - e * T.n_seg);
    const u64 a = conn[e * T.vpb + T.seg[j][0]], b = conn[e * T.vpb + T.seg[j][1]];     // 1-based vertex ids
    keys[t] = (a > b ? a : b) * nv1 + (a > b ? b : a);
}
__global__ void k_face_keys(const int* conn, int64_t n_el, Topo T, u64 nv1, u64* keys) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_el * T.n_faces) return;
    const int f = (int)(t / n_el);                       // (local face, element) order
    const int64_t e = t - (int64_t)f * n_el;
    u64 v[4];
    for (int k = 0; k < T.vpf; ++k) v[k] = conn[e * T.vpb + T.face[f][k]];
    for (int i = 1; i < T.vpf; ++i)                      // insertion sort, ascending
        for (int k = i; k > 0 && v[k] < v[k - 1]; --k) { const u64 s = v[k]; v[k] = v[k - 1]; v[k - 1] = s; }
    keys[t] = (v[0] * nv1 + v[1]) * nv1 + v[2];          // the three smallest ids identify a triangle or a quad
}
// controlpoint_IDs [n_a, n_el] column-major, 1-based
__global__ void k_fill_cp(const int* conn, int64_t n_el, Topo T, int n_a, const u64* seg_keys, const u64* uniq, int64_t n_uniq,
                          int64_t n_vert, int* cp) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_el * (T.vpb + T.n_seg)) return;
    const int64_t e = t / (T.vpb + T.n_seg);
    const int j = (int)(t - e * (T.vpb + T.n_seg));
    if (j < T.vpb) {
        cp[e * n_a + T.vcp[j]] = conn[e * T.vpb + j];
    } else {
        const u64 key = seg_keys[e * T.n_seg + (j - T.vpb)];
        int64_t lo = 0, hi = n_uniq - 1;                 // rank of the key in the sorted unique list
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (uniq[mid] < key) lo = mid + 1; else hi = mid;
        }
        cp[e * n_a + T.scp[j - T.vpb]] = (int)(n_vert + lo + 1);
    }
}
__global__ void k_mid_coords(const u64* uniq, int64_t n_uniq, u64 nv1, int64_t n_vert, const double* x1, const double* x2,
                             const double* x3, double* o1, double* o2, double* o3) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_uniq) return;
    const int64_t hi = (int64_t)(uniq[t] / nv1) - 1, lo = (int64_t)(uniq[t] % nv1) - 1;
    o1[n_vert + t] = 0.5 * x1[hi] + 0.5 * x1[lo];
    o2[n_vert + t] = 0.5 * x2[hi] + 0.5 * x2[lo];
    o3[n_vert + t] = 0.5 * x3[hi] + 0.5 * x3[lo];
}
struct IsBoundary {
    const u64* keys;        // face keys in (face, element) order
    const u64* sorted;      // the same keys sorted
    int64_t n;
    __device__ bool operator()(int64_t t) const {
        const u64 k = keys[t];
        int64_t lo = 0, hi = n;                          // lower bound
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (sorted[mid] < k) lo = mid + 1; else hi = mid; }
        return lo + 1 >= n || sorted[lo + 1] != k;       // the key occurs exactly once
    }
};
__global__ void k_bfacets(const int64_t* ids, int64_t n_bf, int64_t n_el, const int* conn, Topo T, const double* x1, const double* x2,
                          const double* x3, int* f_el, int* f_eidx, double* cen) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_bf) return;
    const int f = (int)(ids[t] / n_el);
    const int64_t e = ids[t] - (int64_t)f * n_el;
    f_el[t] = (int)(e + 1);
    f_eidx[t] = f + 1;
    double c[3] = {0, 0, 0};
    for (int k = 0; k < T.vpf; ++k) {
        const int v = conn[e * T.vpb + T.face[f][k]] - 1;
        c[0] += x1[v]; c[1] += x2[v]; c[2] += x3[v];
    }
    cen[3 * t + 0] = c[0] / T.vpf; cen[3 * t + 1] = c[1] / T.vpf; cen[3 * t + 2] = c[2] / T.vpf;     // [3, n_bf] column-major
}
}  // namespace

#define LAUNCH(kernel, grid, block, ...)                          \
    do {                                                          \
        kernel<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__); \
        ctx->launches++;                                          \
    } while (0)

struct MeshBuild {
    int n_a = 0;
    int64_t n_el = 0, N = 0, n_bf = 0;
    DevBuf<int> cp, f_el, f_eidx;
    DevBuf<double> x1, x2, x3, cen;
};

static void free_build(mfb_ctx* ctx) {
    delete ctx->meshbuild;
    ctx->meshbuild = nullptr;
}
void mfb_meshbuild_free(mfb_ctx* ctx) { free_build(ctx); }

extern "C" int mfb_mesh_build_second_order(mfb_ctx* ctx, int64_t n_vert, const double* x1, const double* x2, const double* x3,
                                           int vpb, int64_t n_el, const int32_t* connections, int n_seg,
                                           const int32_t* segment_vertices, const int32_t* vertex_cp_ids,
                                           const int32_t* segment_cp_ids, int n_faces, int vpf, const int32_t* face_vertices,
                                           int64_t* n_controlpoints, int64_t* n_boundary_facets) {
    if (!ctx) return MFB_ERR_ARG;
    MFB_REQUIRE(n_vert > 0 && n_el > 0 && connections && x1 && x2 && x3 && segment_vertices && vertex_cp_ids && segment_cp_ids &&
                    face_vertices, MFB_ERR_ARG, "mfb_mesh_build_second_order: null or empty input");
    MFB_REQUIRE(vpb >= 4 && vpb <= 16 && n_seg >= 1 && n_seg <= 32 && n_faces >= 1 && n_faces <= 8 && (vpf == 3 || vpf == 4),
                MFB_ERR_ARG, "mfb_mesh_build_second_order: unsupported element topology");
    MFB_REQUIRE((double)(n_vert + 1) * (n_vert + 1) * (n_vert + 1) < 9.0e18, MFB_ERR_ARG, "vertex count too large for 64-bit face keys");
    MFB_CUDA(cudaSetDevice(ctx->device));
    free_build(ctx);
    MeshBuild* B = ctx->meshbuild = new MeshBuild();
    Topo T;
    memset(&T, 0, sizeof(T));
    T.vpb = vpb; T.n_seg = n_seg; T.n_faces = n_faces; T.vpf = vpf;
    const int n_a = vpb + n_seg;
    for (int j = 0; j < n_seg; ++j) {                    // host tables (element-type constants), 1-based in, 0-based kept
        T.seg[j][0] = segment_vertices[2 * j] - 1; T.seg[j][1] = segment_vertices[2 * j + 1] - 1;
        T.scp[j] = segment_cp_ids[j] - 1;
        MFB_REQUIRE(T.seg[j][0] >= 0 && T.seg[j][0] < vpb && T.seg[j][1] >= 0 && T.seg[j][1] < vpb && T.scp[j] >= 0 && T.scp[j] < n_a,
                    MFB_ERR_ARG, "segment table out of range");
    }
    for (int j = 0; j < vpb; ++j) {
        T.vcp[j] = vertex_cp_ids[j] - 1;
        MFB_REQUIRE(T.vcp[j] >= 0 && T.vcp[j] < n_a, MFB_ERR_ARG, "vertex control-point table out of range");
    }
    for (int f = 0; f < n_faces; ++f)
        for (int k = 0; k < vpf; ++k) {
            T.face[f][k] = face_vertices[f * vpf + k] - 1;
            MFB_REQUIRE(T.face[f][k] >= 0 && T.face[f][k] < vpb, MFB_ERR_ARG, "face table out of range");
        }
    auto pol = thrust::cuda::par.on(ctx->stream);
    const u64 nv1 = (u64)n_vert + 1;
    // ---- inputs on the device: connections [vpb, n_el] column-major == [n_el][vpb] row-major ----
    DevBuf<int> conn;
    DevBuf<double> v1, v2, v3;
    MFB_CUDA(conn.alloc(n_el * vpb));
    MFB_CUDA(v1.alloc(n_vert)); MFB_CUDA(v2.alloc(n_vert)); MFB_CUDA(v3.alloc(n_vert));
    MFB_TRY(mfb_stage_in(ctx, connections, n_el * vpb * sizeof(int), conn.p));
    MFB_TRY(mfb_stage_in(ctx, x1, n_vert * sizeof(double), v1.p));
    MFB_TRY(mfb_stage_in(ctx, x2, n_vert * sizeof(double), v2.p));
    MFB_TRY(mfb_stage_in(ctx, x3, n_vert * sizeof(double), v3.p));
    // ---- segments -> mid-edge control points ----
    const int64_t ns = n_el * n_seg;
    DevBuf<u64> skeys, suniq;
    MFB_CUDA(skeys.alloc(ns)); MFB_CUDA(suniq.alloc(ns));
    LAUNCH(k_seg_keys, nblk(ns), TPB, conn.p, n_el, T, nv1, skeys.p);
    MFB_CUDA(cudaMemcpyAsync(suniq.p, skeys.p, ns * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
    thrust::device_ptr<u64> up(suniq.p);
    thrust::sort(pol, up, up + ns);
    const int64_t n_edges = thrust::unique(pol, up, up + ns) - up;
    B->n_a = n_a; B->n_el = n_el; B->N = n_vert + n_edges;
    MFB_REQUIRE(B->N < 2147483647LL, MFB_ERR_ARG, "control-point count exceeds Int32");
    MFB_CUDA(B->cp.alloc(n_el * n_a));
    LAUNCH(k_fill_cp, nblk(n_el * n_a), TPB, conn.p, n_el, T, n_a, skeys.p, suniq.p, n_edges, n_vert, B->cp.p);
    MFB_CUDA(B->x1.alloc(B->N)); MFB_CUDA(B->x2.alloc(B->N)); MFB_CUDA(B->x3.alloc(B->N));
    MFB_CUDA(cudaMemcpyAsync(B->x1.p, v1.p, n_vert * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    MFB_CUDA(cudaMemcpyAsync(B->x2.p, v2.p, n_vert * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    MFB_CUDA(cudaMemcpyAsync(B->x3.p, v3.p, n_vert * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    LAUNCH(k_mid_coords, nblk(n_edges), TPB, suniq.p, n_edges, nv1, n_vert, v1.p, v2.p, v3.p, B->x1.p, B->x2.p, B->x3.p);
    skeys.release(); suniq.release();
    // ---- boundary facets: faces whose key occurs once (get_BoundaryMesh) ----
    const int64_t nf = n_el * n_faces;
    DevBuf<u64> fkeys, fsorted;
    DevBuf<int64_t> ids;
    MFB_CUDA(fkeys.alloc(nf)); MFB_CUDA(fsorted.alloc(nf)); MFB_CUDA(ids.alloc(nf));
    LAUNCH(k_face_keys, nblk(nf), TPB, conn.p, n_el, T, nv1, fkeys.p);
    MFB_CUDA(cudaMemcpyAsync(fsorted.p, fkeys.p, nf * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
    thrust::device_ptr<u64> fp(fsorted.p);
    thrust::sort(pol, fp, fp + nf);
    thrust::device_ptr<int64_t> ip(ids.p);
    IsBoundary pred{fkeys.p, fsorted.p, nf};
    B->n_bf = thrust::copy_if(pol, thrust::counting_iterator<int64_t>(0), thrust::counting_iterator<int64_t>(nf), ip, pred) - ip;
    if (B->n_bf > 0) {
        MFB_CUDA(B->f_el.alloc(B->n_bf)); MFB_CUDA(B->f_eidx.alloc(B->n_bf)); MFB_CUDA(B->cen.alloc(3 * B->n_bf));
        LAUNCH(k_bfacets, nblk(B->n_bf), TPB, ids.p, B->n_bf, n_el, conn.p, T, v1.p, v2.p, v3.p, B->f_el.p, B->f_eidx.p, B->cen.p);
    }
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    MFB_CUDA(cudaGetLastError());
    if (n_controlpoints) *n_controlpoints = B->N;
    if (n_boundary_facets) *n_boundary_facets = B->n_bf;
    return MFB_OK;
}

// copies of the built tables (any pointer may be NULL); the device-resident originals can be handed to mfb_mesh_set /
// mfb_facets_set directly through mfb_mesh_build_device_ptrs
extern "C" int mfb_mesh_build_get(mfb_ctx* ctx, int32_t* controlpoint_IDs, double* x1, double* x2, double* x3,
                                  int32_t* bfacet_element_ID, int32_t* bfacet_element_eindex, double* bfacet_centroids) {
    if (!ctx) return MFB_ERR_ARG;
    MeshBuild* B = ctx->meshbuild;
    MFB_REQUIRE(B != nullptr, MFB_ERR_STATE, "mfb_mesh_build_get: call mfb_mesh_build_second_order first");
    MFB_CUDA(cudaSetDevice(ctx->device));
    if (controlpoint_IDs) MFB_TRY(mfb_stage_out(ctx, B->cp.p, B->n_el * B->n_a * sizeof(int), controlpoint_IDs));
    if (x1) MFB_TRY(mfb_stage_out(ctx, B->x1.p, B->N * sizeof(double), x1));
    if (x2) MFB_TRY(mfb_stage_out(ctx, B->x2.p, B->N * sizeof(double), x2));
    if (x3) MFB_TRY(mfb_stage_out(ctx, B->x3.p, B->N * sizeof(double), x3));
    if (B->n_bf > 0) {
        if (bfacet_element_ID) MFB_TRY(mfb_stage_out(ctx, B->f_el.p, B->n_bf * sizeof(int), bfacet_element_ID));
        if (bfacet_element_eindex) MFB_TRY(mfb_stage_out(ctx, B->f_eidx.p, B->n_bf * sizeof(int), bfacet_element_eindex));
        if (bfacet_centroids) MFB_TRY(mfb_stage_out(ctx, B->cen.p, 3 * B->n_bf * sizeof(double), bfacet_centroids));
    }
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MFB_OK;
}

extern "C" int mfb_mesh_build_device_ptrs(mfb_ctx* ctx, const int32_t** controlpoint_IDs, const double** x1, const double** x2,
                                          const double** x3) {
    if (!ctx) return MFB_ERR_ARG;
    MeshBuild* B = ctx->meshbuild;
    MFB_REQUIRE(B != nullptr, MFB_ERR_STATE, "mfb_mesh_build_device_ptrs: call mfb_mesh_build_second_order first");
    if (controlpoint_IDs) *controlpoint_IDs = B->cp.p;
    if (x1) *x1 = B->x1.p;
    if (x2) *x2 = B->x2.p;
    if (x3) *x3 = B->x3.p;
    return MFB_OK;
}
