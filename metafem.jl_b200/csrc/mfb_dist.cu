// mfb_dist.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// New relative to the reference (single GPU; the only trace of distribution are two TODO comments,
// src/misc/04_GPU_Utils.jl:86-87,102). Element-block partitioning with *unassembled* interface rows:
// every rank assembles only its own elements, so a node shared by several ranks carries a partial
// matrix row / residual entry on each of them. Every vector that comes out of an element loop or an
// SpMV is completed by mfb_halo_add (pack -> grouped ncclSend/ncclRecv with the neighbours -> sum in
// ascending rank order, so all copies of a shared entry are bit-identical). Reductions count a node on
// its owner only and finish with one ncclAllReduce of the whole scalar batch.
// NCCL is bound with dlopen (MFB_NCCL_LIB or libnccl.so.2) so the library loads on machines without it.
#include <dlfcn.h>
#include <nccl.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/reduce.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#include <algorithm>
#include <cstring>

#include "mfb_internal.h"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi& nccl() {
    static NcclApi api;
    if (api.handle || !api.error.empty()) return api;
    const char* name = getenv("MFB_NCCL_LIB");
    api.handle = dlopen(name ? name : "libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!api.handle) {
        api.error = std::string("cannot load NCCL: ") + dlerror();
        return api;
    }
#define BIND(field, sym)                                              \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym)); \
    if (!api.field) api.error = std::string("NCCL symbol missing: ") + sym;
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(AllReduce, "ncclAllReduce");
    BIND(Send, "ncclSend");
    BIND(Recv, "ncclRecv");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
    BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    return api;
}

constexpr int TPB = 256;
inline unsigned nblk(int64_t n) { return (unsigned)((n + TPB - 1) / TPB); }

__global__ void k_pack(const double* v, const int* slots, int64_t n, int nv, double* buf) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n * nv) return;
    buf[t] = v[(size_t)slots[t / nv] * nv + t % nv];
}
__global__ void k_zero_slots(double* v, const int* slots, int64_t n, int nv) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n * nv) return;
    v[(size_t)slots[t / nv] * nv + t % nv] = 0.0;
}
__global__ void k_add_slots(double* v, const int* slots, int64_t n, int nv, const double* buf) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n * nv) return;
    v[(size_t)slots[t / nv] * nv + t % nv] += buf[t];
}
// v = low + v on the union of shared nodes (contributions of lower ranks first)
__global__ void k_merge_low(double* v, const double* low, const int* slots, int64_t n, int nv) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n * nv) return;
    size_t i = (size_t)slots[t / nv] * nv + t % nv;
    v[i] = low[i] + v[i];
}
// One pass over the union of shared nodes: v[node] = sum over the sharing ranks in ascending rank order, this rank's own
// value in its place (contributions of lower ranks first, then own, then higher ranks) -- every rank holding the node
// performs the additions in the same order, so all copies stay bit-identical.
__global__ void k_halo_merge(double* v, const int* uni, const int* uptr, const int* uidx, const int* ulow, int64_t n_union,
                             int nv, const double* recv) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_union * nv) return;
    const int u = (int)(t / nv), k = (int)(t % nv);
    const int lo = uptr[u], hi = uptr[u + 1], nlow = ulow[u];
    const size_t at = (size_t)uni[u] * nv + k;
    double s = 0.0;
    for (int e = lo; e < lo + nlow; ++e) s += recv[(size_t)uidx[e] * nv + k];
    s = s + v[at];
    for (int e = lo + nlow; e < hi; ++e) s += recv[(size_t)uidx[e] * nv + k];
    v[at] = s;
}
__global__ void k_map_nodes(const int* ref_ids_1based, const int* perm, int64_t n, int* out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) out[t] = perm[ref_ids_1based[t] - 1];
}
__global__ void k_owner_internal(const unsigned char* owned_ref, const long long* gid_ref, const int* iperm, int64_t N,
                                 unsigned char* owned, long long* gid) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= N) return;
    owned[t] = owned_ref ? owned_ref[iperm[t]] : 1;
    gid[t] = gid_ref ? gid_ref[iperm[t]] - 1 : iperm[t];
}

}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, n_ranks = 1;
    std::vector<int> neighbors;            // ascending
    std::vector<int64_t> offsets;          // [n_neighbors + 1] into slots
    DevBuf<int> slots;                     // internal node ids, concatenated per neighbour
    DevBuf<int> uni;                       // unique union of slots
    DevBuf<int> uptr, uidx, ulow;          // per union node: receive-buffer positions (ascending neighbour rank), # from lower ranks
    int64_t n_union = 0;
    DevBuf<double> sendbuf, recvbuf, low;
    int buf_nv = 0;
};

#define LAUNCH(kernel, grid, block, ...)                          \
    do {                                                          \
        kernel<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__); \
        ctx->launches++;                                          \
    } while (0)

#define MFB_NCCL(call)                                                                             \
    do {                                                                                           \
        ncclResult_t _r = (call);                                                                  \
        if (_r != ncclSuccess) {                                                                   \
            ctx->err = std::string(#call) + ": " + nccl().GetErrorString(_r);                      \
            return MFB_ERR_NCCL;                                                                   \
        }                                                                                          \
    } while (0)

extern "C" int mfb_comm_unique_id(void* id128) {
    if (!id128) return MFB_ERR_ARG;
    NcclApi& api = nccl();
    if (!api.error.empty()) return MFB_ERR_NCCL;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (api.GetUniqueId(&id) != ncclSuccess) return MFB_ERR_NCCL;
    memcpy(id128, &id, 128);
    return MFB_OK;
}

extern "C" int mfb_comm_init(mfb_ctx* ctx, int rank, int n_ranks, const void* id128) {
    if (!ctx || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return MFB_ERR_ARG;
    NcclApi& api = nccl();
    MFB_REQUIRE(api.error.empty(), MFB_ERR_NCCL, api.error);
    MFB_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->comm) ctx->comm = new Comm();
    Comm* c = ctx->comm;
    c->rank = rank;
    c->n_ranks = n_ranks;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    MFB_NCCL(api.CommInitRank(&c->comm, n_ranks, id, rank));
    return MFB_OK;
}

void mfb_comm_free(mfb_ctx* ctx) {
    Comm* c = ctx->comm;
    if (!c) return;
    if (c->comm) nccl().CommDestroy(c->comm);
    c->slots.release(); c->uni.release(); c->uptr.release(); c->uidx.release(); c->ulow.release();
    c->sendbuf.release(); c->recvbuf.release(); c->low.release();
    ctx->owned.release(); ctx->gid.release();
    delete c;
    ctx->comm = nullptr;
}

// default ownership / global ids for the single-GPU case (also called by mfb_interface_set)
int mfb_node_ids_init(mfb_ctx* ctx, const unsigned char* owned_ref_dev, const long long* gid_ref_dev) {
    MFB_CUDA(ctx->owned.alloc(ctx->N));
    MFB_CUDA(ctx->gid.alloc(ctx->N));
    LAUNCH(k_owner_internal, nblk(ctx->N), TPB, owned_ref_dev, gid_ref_dev, ctx->iperm.p, ctx->N, ctx->owned.p, ctx->gid.p);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_interface_set(mfb_ctx* ctx, int n_neighbors, const int32_t* neighbor_ranks, const int64_t* offsets,
                                 const int32_t* shared_nodes, const uint8_t* owned, const int64_t* global_ids) {
    if (!ctx) return MFB_ERR_ARG;
    MFB_REQUIRE(ctx->perm.p, MFB_ERR_STATE, "mfb_interface_set: call mfb_mesh_set first");
    MFB_REQUIRE(n_neighbors >= 0 && owned && global_ids, MFB_ERR_ARG, "mfb_interface_set: bad arguments");
    MFB_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->comm) ctx->comm = new Comm();
    Comm* c = ctx->comm;
    c->neighbors.assign(neighbor_ranks, neighbor_ranks + n_neighbors);
    c->offsets.assign(offsets, offsets + n_neighbors + 1);
    for (int i = 1; i < n_neighbors; ++i)
        MFB_REQUIRE(c->neighbors[i] > c->neighbors[i - 1], MFB_ERR_ARG, "neighbour ranks must be ascending");
    const int64_t total = n_neighbors ? c->offsets[n_neighbors] : 0;
    DevBuf<int> tmp;
    DevBuf<unsigned char> own;
    DevBuf<long long> gid;
    MFB_CUDA(own.alloc(ctx->N));
    MFB_CUDA(gid.alloc(ctx->N));
    MFB_TRY(mfb_stage_in(ctx, owned, ctx->N, own.p));
    MFB_TRY(mfb_stage_in(ctx, global_ids, ctx->N * sizeof(long long), gid.p));
    MFB_TRY(mfb_node_ids_init(ctx, own.p, gid.p));
    if (total > 0) {
        MFB_CUDA(tmp.alloc(total));
        MFB_CUDA(c->slots.alloc(total));
        MFB_CUDA(c->uni.alloc(total));
        MFB_TRY(mfb_stage_in(ctx, shared_nodes, total * sizeof(int), tmp.p));
        LAUNCH(k_map_nodes, nblk(total), TPB, tmp.p, ctx->perm.p, total, c->slots.p);
        MFB_CUDA(cudaMemcpyAsync(c->uni.p, c->slots.p, total * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
        auto pol = thrust::cuda::par.on(ctx->stream);
        thrust::device_ptr<int> up(c->uni.p);
        thrust::sort(pol, up, up + total);
        c->n_union = thrust::unique(pol, up, up + total) - up;
        // merge lists: for every union node the positions of its copies in the receive buffer, neighbours ascending
        std::vector<int> hs(total), hu(c->n_union);
        MFB_CUDA(cudaMemcpyAsync(hs.data(), c->slots.p, total * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaMemcpyAsync(hu.data(), c->uni.p, c->n_union * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        std::vector<int> uptr(c->n_union + 1, 0), ulow(c->n_union, 0), uidx(total);
        auto uof = [&](int slot) { return (int)(std::lower_bound(hu.begin(), hu.end(), slot) - hu.begin()); };
        for (int64_t p = 0; p < total; ++p) uptr[uof(hs[p]) + 1]++;
        for (int64_t u = 0; u < c->n_union; ++u) uptr[u + 1] += uptr[u];
        std::vector<int> fill(uptr.begin(), uptr.end() - 1);
        for (int i = 0; i < n_neighbors; ++i)                      // ascending neighbour rank == ascending position ranges
            for (int64_t p = c->offsets[i]; p < c->offsets[i + 1]; ++p) {
                const int u = uof(hs[p]);
                uidx[fill[u]++] = (int)p;
                if (c->neighbors[i] < c->rank) ulow[u]++;
            }
        MFB_CUDA(c->uptr.alloc(c->n_union + 1)); MFB_CUDA(c->ulow.alloc(c->n_union)); MFB_CUDA(c->uidx.alloc(total));
        MFB_CUDA(cudaMemcpyAsync(c->uptr.p, uptr.data(), uptr.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        MFB_CUDA(cudaMemcpyAsync(c->ulow.p, ulow.data(), ulow.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        MFB_CUDA(cudaMemcpyAsync(c->uidx.p, uidx.data(), uidx.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    {   // global number of nodes = sum over ranks of owned nodes (normalises every residual norm)
        auto pol = thrust::cuda::par.on(ctx->stream);
        thrust::device_ptr<unsigned char> op(own.p);
        const long long mine = thrust::reduce(pol, op, op + ctx->N, (long long)0);
        DevBuf<double> cnt;
        MFB_CUDA(cnt.alloc(1));
        const double h = (double)mine;
        MFB_CUDA(cudaMemcpyAsync(cnt.p, &h, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        MFB_TRY(mfb_allreduce_sum(ctx, cnt.p, 1));
        double tot = 0;
        MFB_CUDA(cudaMemcpyAsync(&tot, cnt.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->n_global_nodes = tot;
        cnt.release();
    }
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    MFB_CUDA(cudaGetLastError());
    tmp.release(); own.release(); gid.release();
    return MFB_OK;
}

// sum over the ranks sharing each interface entry of v ([N][nv], internal layout), in ascending rank order
int mfb_halo_add(mfb_ctx* ctx, double* v, int nv) {
    Comm* c = ctx->comm;
    if (!c || !c->comm || c->neighbors.empty()) return MFB_OK;
    NcclApi& api = nccl();
    ProfScope ps(ctx, MFB_T_HALO);
    const int nn = (int)c->neighbors.size();
    const int64_t total = c->offsets[nn];
    if (c->buf_nv < nv) {
        MFB_CUDA(c->sendbuf.alloc(total * nv));
        MFB_CUDA(c->recvbuf.alloc(total * nv));
        c->buf_nv = nv;
    }
    LAUNCH(k_pack, nblk(total * nv), TPB, v, c->slots.p, total, nv, c->sendbuf.p);
    MFB_NCCL(api.GroupStart());
    for (int i = 0; i < nn; ++i) {
        const int64_t off = c->offsets[i] * nv, cnt = (c->offsets[i + 1] - c->offsets[i]) * nv;
        MFB_NCCL(api.Send(c->sendbuf.p + off, cnt, ncclDouble, c->neighbors[i], c->comm, ctx->stream));
        MFB_NCCL(api.Recv(c->recvbuf.p + off, cnt, ncclDouble, c->neighbors[i], c->comm, ctx->stream));
    }
    MFB_NCCL(api.GroupEnd());
    LAUNCH(k_halo_merge, nblk(c->n_union * nv), TPB, v, c->uni.p, c->uptr.p, c->uidx.p, c->ulow.p, c->n_union, nv, c->recvbuf.p);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

int mfb_allreduce_sum(mfb_ctx* ctx, double* dev, int n) {
    Comm* c = ctx->comm;
    if (!c || !c->comm || c->n_ranks == 1) return MFB_OK;
    MFB_NCCL(nccl().AllReduce(dev, dev, n, ncclDouble, ncclSum, c->comm, ctx->stream));
    return MFB_OK;
}

bool mfb_is_distributed(mfb_ctx* ctx) { return ctx->comm && ctx->comm->comm && ctx->comm->n_ranks > 1; }
