// mfb_dist.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// New relative to the reference (single GPU; the only trace of distribution are two TODO comments,
// src/misc/04_GPU_Utils.jl:86-87,102). Element-block partitioning with *unassembled* interface rows:
// every rank assembles only its own elements, so a node shared by several ranks carries a partial
// matrix row / residual entry on each of them. Every vector that comes out of an element loop or an
// SpMV is completed by mfb_halo_add (interface entries to the neighbours -- over peer memory, or pack ->
// grouped ncclSend/ncclRecv -> one merge kernel that sums in ascending rank order, so all copies of a shared
// entry are bit-identical). Reductions count a node on its owner only and finish with one allreduce of the whole
// scalar batch (peer-memory mailboxes, or ncclAllReduce).
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
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
    BIND(AllGather, "ncclAllGather");
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

// ---- peer-memory allreduce of the Krylov scalar batches -------------------------------------------------------------
// Every reduction of the solvers ends in k <= P2P_MAXV doubles per rank. NCCL's small-message allreduce costs 20-30 us on
// 8 GPUs; here each rank stores its partial sums straight into a slot of every peer's mailbox over NVLink (the mailboxes are
// cudaMalloc'ed, exchanged as CUDA IPC handles), publishes a sequence number behind a system-scope fence, waits for the
// sequence numbers of all peers in its own mailbox and adds the slots in rank order (bit-identical on every rank). The same
// kernel first folds the per-block partials of the multi-dot kernel, so one launch replaces k_reduce_partials + ncclAllReduce.
// Two mailboxes are used alternately: a peer can only be one round ahead, so a slot is never overwritten while it is read.
// (P2PSlot / P2PPeers: mfb_internal.h)

__global__ void __launch_bounds__(256) k_reduce_allreduce_p2p(const double* partials, int nb, int k, double* out, P2PPeers peers,
                                                              int rank, int n_ranks, unsigned long long seq, int* err) {
    __shared__ double local[P2P_MAXV];
    __shared__ double sh[8];
    const int tid = threadIdx.x;
    for (int d = 0; d < k; ++d) {                       // fold the per-block partials of dot d
        double v = 0.0;
        for (int b = tid; b < nb; b += blockDim.x) v += partials[(size_t)d * nb + b];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) sh[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) {
            double w = 0.0;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) w += sh[i];
            local[d] = w;
        }
        __syncthreads();
    }
    const int par = (int)(seq & 1ull);
    if (tid < n_ranks) {                                // publish to every mailbox (including the own one)
        P2PSlot* dst = peers.box[tid] + par * n_ranks + rank;
        for (int d = 0; d < k; ++d) ((volatile double*)dst->v)[d] = local[d];
        __threadfence_system();
        *((volatile unsigned long long*)&dst->seq) = seq;
    }
    if (tid < n_ranks) {                                // wait for every rank's contribution of this round
        const P2PSlot* src = peers.box[rank] + par * n_ranks + tid;
        long long spins = 0;
        while (*((volatile const unsigned long long*)&src->seq) < seq) {
            if (++spins > 40000000ll) { atomicExch(err, 1); break; }    // ~10 s: a peer died; do not hang the GPU
        }
        __threadfence_system();
    }
    __syncthreads();
    if (tid < k) {
        const P2PSlot* src = peers.box[rank] + par * n_ranks;
        double s = 0.0;
        for (int r = 0; r < n_ranks; ++r) s += ((volatile const double*)src[r].v)[tid];
        out[tid] = s;
    }
}

// ---- peer-memory interface exchange -----------------------------------------------------------------------------------
// The receive buffer of every rank is cudaMalloc'ed once, exported as a CUDA IPC handle and mapped by its neighbours:
// [2 parities][HALO_FLAGS flag words | total * HALO_NVMAX doubles]. k_halo_push packs the interface entries of v straight into
// the neighbours' buffers over NVLink; k_halo_signal (next kernel on the stream, i.e. after all pushes have completed)
// publishes the round number behind a system-scope fence; k_halo_merge_wait spins on the round numbers of its neighbours
// and then does the ordered merge. No NCCL call, no staging copy. Parity double-buffering: a neighbour can be at most one
// round ahead, so a buffer is never overwritten while it is merged.
constexpr int HALO_NVMAX = 6;
constexpr int HALO_MAXNB = 32;
constexpr int HALO_FLAGS = 64;     // flag words (unsigned long long) at the head of each parity region
struct HaloPeers {
    double* data[HALO_MAXNB];                 // neighbour i: data base of its receive region (parity 0)
    long long poff[HALO_MAXNB];               // neighbour i: node offset of MY entries in ITS pack order
    unsigned long long* flag[HALO_MAXNB];     // neighbour i: MY flag word inside its region (parity 0)
    long long stride[HALO_MAXNB];             // neighbour i: bytes between its parity regions
    long long off[HALO_MAXNB + 1];            // my pack offsets (nodes) per neighbour
    int n;
};
__global__ void k_halo_push(const double* v, const int* slots, int nv, HaloPeers P, int parity) {
    const int64_t total = P.off[P.n] * nv;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t node = t / nv;
        int i = 0;
        while (node >= P.off[i + 1]) ++i;
        double* dst = reinterpret_cast<double*>(reinterpret_cast<char*>(P.data[i]) + parity * P.stride[i]);
        dst[(P.poff[i] - P.off[i]) * nv + t] = v[(size_t)slots[node] * nv + (t - node * nv)];
    }
}
__global__ void k_halo_signal(HaloPeers P, int parity, unsigned long long seq) {
    const int i = threadIdx.x;
    if (i >= P.n) return;
    __threadfence_system();
    unsigned long long* f = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(P.flag[i]) + parity * P.stride[i]);
    *((volatile unsigned long long*)f) = seq;
}
__global__ void k_halo_merge_wait(double* v, const int* uni, const int* uptr, const int* uidx, const int* ulow, int64_t n_union,
                                  int nv, const double* recv, const unsigned long long* flags, int nn, unsigned long long seq,
                                  int* err) {
    if (threadIdx.x < nn) {
        long long spins = 0;
        while (*((volatile const unsigned long long*)(flags + threadIdx.x)) < seq)
            if (++spins > 40000000ll) { atomicExch(err, 1); break; }
        __threadfence_system();
    }
    __syncthreads();
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_union * nv) return;
    const int u = (int)(t / nv), k = (int)(t % nv);
    const int lo = uptr[u], hi = uptr[u + 1], nlow = ulow[u];
    const size_t at = (size_t)uni[u] * nv + k;
    double s = 0.0;
    for (int e = lo; e < lo + nlow; ++e) s += ((volatile const double*)recv)[(size_t)uidx[e] * nv + k];
    s = s + v[at];
    for (int e = lo + nlow; e < hi; ++e) s += ((volatile const double*)recv)[(size_t)uidx[e] * nv + k];
    v[at] = s;
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
    DevBuf<double> sendbuf, recvbuf;
    int buf_nv = 0;
    // peer-memory allreduce
    bool p2p = false;
    P2PSlot* mailbox = nullptr;            // own mailbox [2][n_ranks]
    P2PPeers peers;
    std::vector<void*> opened;             // IPC-opened peer mailboxes (closed in mfb_comm_free)
    unsigned long long seq = 0;
    DevBuf<int> p2p_err;
    // peer-memory interface exchange
    bool halo_p2p = false;
    char* hbuf = nullptr;                  // own receive buffer: 2 parity regions
    long long hstride = 0;                 // bytes per parity region
    HaloPeers hpeers;
    unsigned long long hseq = 0;
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
    // ---- mailboxes of the peer-memory allreduce (MFB_P2P=0 falls back to ncclAllReduce) ----
    const char* env = getenv("MFB_P2P");
    if (n_ranks > 1 && n_ranks <= P2P_MAXR && !(env && env[0] == '0')) {
        const size_t bytes = sizeof(P2PSlot) * 2 * n_ranks;
        MFB_CUDA(cudaMalloc((void**)&c->mailbox, bytes));
        MFB_CUDA(cudaMemset(c->mailbox, 0, bytes));
        MFB_CUDA(c->p2p_err.alloc(1));
        MFB_CUDA(cudaMemset(c->p2p_err.p, 0, sizeof(int)));
        cudaIpcMemHandle_t mine;
        MFB_CUDA(cudaIpcGetMemHandle(&mine, c->mailbox));
        DevBuf<unsigned char> hs, hr;
        MFB_CUDA(hs.alloc(sizeof(mine)));
        MFB_CUDA(hr.alloc(sizeof(mine) * n_ranks));
        MFB_CUDA(cudaMemcpyAsync(hs.p, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
        MFB_NCCL(api.AllGather(hs.p, hr.p, sizeof(mine), ncclChar, c->comm, ctx->stream));
        std::vector<cudaIpcMemHandle_t> all(n_ranks);
        MFB_CUDA(cudaMemcpyAsync(all.data(), hr.p, sizeof(mine) * n_ranks, cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        hs.release(); hr.release();
        bool ok = true;
        for (int r = 0; r < n_ranks; ++r) {
            if (r == rank) { c->peers.box[r] = c->mailbox; continue; }
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
            c->opened.push_back(q);
            c->peers.box[r] = (P2PSlot*)q;
        }
        // all ranks must agree: one allreduce of the success flag
        DevBuf<double> flag;
        MFB_CUDA(flag.alloc(1));
        const double f = ok ? 0.0 : 1.0;
        MFB_CUDA(cudaMemcpyAsync(flag.p, &f, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        MFB_NCCL(api.AllReduce(flag.p, flag.p, 1, ncclDouble, ncclSum, c->comm, ctx->stream));
        double tot = 1.0;
        MFB_CUDA(cudaMemcpyAsync(&tot, flag.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        flag.release();
        c->p2p = tot == 0.0;
    }
    return MFB_OK;
}

void mfb_comm_free(mfb_ctx* ctx) {
    Comm* c = ctx->comm;
    if (!c) return;
    for (void* q : c->opened) cudaIpcCloseMemHandle(q);
    if (c->mailbox) cudaFree(c->mailbox);
    if (c->hbuf) cudaFree(c->hbuf);
    c->p2p_err.release();
    if (c->comm) nccl().CommDestroy(c->comm);
    c->slots.release(); c->uni.release(); c->uptr.release(); c->uidx.release(); c->ulow.release();
    c->sendbuf.release(); c->recvbuf.release();
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
    // ---- peer-memory exchange buffers: every rank publishes its IPC handle and its (neighbour, offset) table ----
    // eligibility is decided COLLECTIVELY (a rank with too many neighbours must not skip the AllGather the others enter)
    bool halo_ok = c->p2p && c->comm && c->n_ranks > 1;
    if (halo_ok) {
        DevBuf<double> elig;
        MFB_CUDA(elig.alloc(1));
        const double bad = (n_neighbors <= HALO_MAXNB && n_neighbors <= HALO_FLAGS) ? 0.0 : 1.0;
        MFB_CUDA(cudaMemcpyAsync(elig.p, &bad, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        MFB_NCCL(nccl().AllReduce(elig.p, elig.p, 1, ncclDouble, ncclSum, c->comm, ctx->stream));
        double tot = 1.0;
        MFB_CUDA(cudaMemcpyAsync(&tot, elig.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        halo_ok = tot == 0.0;
    }
    if (halo_ok) {
        NcclApi& api = nccl();
        const int R = c->n_ranks;
        c->hstride = (long long)(HALO_FLAGS * sizeof(unsigned long long) + (size_t)(total > 0 ? total : 1) * HALO_NVMAX * sizeof(double));
        c->hstride = (c->hstride + 255) / 256 * 256;
        if (c->hbuf) { cudaFree(c->hbuf); c->hbuf = nullptr; }
        MFB_CUDA(cudaMalloc((void**)&c->hbuf, 2 * (size_t)c->hstride));
        MFB_CUDA(cudaMemset(c->hbuf, 0, 2 * (size_t)c->hstride));
        struct Rec { cudaIpcMemHandle_t h; long long stride; long long nn; long long nb[HALO_MAXNB]; long long off[HALO_MAXNB]; };
        Rec mine;
        memset(&mine, 0, sizeof(mine));
        MFB_CUDA(cudaIpcGetMemHandle(&mine.h, c->hbuf));
        mine.stride = c->hstride;
        mine.nn = n_neighbors;
        for (int i = 0; i < n_neighbors; ++i) { mine.nb[i] = c->neighbors[i]; mine.off[i] = c->offsets[i]; }
        DevBuf<unsigned char> hs, hr;
        MFB_CUDA(hs.alloc(sizeof(Rec)));
        MFB_CUDA(hr.alloc(sizeof(Rec) * R));
        MFB_CUDA(cudaMemcpyAsync(hs.p, &mine, sizeof(Rec), cudaMemcpyHostToDevice, ctx->stream));
        MFB_NCCL(api.AllGather(hs.p, hr.p, sizeof(Rec), ncclChar, c->comm, ctx->stream));
        std::vector<Rec> all(R);
        MFB_CUDA(cudaMemcpyAsync(all.data(), hr.p, sizeof(Rec) * R, cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        hs.release(); hr.release();
        bool ok = true;
        memset(&c->hpeers, 0, sizeof(c->hpeers));
        c->hpeers.n = n_neighbors;
        for (int i = 0; i <= n_neighbors; ++i) c->hpeers.off[i] = c->offsets[i];
        for (int i = 0; i < n_neighbors && ok; ++i) {
            const Rec& q = all[c->neighbors[i]];
            int me = -1;
            for (int j = 0; j < q.nn; ++j) if (q.nb[j] == c->rank) me = j;
            if (me < 0) { ok = false; break; }
            void* base = nullptr;
            if (cudaIpcOpenMemHandle(&base, q.h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
            c->opened.push_back(base);
            c->hpeers.flag[i] = reinterpret_cast<unsigned long long*>(base) + me;
            c->hpeers.data[i] = reinterpret_cast<double*>(reinterpret_cast<char*>(base) + HALO_FLAGS * sizeof(unsigned long long));
            c->hpeers.poff[i] = q.off[me];
            c->hpeers.stride[i] = q.stride;
        }
        DevBuf<double> flag;
        MFB_CUDA(flag.alloc(1));
        const double f = ok ? 0.0 : 1.0;
        MFB_CUDA(cudaMemcpyAsync(flag.p, &f, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        MFB_NCCL(api.AllReduce(flag.p, flag.p, 1, ncclDouble, ncclSum, c->comm, ctx->stream));
        double tot = 1.0;
        MFB_CUDA(cudaMemcpyAsync(&tot, flag.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        flag.release();
        const char* env = getenv("MFB_P2P_HALO");
        c->halo_p2p = tot == 0.0 && !(env && env[0] == '0');
    }
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
    if (c->halo_p2p && nv <= HALO_NVMAX) {
        c->hseq++;
        const int par = (int)(c->hseq & 1ull);
        // node j of the list shared with neighbour i is entry (off_i + j) of my pack order and entry (poff_i + j) of the
        // neighbour's: it is stored at (poff_i + j) * nv + k of the neighbour's region, where its merge kernel expects it
        const HaloPeers& P = c->hpeers;
        int64_t work = total * nv;
        unsigned grid = nblk(work);
        if (grid > 1184) grid = 1184;
        LAUNCH(k_halo_push, grid ? grid : 1, TPB, v, c->slots.p, nv, P, par);
        LAUNCH(k_halo_signal, 1, HALO_MAXNB, P, par, c->hseq);
        const char* region = c->hbuf + (size_t)par * c->hstride;
        LAUNCH(k_halo_merge_wait, nblk(c->n_union * nv), TPB, v, c->uni.p, c->uptr.p, c->uidx.p, c->ulow.p, c->n_union, nv,
               reinterpret_cast<const double*>(region + HALO_FLAGS * sizeof(unsigned long long)),
               reinterpret_cast<const unsigned long long*>(region), nn, c->hseq, c->p2p_err.p);
        MFB_CUDA(cudaGetLastError());
        return MFB_OK;
    }
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

// out[0..k) = sum over ranks of (sum_b partials[d][b]); one launch over peer memory when the mailboxes are up, else NCCL
int mfb_reduce_allreduce(mfb_ctx* ctx, const double* partials, int nb, int k, double* out) {
    Comm* c = ctx->comm;
    if (!c || !c->comm || c->n_ranks == 1 || !c->p2p || k > P2P_MAXV) return 1;     // caller uses its own reduce (+ mfb_allreduce_sum)
    c->seq++;
    LAUNCH(k_reduce_allreduce_p2p, 1, 256, partials, nb, k, out, c->peers, c->rank, c->n_ranks, c->seq, c->p2p_err.p);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

bool mfb_p2p_next(mfb_ctx* ctx, P2PInfo* out) {
    Comm* c = ctx->comm;
    if (!c || !c->comm || c->n_ranks == 1 || !c->p2p) return false;
    c->seq++;
    out->peers = c->peers;
    out->rank = c->rank;
    out->n_ranks = c->n_ranks;
    out->seq = c->seq;
    out->err = c->p2p_err.p;
    return true;
}

// MFB_ERR_NCCL if a peer-memory allreduce gave up waiting for a peer since the last check (call after a synchronisation)
int mfb_p2p_check(mfb_ctx* ctx) {
    Comm* c = ctx->comm;
    if (!c || !c->p2p) return MFB_OK;
    int h = 0;
    MFB_CUDA(cudaMemcpyAsync(&h, c->p2p_err.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    MFB_REQUIRE(h == 0, MFB_ERR_NCCL, "peer-memory allreduce timed out waiting for another rank");
    return MFB_OK;
}

bool mfb_is_distributed(mfb_ctx* ctx) { return ctx->comm && ctx->comm->comm && ctx->comm->n_ranks > 1; }
