// mfb_krylov.cu -- bandwidth-bound linear algebra of the Newton step for sm_100a.
//
// Replaces (reference file:line)
//   mul!  = CUSPARSE.mv!                      src/misc/04_GPU_Utils.jl:131          -> k_spmv_bsr
//   iterative_Solve!                          src/solver/linear_solver/02_Preconditioner.jl:32-76
//   Pr_Jacobi! / Jacobi_By_Diagonal / Mat_Div_Jacobi     :103-148                    -> k_jacobi_diag, k_scale_copy
//   idrs! (+ modify_Omega)                    src/solver/linear_solver/04_IDRs.jl:1-95
//   bicgstabl_GS!                             src/solver/linear_solver/03_BiCGstabl.jl:18-96
//   initialize_dx!/update_dx!/update_x_star!  src/solver/04_Time_Domain.jl:20-49
// The matrix is block-CSR over the node graph (NV x NV blocks, values [entry][NV*NV]); vectors are
// node-major interleaved [node][NV]. Every vector update of a Krylov step is one fused launch
// (k_lincomb / k_axpby_batch) and every group of reductions is one fused multi-dot launch; scalars
// come back through one pinned 8*k byte copy per group instead of one blocking cuBLAS call each.
#include <cmath>
#include <cstring>

#include "mfb_internal.h"

namespace {

constexpr int TPB = 256;
constexpr int MAXT = 48;   // max terms of one fused vector update
constexpr int MAXD = 24;   // max dots of one fused reduction
constexpr int RED_BLOCKS = 1184;  // 8 x 148 SMs, 256 threads each

inline unsigned nblk(int64_t n) { return (unsigned)((n + TPB - 1) / TPB); }

// ---------------------------------------------------------------------------------------------
// SpMV (main kernel, NV*NV <= 32). One warp per block row, 8 warps walk a contiguous chunk of SPMV_ROWS rows.
//  * Lane l owns ONE fixed position (i,k) of the NV x NV block and steps through the row EPW = 32/(NV*NV) entries at a
//    time: consecutive lanes read consecutive doubles (a fully coalesced stream of the row's values), (i,k) and the x
//    component never change per lane, so per value the lane issues 3 loads + 1 DFMA and nothing else.
//  * The loop is software-pipelined: the streaming loads (values, column ids) of batch j+1 are issued BEFORE the
//    dependent x gathers of batch j are consumed, and the next row's pointers are fetched one row ahead. The round-1
//    kernel had the chain nodeptr -> (vals, cols) -> x[col] exposed on every batch and sat at 46 % of DRAM peak;
//    measured on the real 88^3 pattern: 3.16 ms -> 2.26 ms (profiles/exp/spmv_exp2.cu, profiles/spmv_experiments_r1.md).
//  * Values are streamed with evict-first loads so x stays resident in L1/L2; the Morton node order keeps the gathers local.
constexpr int SPMV_ROWS = 64;
template <int NV> struct SpmvUnroll { static constexpr int value = NV == 1 ? 2 : NV == 2 ? 3 : NV == 3 ? 5 : NV == 4 ? 6 : 8; };

template <int NV>
__global__ void __launch_bounds__(256) k_spmv_bsr(const int* __restrict__ nodeptr, const int* __restrict__ nodecol,
                                                  const double* __restrict__ K, const double* __restrict__ x,
                                                  double* __restrict__ y, int64_t N) {
    constexpr int B = NV * NV;
    constexpr int EPW = 32 / B;        // entries per warp step
    constexpr int ACTIVE = EPW * B;    // active lanes (27 of 32 for NV = 3)
    constexpr int UNR = SpmvUnroll<NV>::value;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int le = lane / B, ik = lane - le * B, i = ik / NV, k = ik - i * NV;
    const double* xk = x + k;
    const int64_t row0 = blockIdx.x * (int64_t)SPMV_ROWS + warp;
    const int64_t rend = min((blockIdx.x + 1) * (int64_t)SPMV_ROWS, N);
    if (row0 >= rend) return;
    int ps = nodeptr[row0], pt = nodeptr[row0 + 1];
    for (int64_t row = row0; row < rend; row += 8) {
        const int s = ps, deg = (lane < ACTIVE) ? pt - ps : 0;
        if (row + 8 < rend) { ps = __ldg(nodeptr + row + 8); pt = __ldg(nodeptr + row + 9); }
        const double* Kp = K + (size_t)s * B + lane;
        const int* Cp = nodecol + s + le;
        double a[UNR], v[UNR];
        int c[UNR];
        int e = le;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            a[u] = 0.0;
            const bool ok = e + u * EPW < deg;
            v[u] = ok ? __ldcs(Kp + u * ACTIVE) : 0.0;
            c[u] = ok ? __ldg(Cp + u * EPW) : 0;
        }
        while (e < deg) {
            double vn[UNR];
            int cn[UNR];
            e += UNR * EPW;
            Kp += UNR * ACTIVE;
            Cp += UNR * EPW;
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const bool ok = e + u * EPW < deg;
                vn[u] = ok ? __ldcs(Kp + u * ACTIVE) : 0.0;
                cn[u] = ok ? __ldg(Cp + u * EPW) : 0;
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) a[u] += v[u] * __ldg(xk + (size_t)c[u] * NV);
#pragma unroll
            for (int u = 0; u < UNR; ++u) { v[u] = vn[u]; c[u] = cn[u]; }
        }
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < UNR; ++u) acc += a[u];
        double t = acc;
#pragma unroll
        for (int d = 1; d < NV; ++d) t += __shfl_down_sync(0xffffffffu, acc, d);          // sum over k
        double r = t;
        if constexpr ((B & (B - 1)) == 0) {                                                // EPW is a power of two: butterfly
#pragma unroll
            for (int o = 16; o >= B; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        } else {
#pragma unroll
            for (int d = 1; d < EPW; ++d) r += __shfl_down_sync(0xffffffffu, t, d * B);   // sum over the EPW entries
        }
        if (le == 0 && k == 0 && lane < ACTIVE) y[(size_t)row * NV + i] = r;
    }
}

// SpMV (fallback, any NV): the row's values as one flat stream, (entry, i, k) recomputed per value.
template <int NV>
__global__ void __launch_bounds__(256) k_spmv_bsr_flat(const int* __restrict__ nodeptr, const int* __restrict__ nodecol,
                                                  const double* __restrict__ K, const double* __restrict__ x,
                                                  double* __restrict__ y, int64_t N) {
    constexpr int B = NV * NV;
    constexpr int UNR = 4;   // independent value / index / x loads in flight per lane
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= N) return;
    const int s = nodeptr[row], t = nodeptr[row + 1];
    const double* Kr = K + (size_t)s * B;
    const int* Cr = nodecol + s;
    const int nflat = (t - s) * B;
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0;
    for (int f0 = lane; f0 < nflat; f0 += 32 * UNR) {
        double v[UNR];
        int col[UNR], kk[UNR], ii[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int f = f0 + 32 * u;
            const bool ok = f < nflat;
            const int fc = ok ? f : 0;
            v[u] = ok ? __ldcs(Kr + fc) : 0.0;   // streamed once: keep x resident in L1/L2 instead
            const int ent = fc / B, ik = fc - ent * B;
            ii[u] = ik / NV;
            kk[u] = ik - ii[u] * NV;
            col[u] = __ldg(Cr + ent);
        }
        double xv[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) xv[u] = __ldg(x + (size_t)col[u] * NV + kk[u]);
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const double p = v[u] * xv[u];
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] += (i == ii[u]) ? p : 0.0;
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    if (lane < NV) {
        double r = acc[0];
#pragma unroll
        for (int i = 1; i < NV; ++i) r = (lane == i) ? acc[i] : r;
        y[(size_t)row * NV + lane] = r;
    }
}

// jac[node][v] = K[diag(node)][v][v] (partial on interface nodes; completed by the halo sum, then k_jacobi_abs),
// 1 when the (v,v) block is not populated (Jacobi_By_Diagonal leaves the initial 1.0)
__global__ void k_jacobi_diag(const int* nodeptr, const int* nodecol, const double* K, const int* diag_ok, int64_t N,
                              int nv, const unsigned char* owned, double* jac) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= N * nv) return;
    int64_t a = t / nv;
    int v = (int)(t % nv);
    double j = (owned == nullptr || owned[a]) ? 1.0 : 0.0;   // the default 1.0 must be counted once across ranks
    if (diag_ok[v]) {
        j = 0.0;
        int lo = nodeptr[a], hi = nodeptr[a + 1] - 1;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (nodecol[mid] < a) lo = mid + 1; else hi = mid;
        }
        if (nodecol[lo] == a) j = K[(size_t)lo * nv * nv + v * nv + v];
    }
    jac[t] = j;
}

__global__ void k_jacobi_abs(double* jac, int64_t n) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) jac[t] = fabs(jac[t]);
}

// Ks[ent][i][k] = K[ent][i][k] / jac[col][k]   (gather-copy of 02_Preconditioner.jl:35 fused with Mat_Div_Jacobi)
__global__ void k_scale_copy(const int* nodecol_of_entry, const double* K, const double* jac, int64_t total, int nv,
                             double* Ks) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= total) return;
    int B = nv * nv;
    int64_t ent = t / B;
    int k = (int)(t % B) % nv;
    Ks[t] = K[t] / jac[(size_t)nodecol_of_entry[ent] * nv + k];
}

struct LinComb {
    double* y;
    double ay;            // y = ay*y + sum c_i x_i   (ay == 0 -> y is not read)
    int n;
    double c[MAXT];
    const double* x[MAXT];
};

// optional fused squared norm of the result (partials[block]); 4 independent elements per thread for memory-level parallelism
__global__ void __launch_bounds__(TPB) k_lincomb(LinComb L, int64_t n, double* partial_norm2, const unsigned char* owned, int nv) {
    double nrm = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        double sv[4];
        int64_t idx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            idx[u] = i0 + u * stride;
            if (idx[u] >= n) idx[u] = -1;
            sv[u] = (L.ay != 0.0 && idx[u] >= 0) ? L.ay * L.y[idx[u]] : 0.0;
        }
        for (int k = 0; k < L.n; ++k) {
            const double c = L.c[k];
            const double* __restrict__ xp = L.x[k];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (idx[u] >= 0) sv[u] += c * xp[idx[u]];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (idx[u] >= 0) {
                L.y[idx[u]] = sv[u];
                if (owned == nullptr || owned[idx[u] / nv]) nrm += sv[u] * sv[u];
            }
    }
    if (partial_norm2) {
        __shared__ double sh[TPB / 32];
        for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = nrm;
        __syncthreads();
        if (threadIdx.x < 32) {
            double v = threadIdx.x < TPB / 32 ? sh[threadIdx.x] : 0.0;
            for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (threadIdx.x == 0) partial_norm2[blockIdx.x] = v;
        }
    }
}

struct AxpbyBatch {
    int n;
    double a[MAXT], b[MAXT];
    double* y[MAXT];
    const double* x[MAXT];
};
// y_k = a_k*y_k + b_k*x_k for k < n, independent updates in one launch
__global__ void __launch_bounds__(TPB) k_axpby_batch(AxpbyBatch Bt, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < n; i0 += 2 * stride) {
        const int64_t i1 = i0 + stride;
        const bool ok1 = i1 < n;
        for (int k = 0; k < Bt.n; ++k) {
            double* __restrict__ yp = Bt.y[k];
            const double* __restrict__ xp = Bt.x[k];
            const double y0 = yp[i0], x0 = xp[i0];
            const double y1 = ok1 ? yp[i1] : 0.0, x1 = ok1 ? xp[i1] : 0.0;
            yp[i0] = Bt.a[k] * y0 + Bt.b[k] * x0;
            if (ok1) yp[i1] = Bt.a[k] * y1 + Bt.b[k] * x1;
        }
    }
}

struct MultiDot {
    int n;
    const double* x[MAXD];
    const double* y[MAXD];
};
// partials[k][block] = sum over the block's grid-stride slice of x_k*y_k (warp-shuffle + smem reduction)
__global__ void __launch_bounds__(TPB) k_multidot(MultiDot M, int64_t n, double* partials, const unsigned char* owned, int nv) {
    __shared__ double sh[TPB / 32];
    for (int k0 = 0; k0 < M.n; k0 += 4) {
        double acc[4] = {0, 0, 0, 0};
        const int nk = min(4, M.n - k0);
        const int64_t stride = (int64_t)gridDim.x * blockDim.x;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += 2 * stride) {
            const int64_t i1 = i + stride;
            const bool ok1 = i1 < n;
            const double w0 = (owned == nullptr || owned[i / nv]) ? 1.0 : 0.0;        // count every node once (its owner)
            const double w1 = (ok1 && (owned == nullptr || owned[i1 / nv])) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < nk) {
                    const double a0 = M.x[k0 + k][i], b0 = M.y[k0 + k][i];
                    const double a1 = ok1 ? M.x[k0 + k][i1] : 0.0, b1 = ok1 ? M.y[k0 + k][i1] : 0.0;
                    acc[k] += w0 * (a0 * b0) + w1 * (a1 * b1);
                }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k >= nk) break;
            double v = acc[k];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
            __syncthreads();
            if (threadIdx.x < 32) {
                double w = threadIdx.x < TPB / 32 ? sh[threadIdx.x] : 0.0;
                for (int o = 4; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
                if (threadIdx.x == 0) partials[(size_t)(k0 + k) * gridDim.x + blockIdx.x] = w;
            }
        }
    }
}

// out[k] = sum_b partials[k][b]  (fixed order: deterministic run to run)
__global__ void k_reduce_partials(const double* partials, int nb, double* out) {
    __shared__ double sh[32];
    const int k = blockIdx.x;
    double v = 0.0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) v += partials[(size_t)k * nb + b];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double w = threadIdx.x < blockDim.x / 32 ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (threadIdx.x == 0) out[k] = w;
    }
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// FEM_rand: uniform [0,1) (reference uses unseeded cuRAND, 04_GPU_Utils.jl:22); counter-based and seeded here
// keyed by the GLOBAL reference node id, so the vector does not depend on the internal numbering or on the partition
__global__ void k_rand(double* p, int64_t n, unsigned long long seed, unsigned long long stream_id, const long long* gid, int nv) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long g = (unsigned long long)gid[i / nv] * nv + (unsigned long long)(i % nv);
        unsigned long long h = splitmix64(splitmix64(seed ^ (stream_id * 0xD1B54A32D192ED03ull)) + g);
        p[i] = (double)(h >> 11) * (1.0 / 9007199254740992.0);
    }
}

__global__ void k_div(double* y, const double* x, const double* d, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) y[i] = x[i] / d[i];
}


}  // namespace

#define LAUNCH(kernel, grid, block, ...)                          \
    do {                                                          \
        kernel<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__); \
        ctx->launches++;                                          \
    } while (0)

// ---------------------------------------------------------------------------------------------
int mfb_spmv_internal(mfb_ctx* ctx, const double* K, const double* x, double* y) {
    const int64_t N = ctx->N;
    unsigned grid = (unsigned)((N + SPMV_ROWS - 1) / SPMV_ROWS);
    const unsigned gridf = (unsigned)((N * 32 + 255) / 256);
    ProfScope ps(ctx, MFB_T_SPMV);
    switch (ctx->n_var) {
        case 1: LAUNCH(k_spmv_bsr<1>, grid, 256, ctx->nodeptr.p, ctx->nodecol.p, K, x, y, N); break;
        case 2: LAUNCH(k_spmv_bsr<2>, grid, 256, ctx->nodeptr.p, ctx->nodecol.p, K, x, y, N); break;
        case 3: LAUNCH(k_spmv_bsr<3>, grid, 256, ctx->nodeptr.p, ctx->nodecol.p, K, x, y, N); break;
        case 4: LAUNCH(k_spmv_bsr<4>, grid, 256, ctx->nodeptr.p, ctx->nodecol.p, K, x, y, N); break;
        case 5: LAUNCH(k_spmv_bsr<5>, grid, 256, ctx->nodeptr.p, ctx->nodecol.p, K, x, y, N); break;
        case 6: LAUNCH(k_spmv_bsr_flat<6>, gridf, 256, ctx->nodeptr.p, ctx->nodecol.p, K, x, y, N); break;
        default: ctx->err = "n_var > 6 not supported by the block SpMV"; return MFB_ERR_ARG;
    }
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

namespace {

struct Solver {
    mfb_ctx* ctx;
    int64_t n;
    const double* A;
    int spmv = 0;

    const unsigned char* mask() const { return mfb_is_distributed(ctx) ? ctx->owned.p : nullptr; }
    int mul(double* y, const double* x) {
        spmv++;
        MFB_TRY(mfb_spmv_internal(ctx, A, x, y));
        return mfb_halo_add(ctx, y, ctx->n_var);          // complete the interface rows (no-op on one GPU)
    }
    // dots: results in ctx->h_scal[0..k)
    int dots(int k, const double* const* xs, const double* const* ys) {
        MultiDot M;
        M.n = k;
        for (int i = 0; i < k; ++i) { M.x[i] = xs[i]; M.y[i] = ys[i]; }
        double* partials = ctx->scal.p + 64;
        LAUNCH(k_multidot, RED_BLOCKS, TPB, M, n, partials, mask(), ctx->n_var);
        LAUNCH(k_reduce_partials, k, 256, partials, RED_BLOCKS, ctx->scal.p);
        MFB_TRY(mfb_allreduce_sum(ctx, ctx->scal.p, k));
        MFB_CUDA(cudaMemcpyAsync(ctx->h_scal, ctx->scal.p, k * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        return MFB_OK;
    }
    int dot1(const double* x, const double* y, double* out) {
        const double* xs[1] = {x};
        const double* ys[1] = {y};
        MFB_TRY(dots(1, xs, ys));
        *out = ctx->h_scal[0];
        return MFB_OK;
    }
    // y = ay*y + sum c_i x_i ; if norm2 != nullptr also returns ||y||^2 (one extra sync)
    int lincomb(double* y, double ay, int k, const double* c, const double* const* xs, double* norm2 = nullptr) {
        LinComb L;
        L.y = y; L.ay = ay; L.n = k;
        for (int i = 0; i < k; ++i) { L.c[i] = c[i]; L.x[i] = xs[i]; }
        double* partials = norm2 ? ctx->scal.p + 64 : nullptr;
        LAUNCH(k_lincomb, RED_BLOCKS, TPB, L, n, partials, mask(), ctx->n_var);
        if (norm2) {
            LAUNCH(k_reduce_partials, 1, 256, partials, RED_BLOCKS, ctx->scal.p);
            MFB_TRY(mfb_allreduce_sum(ctx, ctx->scal.p, 1));
            MFB_CUDA(cudaMemcpyAsync(ctx->h_scal, ctx->scal.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            MFB_CUDA(cudaStreamSynchronize(ctx->stream));
            *norm2 = ctx->h_scal[0];
        }
        return MFB_OK;
    }
    int axpby_batch(int k, double* const* ys, const double* a, const double* b, const double* const* xs) {
        AxpbyBatch Bt;
        Bt.n = k;
        for (int i = 0; i < k; ++i) { Bt.y[i] = ys[i]; Bt.a[i] = a[i]; Bt.b[i] = b[i]; Bt.x[i] = xs[i]; }
        LAUNCH(k_axpby_batch, RED_BLOCKS, TPB, Bt, n);
        return MFB_OK;
    }
    double n_global = 0.0;   // global number of DOFs (sum of owned nodes over ranks)
    double nn(double norm2) const { return std::sqrt(norm2) / std::sqrt(n_global > 0 ? n_global : (double)n); }  // normalized_norm
};

// r = b - A x, returns normalized norm
int true_residual(Solver& S, double* r, const double* b, const double* x, double* res) {
    MFB_TRY(S.mul(r, x));
    double c[1] = {1.0};
    const double* xs[1] = {b};
    double n2;
    MFB_TRY(S.lincomb(r, -1.0, 1, c, xs, &n2));
    *res = S.nn(n2);
    return MFB_OK;
}

// idrs!  (04_IDRs.jl:26-95). Vectors: P[s], U[s], G[s], Ar.
int idrs(Solver& S, double* x, const double* b, double* r, double tol, int maxiter, int s, uint64_t seed, int pass,
         std::vector<double*>& W, int* iters) {
    mfb_ctx* ctx = S.ctx;
    const int64_t n = S.n;
    double res;
    MFB_TRY(true_residual(S, r, b, x, &res));
    if (res <= tol) { *iters = 0; return MFB_OK; }
    int iter = 1;
    double** P = &W[0];
    double** U = &W[s];
    double** G = &W[2 * s];
    double* Ar = W[3 * s];
    for (int k = 0; k < s; ++k) {
        LAUNCH(k_rand, RED_BLOCKS, TPB, P[k], n, (unsigned long long)seed, (unsigned long long)(pass * 64 + k), ctx->gid.p, ctx->n_var);
        MFB_CUDA(cudaMemsetAsync(U[k], 0, n * sizeof(double), ctx->stream));
        MFB_CUDA(cudaMemsetAsync(G[k], 0, n * sizeof(double), ctx->stream));
    }
    std::vector<double> M(s * s, 0.0), f(s, 0.0), c(s, 0.0);
    for (int i = 0; i < s; ++i) M[i * s + i] = 1.0;  // M[i][k] row-major
    double omega = 1.0;
    std::vector<const double*> xs(2 * s + 2), ys(s + 2);
    std::vector<double> cf(2 * s + 2);
    while (true) {
        for (int i = 0; i < s; ++i) { xs[i] = P[i]; ys[i] = r; }
        MFB_TRY(S.dots(s, xs.data(), ys.data()));
        for (int i = 0; i < s; ++i) f[i] = ctx->h_scal[i];
        for (int k = 0; k < s; ++k) {
            // c = LowerTriangular(M[k:s,k:s]) \ f[k:s]
            for (int i = k; i < s; ++i) {
                double v = f[i];
                for (int j = k; j < i; ++j) v -= M[i * s + j] * c[j];
                c[i] = v / M[i * s + i];
            }
            // U[k] = sum c_i U[i] + omega*(r - sum c_i G[i])
            int nt = 0;
            for (int i = k; i < s; ++i) {
                if (i != k) { cf[nt] = c[i]; xs[nt++] = U[i]; }
                cf[nt] = -omega * c[i]; xs[nt++] = G[i];
            }
            cf[nt] = omega; xs[nt++] = r;
            MFB_TRY(S.lincomb(U[k], c[k], nt, cf.data(), xs.data()));
            MFB_TRY(S.mul(G[k], U[k]));
            for (int i = 0; i < k; ++i) {
                double d;
                MFB_TRY(S.dot1(P[i], G[k], &d));
                double alpha = d / M[i * s + i];
                double* yy[2] = {G[k], U[k]};
                double a[2] = {1.0, 1.0}, bb[2] = {-alpha, -alpha};
                const double* xx[2] = {G[i], U[i]};
                MFB_TRY(S.axpby_batch(2, yy, a, bb, xx));
            }
            for (int i = k; i < s; ++i) { xs[i - k] = P[i]; ys[i - k] = G[k]; }
            MFB_TRY(S.dots(s - k, xs.data(), ys.data()));
            for (int i = k; i < s; ++i) M[i * s + k] = ctx->h_scal[i - k];
            double beta = f[k] / M[k * s + k];
            {
                double one = beta;
                const double* xx[1] = {U[k]};
                MFB_TRY(S.lincomb(x, 1.0, 1, &one, xx));
                double mb = -beta, n2;
                const double* gg[1] = {G[k]};
                MFB_TRY(S.lincomb(r, 1.0, 1, &mb, gg, &n2));
                res = S.nn(n2);
            }
            if (res <= tol || iter >= maxiter) { *iters = iter; return MFB_OK; }
            for (int i = k + 1; i < s; ++i) f[i] -= beta * M[i * s + k];
            iter++;
        }
        MFB_TRY(S.mul(Ar, r));
        {
            const double* a3[3] = {Ar, r, Ar};
            const double* b3[3] = {Ar, r, r};
            MFB_TRY(S.dots(3, a3, b3));
            double n1 = std::sqrt(ctx->h_scal[0]), n2 = std::sqrt(ctx->h_scal[1]), d = ctx->h_scal[2];
            const double angle = std::sqrt(2.0) / 2;
            double rho = std::fabs(d / (n1 * n2));
            omega = d / (n1 * n1);
            if (rho < angle) omega = omega * angle / rho;
        }
        {
            double om = omega, n2;
            const double* xx[1] = {r};
            MFB_TRY(S.lincomb(x, 1.0, 1, &om, xx));
            double mo = -omega;
            const double* aa[1] = {Ar};
            MFB_TRY(S.lincomb(r, 1.0, 1, &mo, aa, &n2));
            res = S.nn(n2);
        }
        if (res <= tol || iter >= maxiter) { *iters = iter; return MFB_OK; }
        iter++;
    }
}

// bicgstabl_GS!  (03_BiCGstabl.jl:18-96). Vectors: R[1..s] (R[0] = r), U[0..s], r_shadow.
int bicgstabl_gs(Solver& S, double* x, const double* b, double* r, double tol, int maxiter, int s, uint64_t seed,
                 int pass, std::vector<double*>& W, int* iters) {
    mfb_ctx* ctx = S.ctx;
    const int64_t n = S.n;
    double res;
    MFB_TRY(true_residual(S, r, b, x, &res));
    if (res <= tol) { *iters = 0; return MFB_OK; }
    int iter = 1;
    std::vector<double*> R(s + 1), U(s + 1);
    R[0] = r;
    for (int i = 1; i <= s; ++i) R[i] = W[i - 1];
    for (int i = 0; i <= s; ++i) U[i] = W[s + i];
    double* r_shadow = W[2 * s + 1];
    LAUNCH(k_rand, RED_BLOCKS, TPB, r_shadow, n, (unsigned long long)seed, (unsigned long long)(pass * 64 + 63), ctx->gid.p, ctx->n_var);
    for (int i = 1; i <= s; ++i) MFB_CUDA(cudaMemsetAsync(R[i], 0, n * sizeof(double), ctx->stream));
    for (int i = 0; i <= s; ++i) MFB_CUDA(cudaMemsetAsync(U[i], 0, n * sizeof(double), ctx->stream));
    std::vector<double> gam(s, 0.0), gamp(s, 0.0), gampp(s, 0.0), sig(s, 0.0), tau(s * s, 0.0);
    double omega = 1.0, rho0 = 1.0, alpha = 0.0;
    std::vector<double*> yy(s + 2);
    std::vector<const double*> xx(2 * s + 2);
    std::vector<double> ca(2 * s + 2), cb(2 * s + 2);
    while (true) {
        rho0 *= -omega;
        for (int j = 0; j < s; ++j) {
            double rho1;
            MFB_TRY(S.dot1(r_shadow, R[j], &rho1));
            double beta = alpha * rho1 / rho0;
            rho0 = rho1;
            for (int i = 0; i <= j; ++i) { yy[i] = U[i]; ca[i] = -beta; cb[i] = 1.0; xx[i] = R[i]; }
            MFB_TRY(S.axpby_batch(j + 1, yy.data(), ca.data(), cb.data(), xx.data()));
            MFB_TRY(S.mul(U[j + 1], U[j]));
            double d;
            MFB_TRY(S.dot1(r_shadow, U[j + 1], &d));
            alpha = rho0 / d;
            for (int i = 0; i <= j; ++i) { yy[i] = R[i]; ca[i] = 1.0; cb[i] = -alpha; xx[i] = U[i + 1]; }
            yy[j + 1] = x; ca[j + 1] = 1.0; cb[j + 1] = alpha; xx[j + 1] = U[0];
            MFB_TRY(S.axpby_batch(j + 2, yy.data(), ca.data(), cb.data(), xx.data()));
            MFB_TRY(S.mul(R[j + 1], R[j]));
        }
        for (int j = 0; j < s; ++j) {
            for (int i = 0; i < j; ++i) {
                double d;
                MFB_TRY(S.dot1(R[i + 1], R[j + 1], &d));
                tau[i * s + j] = d / sig[i];
                double mt = -tau[i * s + j];
                const double* x1[1] = {R[i + 1]};
                MFB_TRY(S.lincomb(R[j + 1], 1.0, 1, &mt, x1));
            }
            const double* a2[2] = {R[j + 1], R[0]};
            const double* b2[2] = {R[j + 1], R[j + 1]};
            MFB_TRY(S.dots(2, a2, b2));
            sig[j] = ctx->h_scal[0];
            gamp[j] = ctx->h_scal[1] / sig[j];
        }
        gam[s - 1] = gamp[s - 1];
        omega = gam[s - 1];
        for (int j = s - 2; j >= 0; --j) {
            double d = 0.0;
            for (int i = j + 1; i < s; ++i) d += tau[j * s + i] * gam[i];
            gam[j] = gamp[j] - d;
        }
        for (int j = 0; j < s - 1; ++j) {
            double d = 0.0;
            for (int i = j + 1; i < s - 1; ++i) d += tau[j * s + i] * gam[i + 1];
            gampp[j] = gam[j + 1] + d;
        }
        // x += gam[0]*R[0] + sum gampp[j]*R[j+1]      (uses R[0] before its update, as the reference does)
        int nt = 0;
        ca[nt] = gam[0]; xx[nt++] = R[0];
        for (int j = 0; j < s - 1; ++j) { ca[nt] = gampp[j]; xx[nt++] = R[j + 1]; }
        MFB_TRY(S.lincomb(x, 1.0, nt, ca.data(), xx.data()));
        // U[0] -= gam[s-1]*U[s] + sum gam[j]*U[j+1]
        nt = 0;
        ca[nt] = -gam[s - 1]; xx[nt++] = U[s];
        for (int j = 0; j < s - 1; ++j) { ca[nt] = -gam[j]; xx[nt++] = U[j + 1]; }
        MFB_TRY(S.lincomb(U[0], 1.0, nt, ca.data(), xx.data()));
        // R[0] -= gamp[s-1]*R[s] + sum gamp[j]*R[j+1]   (+ fused norm)
        nt = 0;
        ca[nt] = -gamp[s - 1]; xx[nt++] = R[s];
        for (int j = 0; j < s - 1; ++j) { ca[nt] = -gamp[j]; xx[nt++] = R[j + 1]; }
        double n2;
        MFB_TRY(S.lincomb(R[0], 1.0, nt, ca.data(), xx.data(), &n2));
        iter += s;
        if (S.nn(n2) <= tol || iter >= maxiter) { *iters = iter; return MFB_OK; }
    }
}

int ensure_scalars(mfb_ctx* ctx) {
    if (!ctx->scal.p) {
        MFB_CUDA(ctx->scal.alloc(64 + (size_t)MAXD * RED_BLOCKS));
        MFB_CUDA(cudaMallocHost((void**)&ctx->h_scal, 64 * sizeof(double)));
    }
    return MFB_OK;
}

int ensure_work(mfb_ctx* ctx, int count, int64_t n) {
    if ((int)ctx->work.size() < count) ctx->work.resize(count);
    for (int i = 0; i < count; ++i) MFB_CUDA(ctx->work[i].alloc(n));
    return MFB_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" int mfb_krylov_solve(mfb_ctx* ctx, int method, int s, int maxiter, int max_pass, double tol, uint64_t seed,
                                double* delta_out, mfb_solve_info* info) {
    if (!ctx) return MFB_ERR_ARG;
    MFB_REQUIRE(ctx->U > 0 && ctx->K_total.p, MFB_ERR_STATE, "mfb_krylov_solve: pattern/matrix not built");
    MFB_REQUIRE(method == MFB_IDRS || method == MFB_BICGSTABL_GS, MFB_ERR_ARG, "unknown Krylov method");
    MFB_REQUIRE(s >= 1 && 3 * s + 4 <= MAXT && s <= MAXD, MFB_ERR_ARG, "s out of range");
    MFB_CUDA(cudaSetDevice(ctx->device));
    MFB_TRY(ensure_scalars(ctx));
    ProfScope ps(ctx, MFB_T_SOLVE);
    const int nv = ctx->n_var;
    const int64_t n = ctx->N * nv;
    const int64_t nval = ctx->U * nv * nv;
    // workspace: [0] Ks (scaled copy), then vectors
    const int nvec = (method == MFB_IDRS ? 3 * s + 1 : 2 * s + 2) + 3;  // + x, r, (spare)
    MFB_TRY(ensure_work(ctx, nvec, n));
    DevBuf<double> Ks;  // scaled matrix copy, freed on return (the reference allocates K_vals per solve too)
    MFB_CUDA(Ks.alloc(nval));
    MFB_CUDA(ctx->jac.alloc(n));
    MFB_CUDA(ctx->delta.alloc(n));
    {
        std::vector<int> diag_ok(nv);
        for (int v = 0; v < nv; ++v) diag_ok[v] = ctx->block_of[v * nv + v] >= 0;
        DevBuf<int> d_ok;
        MFB_CUDA(d_ok.alloc(nv));
        MFB_CUDA(cudaMemcpyAsync(d_ok.p, diag_ok.data(), nv * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(k_jacobi_diag, nblk(n), TPB, ctx->nodeptr.p, ctx->nodecol.p, ctx->K_total.p, d_ok.p, ctx->N, nv,
               mfb_is_distributed(ctx) ? ctx->owned.p : (const unsigned char*)nullptr, ctx->jac.p);
        MFB_TRY(mfb_halo_add(ctx, ctx->jac.p, nv));
        LAUNCH(k_jacobi_abs, nblk(n), TPB, ctx->jac.p, n);
        LAUNCH(k_scale_copy, nblk(nval), TPB, ctx->nodecol.p, ctx->K_total.p, ctx->jac.p, nval, nv, Ks.p);
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        d_ok.release();
    }
    std::vector<double*> W(nvec - 2);
    for (int i = 0; i < nvec - 2; ++i) W[i] = ctx->work[i].p;
    double* x = ctx->work[nvec - 2].p;
    double* r = ctx->work[nvec - 1].p;
    const double* b = ctx->residue.p;
    MFB_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double), ctx->stream));
    MFB_CUDA(cudaMemcpyAsync(r, b, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    Solver S{ctx, n, Ks.p};
    S.n_global = ctx->n_global_nodes * nv;
    mfb_solve_info inf;
    memset(&inf, 0, sizeof(inf));
    {
        double d;
        MFB_TRY(S.dot1(b, b, &d));
        inf.initial_residual = S.nn(d);
    }
    int pass = 1;
    double res = inf.initial_residual;
    while (true) {
        int it = 0;
        if (method == MFB_IDRS) MFB_TRY(idrs(S, x, b, r, tol, maxiter, s, seed, pass, W, &it));
        else MFB_TRY(bicgstabl_gs(S, x, b, r, tol, maxiter, s, seed, pass, W, &it));
        inf.iterations += it;
        MFB_TRY(true_residual(S, r, b, x, &res));
        if (res < tol || pass >= max_pass) break;
        pass++;
    }
    inf.passes = pass;
    inf.residual = res;
    inf.converged = res < tol;
    inf.spmv_count = S.spmv;
    LAUNCH(k_div, nblk(n), TPB, ctx->delta.p, x, ctx->jac.p, n);   // Pr(x) = x ./ jac_vec  (:75)
    ctx->have_delta = true;
    if (delta_out) {
        DevBuf<double> tmp;
        MFB_CUDA(tmp.alloc(n));
        MFB_TRY(mfb_to_reference(ctx, ctx->delta.p, tmp.p, 1));
        MFB_TRY(mfb_stage_out(ctx, tmp.p, n * sizeof(double), delta_out));
        tmp.release();
    }
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    Ks.release();
    if (info) *info = inf;
    return inf.converged ? MFB_OK : MFB_NOT_CONVERGED;
}

extern "C" int mfb_spmv(mfb_ctx* ctx, int which, const double* x, double* y, int64_t n) {
    if (!ctx) return MFB_ERR_ARG;
    MFB_REQUIRE(ctx->U > 0, MFB_ERR_STATE, "mfb_spmv: pattern not built");
    MFB_REQUIRE(n == ctx->N * ctx->n_var, MFB_ERR_ARG, "mfb_spmv: wrong vector length");
    MFB_CUDA(cudaSetDevice(ctx->device));
    DevBuf<double> xr, xi, yi, yr;
    MFB_CUDA(xr.alloc(n)); MFB_CUDA(xi.alloc(n)); MFB_CUDA(yi.alloc(n)); MFB_CUDA(yr.alloc(n));
    MFB_TRY(mfb_stage_in(ctx, x, n * sizeof(double), xr.p));
    MFB_TRY(mfb_to_internal(ctx, xr.p, xi.p, 1));
    const double* K = which == MFB_MAT_K_LINEAR ? ctx->K_linear.p : ctx->K_total.p;
    MFB_TRY(mfb_spmv_internal(ctx, K, xi.p, yi.p));
    MFB_TRY(mfb_to_reference(ctx, yi.p, yr.p, 1));
    MFB_TRY(mfb_stage_out(ctx, yr.p, n * sizeof(double), y));
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    xr.release(); xi.release(); yi.release(); yr.release();
    return MFB_OK;
}

// ---- time stepping helpers (04_Time_Domain.jl) ------------------------------------------------
namespace {
__global__ void k_init_dx(double* dx, const double* x, int64_t n, int L1, double dt, double g0, double g1) {
    // dx = 0; for l = L..1: dx[l-1] = dt*(x[l] + gamma[l]*dx[l])   (one thread per basic DOF, all levels)
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double g[2] = {g0, g1};
    double hi = 0.0;
    dx[(size_t)(L1 - 1) * n + i] = 0.0;
    for (int l = L1 - 1; l >= 1; --l) {
        double lo = dt * (x[(size_t)l * n + i] + g[l - 1] * hi);
        dx[(size_t)(l - 1) * n + i] = lo;
        hi = lo;
    }
}
__global__ void k_x_star(double* xs, const double* x, const double* dx, int64_t n, int L1, double a0, double a1, double a2) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n * L1) return;
    double a[3] = {a0, a1, a2};
    xs[i] = x[i] + a[i / n] * dx[i];
}
__global__ void k_update_dx(double* dx, const double* delta, int64_t n, int L1, double b0, double b1, double b2) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n * L1) return;
    double b[3] = {b0, b1, b2};
    dx[i] += b[i / n] * delta[i % n];
}
__global__ void k_add(double* x, const double* dx, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) x[i] += dx[i];
}
}  // namespace

extern "C" int mfb_initialize_dx(mfb_ctx* ctx, double dt, const double* gamma, int n_gamma) {
    if (!ctx) return MFB_ERR_ARG;
    MFB_REQUIRE(ctx->x.p, MFB_ERR_STATE, "vectors not allocated (call mfb_pattern_build)");
    MFB_REQUIRE(ctx->L1 <= 3 && n_gamma >= ctx->L1 - 1, MFB_ERR_ARG, "bad gamma_params");
    int64_t n = ctx->N * ctx->n_var;
    double g0 = n_gamma > 0 ? gamma[0] : 0.0, g1 = n_gamma > 1 ? gamma[1] : 0.0;
    LAUNCH(k_init_dx, nblk(n), TPB, ctx->dx.p, ctx->x.p, n, ctx->L1, dt, g0, g1);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_update_x_star(mfb_ctx* ctx, const double* alpha, int n_alpha) {
    if (!ctx) return MFB_ERR_ARG;
    MFB_REQUIRE(ctx->x.p, MFB_ERR_STATE, "vectors not allocated (call mfb_pattern_build)");
    MFB_REQUIRE(n_alpha >= ctx->L1 && ctx->L1 <= 3, MFB_ERR_ARG, "bad alpha_params");
    int64_t n = ctx->N * ctx->n_var;
    double a[3] = {alpha[0], n_alpha > 1 ? alpha[1] : 0.0, n_alpha > 2 ? alpha[2] : 0.0};
    LAUNCH(k_x_star, nblk(n * ctx->L1), TPB, ctx->x_star.p, ctx->x.p, ctx->dx.p, n, ctx->L1, a[0], a[1], a[2]);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_update_dx(mfb_ctx* ctx, const double* beta, int n_beta, double sign) {
    if (!ctx) return MFB_ERR_ARG;
    MFB_REQUIRE(ctx->have_delta, MFB_ERR_STATE, "mfb_update_dx: no solve result available");
    MFB_REQUIRE(n_beta >= ctx->L1 && ctx->L1 <= 3, MFB_ERR_ARG, "bad beta_params");
    int64_t n = ctx->N * ctx->n_var;
    double b[3] = {sign * beta[0], n_beta > 1 ? sign * beta[1] : 0.0, n_beta > 2 ? sign * beta[2] : 0.0};
    LAUNCH(k_update_dx, nblk(n * ctx->L1), TPB, ctx->dx.p, ctx->delta.p, n, ctx->L1, b[0], b[1], b[2]);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_commit_step(mfb_ctx* ctx) {
    if (!ctx) return MFB_ERR_ARG;
    MFB_REQUIRE(ctx->x.p, MFB_ERR_STATE, "vectors not allocated (call mfb_pattern_build)");
    int64_t n = ctx->N * ctx->n_var * ctx->L1;
    LAUNCH(k_add, nblk(n), TPB, ctx->x.p, ctx->dx.p, n);
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

extern "C" int mfb_residue_norm(mfb_ctx* ctx, double* out) {
    if (!ctx || !out) return MFB_ERR_ARG;
    MFB_REQUIRE(ctx->residue.p, MFB_ERR_STATE, "vectors not allocated (call mfb_pattern_build)");
    MFB_TRY(ensure_scalars(ctx));
    Solver S{ctx, ctx->N * ctx->n_var, nullptr};
    S.n_global = ctx->n_global_nodes * ctx->n_var;
    double d;
    MFB_TRY(S.dot1(ctx->residue.p, ctx->residue.p, &d));
    *out = S.nn(d);
    return MFB_OK;
}
