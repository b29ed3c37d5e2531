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
#include <algorithm>
#include <cmath>
#include <cstring>
#include <initializer_list>

#include "mfb_internal.h"

namespace {

constexpr int TPB = 256;
constexpr int MAXT = 48;   // max terms of one fused linear combination (idrs! needs 2 s + 1)
constexpr int MAXB = 24;   // max independent updates of one batched axpby (bicgstabl needs s + 2)
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

template <int NV, int UNR = SpmvUnroll<NV>::value, int MINB = 0, int ROWS = SPMV_ROWS>
__global__ void __launch_bounds__(256, MINB) k_spmv_bsr(const int* __restrict__ nodeptr, const int* __restrict__ nodecol,
                                                        const double* __restrict__ K, const double* __restrict__ x,
                                                        double* __restrict__ y, int64_t N) {
    constexpr int B = NV * NV;
    constexpr int EPW = 32 / B;        // entries per warp step
    constexpr int ACTIVE = EPW * B;    // active lanes (27 of 32 for NV = 3)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int le = lane / B, ik = lane - le * B, i = ik / NV, k = ik - i * NV;
    const double* xk = x + k;
    const int64_t row0 = blockIdx.x * (int64_t)ROWS + warp;
    const int64_t rend = min((blockIdx.x + 1) * (int64_t)ROWS, N);
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

// Same kernel with the software pipeline running ACROSS row boundaries: while the last batch of a row is consumed, the
// first batch of the warp's next row is already in flight (its row pointers were fetched one row earlier), so the
// dependent chain nodeptr -> (values, column ids) -> x[col] is never restarted from an empty pipeline.
template <int NV, int UNR, int ROWS = SPMV_ROWS>
__global__ void __launch_bounds__(256) k_spmv_bsr_x(const int* __restrict__ nodeptr, const int* __restrict__ nodecol,
                                                    const double* __restrict__ K, const double* __restrict__ x,
                                                    double* __restrict__ y, int64_t N) {
    constexpr int B = NV * NV, EPW = 32 / B, ACTIVE = EPW * B;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int le = lane / B, ik = lane - le * B, i = ik / NV, k = ik - i * NV;
    const bool on = lane < ACTIVE;
    const double* xk = x + k;
    int64_t row = blockIdx.x * (int64_t)ROWS + warp;
    const int64_t rend = min((blockIdx.x + 1) * (int64_t)ROWS, N);
    if (row >= rend) return;
    int s = nodeptr[row], t = nodeptr[row + 1];
    int64_t nrow = row + 8;
    int ns = 0, nt = 0;
    if (nrow < rend) { ns = __ldg(nodeptr + nrow); nt = __ldg(nodeptr + nrow + 1); }
    int base = 0;                                            // first entry of the current batch within its row (warp-uniform)
    const double* Kp = K + (size_t)s * B + lane;
    const int* Cp = nodecol + s + le;
    double v[UNR], a[UNR];
    int c[UNR];
    {
        const int deg = on ? t - s : 0;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            a[u] = 0.0;
            const bool ok = le + u * EPW < deg;
            v[u] = ok ? __ldcs(Kp + u * ACTIVE) : 0.0;
            c[u] = ok ? __ldg(Cp + u * EPW) : 0;
        }
    }
    while (true) {
        const bool last = base + UNR * EPW >= t - s;         // this batch finishes the row
        int64_t row_n = row;
        int s_n = s, t_n = t, base_n = base + UNR * EPW;
        if (last) {
            row_n = nrow; s_n = ns; t_n = nt; base_n = 0;
            nrow += 8;
            if (nrow < rend) { ns = __ldg(nodeptr + nrow); nt = __ldg(nodeptr + nrow + 1); }
            Kp = K + (size_t)s_n * B + lane;
            Cp = nodecol + s_n + le;
        } else {
            Kp += UNR * ACTIVE;
            Cp += UNR * EPW;
        }
        const bool more = row_n < rend;
        double vn[UNR];
        int cn[UNR];
        {
            const int deg = (on && more) ? t_n - s_n : 0;
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const bool ok = base_n + le + u * EPW < deg;
                vn[u] = ok ? __ldcs(Kp + u * ACTIVE) : 0.0;
                cn[u] = ok ? __ldg(Cp + u * EPW) : 0;
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) a[u] += v[u] * __ldg(xk + (size_t)c[u] * NV);
        if (last) {
            double acc = 0.0;
#pragma unroll
            for (int u = 0; u < UNR; ++u) { acc += a[u]; a[u] = 0.0; }
            double tt = acc;
#pragma unroll
            for (int d = 1; d < NV; ++d) tt += __shfl_down_sync(0xffffffffu, acc, d);          // sum over k
            double r = tt;
            if constexpr ((B & (B - 1)) == 0) {
#pragma unroll
                for (int o = 16; o >= B; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
            } else {
#pragma unroll
                for (int d = 1; d < EPW; ++d) r += __shfl_down_sync(0xffffffffu, tt, d * B);   // sum over the EPW entries
            }
            if (le == 0 && k == 0 && on) y[(size_t)row * NV + i] = r;
        }
        if (!more) break;
        row = row_n; s = s_n; t = t_n; base = base_n;
#pragma unroll
        for (int u = 0; u < UNR; ++u) { v[u] = vn[u]; c[u] = cn[u]; }
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
    const double one = (owned == nullptr || owned[a]) ? 1.0 : 0.0;   // the default 1.0 must be counted once across ranks
    double j = one;
    if (diag_ok[v] && nodeptr[a] < nodeptr[a + 1]) {
        int lo = nodeptr[a], hi = nodeptr[a + 1] - 1;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (nodecol[mid] < a) lo = mid + 1; else hi = mid;
        }
        // a partitioned mesh holds the (a,a) block on every rank that holds node a (partial sums: completed by the halo add);
        // a row WITHOUT a stored diagonal block keeps the 1.0 of Jacobi_By_Diagonal (02_Preconditioner.jl:114-120)
        if (nodecol[lo] == a) j = K[(size_t)lo * nv * nv + v * nv + v];
    }
    jac[t] = j;
}

__global__ void k_jacobi_abs(double* jac, int64_t n) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) jac[t] = fabs(jac[t]);
}

// y[col][k] += sum_i K[ent][i][k] * x[row][i]  -- transposed block SpMV for lsqr! (tmul!, 06_LSQR.jl:25,42); y zeroed by the
// caller. One warp per block row, lane <-> fixed (i,k) like the forward kernel; the NV partial products of one output
// component sit NV lanes apart and are folded with shuffles before ONE red.global.add per (entry, k).
template <int NV>
__global__ void __launch_bounds__(256) k_spmv_bsr_t(const int* __restrict__ nodeptr, const int* __restrict__ nodecol,
                                                    const double* __restrict__ K, const double* __restrict__ x,
                                                    double* __restrict__ y, int64_t N) {
    constexpr int B = NV * NV, EPW = 32 / B, ACTIVE = EPW * B;
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= N) return;
    const int le = lane / B, ik = lane - le * B, i = ik / NV, k = ik - i * NV;
    const bool on = lane < ACTIVE;
    const double xi = on ? x[(size_t)row * NV + i] : 0.0;
    const int s = nodeptr[row], t = nodeptr[row + 1];
    for (int e0 = s; e0 < t; e0 += EPW) {
        const int e = e0 + le;
        const bool live = on && e < t;
        double v = live ? K[(size_t)e * B + ik] * xi : 0.0;
        // fold over i: lanes ik = i*NV + k, i = 0..NV-1 -> lane with i == 0 collects
#pragma unroll
        for (int j = 1; j < NV; ++j) {
            const double o = __shfl_down_sync(0xffffffffu, v, j * NV);
            if (i == 0) v += o;
        }
        if (live && i == 0) atomicAdd(y + (size_t)nodecol[e] * NV + k, v);
    }
}

// jac[col][k] += K[ent][i][k]^2 (Jacobi2_By_Colomn, 02_Preconditioner.jl:122-129) / jac[row][i] += K^2 (Jacobi_By_Row :169-176)
__global__ void k_jacobi_sq(const int* nodeptr, const int* nodecol, const double* K, int64_t N, int nv, int by_row, double* jac) {
    const int B = nv * nv;
    const int64_t row = blockIdx.x;
    for (int e = nodeptr[row]; e < nodeptr[row + 1]; ++e)
        for (int t = threadIdx.x; t < B; t += blockDim.x) {
            const double v = K[(size_t)e * B + t];
            const int i = t / nv, k = t - i * nv;
            atomicAdd(by_row ? jac + (size_t)row * nv + i : jac + (size_t)nodecol[e] * nv + k, v * v);
        }
}
__global__ void k_sqrt(double* v, int64_t n) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) v[t] = sqrt(v[t]);
}
__global__ void k_fill1(double* v, int64_t n) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) v[t] = 1.0;
}
__global__ void k_div_inplace(double* y, const double* d, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) y[i] /= d[i];
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
    const double* cp[MAXT];   // optional device-resident factor: term i is c[i] * (*cp[i]) * x_i  (nullptr -> c[i] * x_i)
    const double* ayp;        // optional device-resident factor of ay
};

// optional fused squared norm of the result (partials[block]); 4 independent elements per thread for memory-level parallelism
__global__ void __launch_bounds__(TPB) k_lincomb(LinComb L, int64_t n, double* partial_norm2, const unsigned char* owned, int nv) {
    double nrm = 0.0;
    const double ay = L.ayp ? L.ay * __ldg(L.ayp) : L.ay;     // (never write to the by-value parameter struct: that would spill all of it to local memory)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        double sv[4];
        int64_t idx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            idx[u] = i0 + u * stride;
            if (idx[u] >= n) idx[u] = -1;
            sv[u] = (ay != 0.0 && idx[u] >= 0) ? ay * L.y[idx[u]] : 0.0;
        }
        for (int k = 0; k < L.n; ++k) {
            const double c = L.cp[k] ? L.c[k] * __ldg(L.cp[k]) : L.c[k];
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
    double a[MAXB], b[MAXB];
    double* y[MAXB];
    const double* x[MAXB];
    const double* ap[MAXB];   // optional device-resident factors of a_k and b_k (nullptr -> 1)
    const double* bp[MAXB];
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
            const double ak = Bt.ap[k] ? Bt.a[k] * __ldg(Bt.ap[k]) : Bt.a[k];
            const double bk = Bt.bp[k] ? Bt.b[k] * __ldg(Bt.bp[k]) : Bt.b[k];
            const double y0 = yp[i0], x0 = xp[i0];
            const double y1 = ok1 ? yp[i1] : 0.0, x1 = ok1 ? xp[i1] : 0.0;
            yp[i0] = ak * y0 + bk * x0;
            if (ok1) yp[i1] = ak * y1 + bk * x1;
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


// =============================================================================================================
// Fused reductions. Every group of dot products of a Krylov step is ONE launch: the kernel that produces the vector
// (SpMV, vector update) also accumulates the dots, each block leaves its partial sums in global memory and the block that
// finishes last (ticket counter) folds them in a fixed order (bit-reproducible), sums them over the ranks through the
// peer-memory mailboxes when the mesh is partitioned, and applies the scalar recurrence that consumes them (rho/alpha/beta,
// tau/sigma/gamma, the convergence test) to the device-resident scalars. The reference does one blocking cuBLAS dot per
// scalar (03_BiCGstabl.jl:43-93, 04_IDRs.jl:47-93); round 1 of this library needed multidot + fold + one-thread kernel.
// sc layout (doubles): [0] rho0 [1] alpha [2] omega [3] beta [4] res2 [5] STOP flag [6] iteration counter [7] spare,
// then the per-method arrays (SC_SIG ... for bicgstabl_GS!, ID_* for idrs!).
enum { SC_RHO0 = 0, SC_ALPHA = 1, SC_OMEGA = 2, SC_BETA = 3, SC_RES2 = 4, SC_STOP = 5, SC_ITER = 6, SC_SIG = 8, SC_GAMP = 24,
       SC_GAM = 40, SC_GAMPP = 56, SC_TAU = 72, SC_COUNT = 72 + 16 * 16 };
enum { OP_NONE = 0, OP_BETA0, OP_BETA, OP_ALPHA, OP_TAU, OP_SIG, OP_FINAL,
       OP_IDR_FC, OP_IDR_ALPHA, OP_IDR_M, OP_IDR_RES, OP_IDR_OMEGA };
enum { IDS_OMEGA = 0, IDS_BETA = 1, IDS_ALPHA = 2, IDS_F = 8 };   // idrs! scalars: f [8..8+S), c, -omega c, M [S][S] behind them
struct ScOp { int op, i, j, off; };
struct RedCtx {
    double* sc;          // device scalars (nullptr: no stop flag, no scalar ops)
    double* red;         // reduced values
    double* partials;    // [n_dot][gridDim.x]
    unsigned* counter;   // ticket of the last-block detection (left at 0)
    double tol, inv_sqrt_n;
    int maxiter, S;
    int mode;            // 0 one GPU | 1 peer-memory allreduce in the tail | 2 local fold only (host: ncclAllReduce + k_apply_ops)
    int rank, n_ranks;
    unsigned long long seq;
    int* err;
    P2PPeers peers;
};

__device__ __forceinline__ bool stopped(const RedCtx& R) {
    return R.sc != nullptr && *reinterpret_cast<const volatile double*>(R.sc + SC_STOP) != 0.0;
}

__device__ void sc_gamma(double* sc, int S) {                                                          // 03_BiCGstabl.jl:72-81
    double* gam = sc + SC_GAM;
    const double* gamp = sc + SC_GAMP;
    const double* tau = sc + SC_TAU;
    gam[S - 1] = gamp[S - 1];
    sc[SC_OMEGA] = gam[S - 1];
    for (int j = S - 2; j >= 0; --j) {
        double d = 0.0;
        for (int i = j + 1; i < S; ++i) d += tau[j * S + i] * gam[i];
        gam[j] = gamp[j] - d;
    }
    for (int j = 0; j < S - 1; ++j) {
        double d = 0.0;
        for (int i = j + 1; i < S - 1; ++i) d += tau[j * S + i] * gam[i + 1];
        sc[SC_GAMPP + j] = gam[j + 1] + d;
    }
}

// scalar recurrences of bicgstabl_GS! (03_BiCGstabl.jl:43-93), applied by ONE thread to freshly reduced values rd[]
// c = LowerTriangular(M[k:s,k:s]) \ f[k:s] and -omega c (04_IDRs.jl:52)
__device__ void idr_csolve(double* sc, int k, int S) {
    double* f = sc + IDS_F; double* c = f + S; double* oc = f + 2 * S; const double* M = f + 3 * S;
    for (int i = k; i < S; ++i) {
        double v = f[i];
        for (int j = k; j < i; ++j) v -= M[i * S + j] * c[j];
        c[i] = v / M[i * S + i];
        oc[i] = -sc[IDS_OMEGA] * c[i];
    }
}
__device__ void sc_apply(const RedCtx& R, const ScOp& o, const double* red) {
    double* sc = R.sc;
    const int S = R.S;
    const double* rd = red + o.off;
    switch (o.op) {
        case OP_BETA0:                                   // rho0 *= -omega (:44), then beta of j = 0
            sc[SC_RHO0] *= -sc[SC_OMEGA];
            // fallthrough
        case OP_BETA: {                                  // (:46-48)
            const double rho1 = rd[0];
            sc[SC_BETA] = sc[SC_ALPHA] * rho1 / sc[SC_RHO0];
            sc[SC_RHO0] = rho1;
        } break;
        case OP_ALPHA: sc[SC_ALPHA] = sc[SC_RHO0] / rd[0]; break;                                      // (:54)
        case OP_TAU: sc[SC_TAU + o.i * S + o.j] = rd[0] / sc[SC_SIG + o.i]; break;                     // (:66)
        case OP_SIG:                                                                                   // (:69-70, 72-81)
            sc[SC_SIG + o.j] = rd[0];
            sc[SC_GAMP + o.j] = rd[1] / rd[0];
            if (o.j == S - 1) sc_gamma(sc, S);
            break;
        case OP_FINAL: {                                 // convergence test of the outer iteration (:92-93), then the next one's opening
            const double n2 = rd[0];
            sc[SC_RES2] = n2;
            const double iter = sc[SC_ITER] + S;
            sc[SC_ITER] = iter;
            const double nrm = sqrt(n2) * R.inv_sqrt_n;
            if (!(nrm > R.tol) || iter >= (double)R.maxiter) {
                sc[SC_STOP] = 1.0;                       // also on NaN (breakdown)
            } else {
                sc[SC_RHO0] *= -sc[SC_OMEGA];
                const double rho1 = rd[1];
                sc[SC_BETA] = sc[SC_ALPHA] * rho1 / sc[SC_RHO0];
                sc[SC_RHO0] = rho1;
            }
        } break;
        // ---- idrs! (04_IDRs.jl:26-95): f [IDS_F + i], c [IDS_F + S + i], -omega c [IDS_F + 2S + i], M [IDS_F + 3S + i S + k] ----
        case OP_IDR_FC:                                  // o.i == 1: f = P' r first (:47-49); then c of inner step o.j
            if (o.i == 1)
                for (int i = 0; i < S; ++i) sc[IDS_F + i] = rd[i];
            idr_csolve(sc, o.j, S);
            break;
        case OP_IDR_ALPHA: sc[IDS_ALPHA] = rd[0] / sc[IDS_F + 3 * S + o.i * S + o.i]; break;           // (:66)
        case OP_IDR_M: {                                 // M[i][k] = P[i]' G[k], i >= k (:71-73); beta (:76); f update (:85)
            double* f = sc + IDS_F; double* M = f + 3 * S;
            const int k = o.j;
            for (int i = k; i < S; ++i) M[i * S + k] = rd[i - k];
            const double beta = f[k] / M[k * S + k];
            sc[IDS_BETA] = beta;
            for (int i = k + 1; i < S; ++i) f[i] -= beta * M[i * S + k];
        } break;
        case OP_IDR_RES: {                               // convergence test after an inner step (:81-84, :92-93); o.j = next k or -1
            const double n2 = rd[0];
            sc[SC_RES2] = n2;
            const double nrm = sqrt(n2) * R.inv_sqrt_n;
            if (!(nrm > R.tol) || sc[SC_ITER] >= (double)R.maxiter) {
                sc[SC_STOP] = 1.0;
            } else {
                sc[SC_ITER] += 1.0;
                if (o.j >= 0) idr_csolve(sc, o.j, S);      // c of the next inner step
            }
        } break;
        case OP_IDR_OMEGA: {                             // modify_Omega (:1-8) from |Ar|^2, |r|^2, Ar'r
            const double n1 = sqrt(rd[0]), n2 = sqrt(rd[1]), d = rd[2];
            const double angle = 0.70710678118654752440;
            const double rho = fabs(d / (n1 * n2));
            double omega = d / (n1 * n1);
            if (rho < angle) omega = omega * angle / rho;
            sc[IDS_OMEGA] = omega;
        } break;
        default: break;
    }
}

__global__ void k_apply_ops(RedCtx R, ScOp o0, ScOp o1) {
    if (stopped(R)) return;
    if (o0.op != OP_NONE) sc_apply(R, o0, R.red);
    if (o1.op != OP_NONE) sc_apply(R, o1, R.red);
}

// sum of v over the 256 threads of the block, valid in thread 0 (sh: 8 doubles)
__device__ __forceinline__ double block_sum256(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double w = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) w += sh[i];
    }
    return w;
}

// Tail of every reducing kernel (256 threads per block; thread 0 of each block has already stored the block's partial sums at
// partials[d * gridDim.x + blockIdx.x]). The last block to arrive folds, exchanges with the other ranks, applies the scalar ops.
__device__ __noinline__ void reduce_tail(const RedCtx& R, int nd, const ScOp& o0, const ScOp& o1, int n_partials = 0) {
    __shared__ int s_last;
    __shared__ double s_w[8];
    __shared__ double s_loc[P2P_MAXV];
    const int tid = threadIdx.x;
    if (tid == 0) {
        __threadfence();
        const unsigned t = atomicAdd(R.counter, 1u);
        s_last = (t == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int nb = n_partials > 0 ? n_partials : (int)gridDim.x;    // kernels that run VIRTUAL blocks store one partial per virtual block
    for (int d = 0; d < nd; ++d) {
        double v = 0.0;
        for (int b = tid; b < nb; b += 256) v += __ldcg(R.partials + (size_t)d * nb + b);
        const double w = block_sum256(v, s_w);
        if (tid == 0) s_loc[d] = w;
    }
    __syncthreads();
    if (R.mode == 1) {
        const int par = (int)(R.seq & 1ull);
        if (tid < R.n_ranks) {                              // publish to every mailbox (including the own one)
            P2PSlot* dst = R.peers.box[tid] + par * R.n_ranks + R.rank;
            for (int d = 0; d < nd; ++d) ((volatile double*)dst->v)[d] = s_loc[d];
            __threadfence_system();
            *((volatile unsigned long long*)&dst->seq) = R.seq;
        }
        if (tid < R.n_ranks) {                              // wait for every rank's contribution of this round
            const P2PSlot* src = R.peers.box[R.rank] + par * R.n_ranks + tid;
            long long spins = 0;
            while (*((volatile const unsigned long long*)&src->seq) < R.seq) {
                if (++spins > 40000000ll) { atomicExch(R.err, 1); break; }    // ~10 s: a peer died; do not hang the GPU
            }
            __threadfence_system();
        }
        __syncthreads();
        double sum = 0.0;
        if (tid < nd) {
            const P2PSlot* src = R.peers.box[R.rank] + par * R.n_ranks;
            for (int r = 0; r < R.n_ranks; ++r) sum += ((volatile const double*)src[r].v)[tid];   // rank order: same bits everywhere
        }
        __syncthreads();
        if (tid < nd) s_loc[tid] = sum;
        __syncthreads();
    }
    if (tid == 0) {
        for (int d = 0; d < nd; ++d) R.red[d] = s_loc[d];
        if (R.mode != 2 && R.sc != nullptr) {
            if (o0.op != OP_NONE) sc_apply(R, o0, s_loc);
            if (o1.op != OP_NONE) sc_apply(R, o1, s_loc);
        }
        *R.counter = 0u;
    }
}

// ---- fused vector program: a short list of updates y_u = ay_u * y_u + sum_t c_t x_t (coefficients = host constant times an
// optional device-resident scalar), executed in order for every index, followed by dot products whose operands are vectors
// or the result of the LAST update (still in registers), the cross-block fold and the scalar ops. The programs of one outer
// iteration are built once per solve and live in device memory (8-byte kernel argument instead of a 2-3 kB struct).
constexpr int FU_MAXU = 18;   // updates per program (bicgstabl needs s + 2)
constexpr int FU_MAXT = 56;   // terms per program
constexpr int FU_MAXD = 24;   // dots per program
struct FusedTerm { const double* x; const double* cp; double c; };
struct FusedUpd { double* y; const double* ayp; double ay; int t0, nt; };
struct Fused {
    int n_upd, n_term, n_dot, pad;
    FusedUpd u[FU_MAXU];
    FusedTerm t[FU_MAXT];
    const double* dx[FU_MAXD];    // dot d = sum_i X_i * Y_i, X = dx[d] (nullptr: result of the last update), Y likewise
    const double* dy[FU_MAXD];
    ScOp ops[2];
};
static_assert(sizeof(Fused) % 8 == 0, "Fused is copied to shared memory in 8-byte words");

__device__ __forceinline__ double2 ld2(const double* p, int64_t pair, int64_t n) {
    if (2 * pair + 1 < n) return *reinterpret_cast<const double2*>(p + 2 * pair);
    return make_double2(p[2 * pair], 0.0);
}
__device__ __forceinline__ void st2(double* p, int64_t pair, int64_t n, double2 v) {
    if (2 * pair + 1 < n) *reinterpret_cast<double2*>(p + 2 * pair) = v;
    else p[2 * pair] = v.x;
}

template <int ND>
__global__ void __launch_bounds__(256) k_fused(const Fused* __restrict__ Fg, const RedCtx R, int64_t n, const unsigned char* owned,
                                               int nv) {
    if (stopped(R)) return;
    __shared__ Fused F;
    __shared__ double s_c[FU_MAXT], s_ay[FU_MAXU], s_w[8];
    __shared__ int s_single;
    const int tid = threadIdx.x;
    {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(Fg);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(&F);
        for (int i = tid; i < (int)(sizeof(Fused) / 8); i += 256) dst[i] = src[i];
    }
    __syncthreads();
    if (tid < F.n_term) s_c[tid] = F.t[tid].cp ? F.t[tid].c * __ldcg(F.t[tid].cp) : F.t[tid].c;
    if (tid < F.n_upd) s_ay[tid] = F.u[tid].ayp ? F.u[tid].ay * __ldcg(F.u[tid].ayp) : F.u[tid].ay;
    if (tid == 0) {
        int single = 1;
        for (int u = 0; u < F.n_upd; ++u) single &= (F.u[u].nt == 1);
        s_single = single;
    }
    __syncthreads();
    double acc[ND > 0 ? ND : 1];
#pragma unroll
    for (int d = 0; d < (ND > 0 ? ND : 1); ++d) acc[d] = 0.0;
    const int64_t np = (n + 1) >> 1;                      // element pairs: 128-bit loads and stores
    const int64_t stride = (int64_t)gridDim.x * 256;
    const bool single = s_single != 0;
    const double2 zero2 = make_double2(0.0, 0.0);
    for (int64_t p0 = blockIdx.x * (int64_t)256 + tid; p0 < np; p0 += stride) {
        double2 r0 = zero2;                               // result of the last update (dot operand)
        if (single) {
            // independent single-term updates y_u = ay_u y_u + c_u x_u (the BiCG part): four at a time, eight loads in flight
            for (int u0 = 0; u0 < F.n_upd; u0 += 4) {
                double2 yv[4], xv[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int u = u0 + g;
                    const bool live = u < F.n_upd;
                    yv[g] = (live && s_ay[u] != 0.0) ? ld2(F.u[u].y, p0, n) : zero2;
                    xv[g] = live ? ld2(F.t[F.u[u].t0].x, p0, n) : zero2;
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int u = u0 + g;
                    if (u < F.n_upd) {
                        const double ay = s_ay[u], c = s_c[F.u[u].t0];
                        r0 = make_double2(ay * yv[g].x + c * xv[g].x, ay * yv[g].y + c * xv[g].y);
                        st2(F.u[u].y, p0, n, r0);
                    }
                }
            }
        } else {
            for (int u = 0; u < F.n_upd; ++u) {
                double* yp = F.u[u].y;
                const double ay = s_ay[u];
                double2 a0 = (ay != 0.0) ? ld2(yp, p0, n) : zero2;
                const int t0 = F.u[u].t0, t1 = t0 + F.u[u].nt;
                double2 xv[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) xv[g] = (t0 + g < t1) ? ld2(F.t[t0 + g].x, p0, n) : zero2;
                a0.x *= ay; a0.y *= ay;
                for (int t = t0; t < t1; t += 4) {                       // four term loads in flight, the next four issued before use
                    double2 xn[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) xn[g] = (t + 4 + g < t1) ? ld2(F.t[t + 4 + g].x, p0, n) : zero2;
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        if (t + g < t1) {
                            const double c = s_c[t + g];
                            a0.x += c * xv[g].x; a0.y += c * xv[g].y;
                        }
#pragma unroll
                    for (int g = 0; g < 4; ++g) xv[g] = xn[g];
                }
                st2(yp, p0, n, a0);
                r0 = a0;
            }
        }
        if constexpr (ND > 0) {
            double2 w0 = make_double2(1.0, 2 * p0 + 1 < n ? 1.0 : 0.0);
            if (owned != nullptr) {                         // count every node once (on its owner)
                w0.x = owned[(2 * p0) / nv] ? 1.0 : 0.0;
                if (w0.y != 0.0) w0.y = owned[(2 * p0 + 1) / nv] ? 1.0 : 0.0;
            }
            constexpr int DG = ND < 4 ? ND : 4;
#pragma unroll
            for (int d0 = 0; d0 < ND; d0 += DG) {
                if (d0 < F.n_dot) {
                    double2 xv[DG], yv[DG];
#pragma unroll
                    for (int g = 0; g < DG; ++g) {
                        const int d = d0 + g;
                        const bool live = d < F.n_dot;
                        const double* xp = live ? F.dx[d] : nullptr;
                        const double* yp = live ? F.dy[d] : nullptr;
                        xv[g] = xp ? ld2(xp, p0, n) : r0;
                        yv[g] = yp ? (yp == xp ? xv[g] : ld2(yp, p0, n)) : r0;
                    }
#pragma unroll
                    for (int g = 0; g < DG; ++g)
                        if (d0 + g < F.n_dot) acc[d0 + g] += w0.x * (xv[g].x * yv[g].x) + w0.y * (xv[g].y * yv[g].y);
                }
            }
        }
    }
    if constexpr (ND > 0) {
#pragma unroll
        for (int d = 0; d < ND; ++d)
            if (d < F.n_dot) {
                const double w = block_sum256(acc[d], s_w);
                if (tid == 0) R.partials[(size_t)d * gridDim.x + blockIdx.x] = w;
            }
        reduce_tail(R, F.n_dot, F.ops[0], F.ops[1]);
    }
}

// Programs without updates (the dots that follow a product: r_shadow' A u, and the opening row of the MR part). In k_fused such a
// program has one or two 16-byte loads per thread in flight -- 57 us for a 134 MB dot, 2.3 TB/s (profiles/launches_r2g). Here
// every thread takes J pairs per step and issues all 2 J ND loads before the first multiply.
template <int ND, int J>
__global__ void __launch_bounds__(256, 3) k_dots(const Fused* __restrict__ Fg, const RedCtx R, int64_t n, const unsigned char* owned,
                                                 int nv, int vblocks) {
    if (stopped(R)) return;
    __shared__ const double* s_x[ND];
    __shared__ const double* s_y[ND];
    __shared__ ScOp s_ops[2];
    __shared__ int s_nd;
    __shared__ double s_w[8];
    const int tid = threadIdx.x;
    if (tid < ND) { s_x[tid] = Fg->dx[tid]; s_y[tid] = Fg->dy[tid]; }
    if (tid == 32) { s_ops[0] = Fg->ops[0]; s_ops[1] = Fg->ops[1]; s_nd = Fg->n_dot; }
    __syncthreads();
    const int nd = s_nd;
    const double* xp[ND];
    const double* yp[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) { xp[d] = d < nd ? s_x[d] : nullptr; yp[d] = d < nd ? s_y[d] : nullptr; }
    // The grid is ONE wave of CTAs; each CTA runs the VIRTUAL blocks vb = blockIdx.x, blockIdx.x + gridDim.x, ... of a vblocks-wide
    // grid: which elements a thread sums, in which order, and which partial they land in depend on vblocks only, so the result
    // is bit-identical for every physical grid size (the iteration count of BiCGStab hangs on these roundings)
    const int64_t np = (n + 1) >> 1;
    const int64_t stride = (int64_t)vblocks * 256;
    const double2 zero2 = make_double2(0.0, 0.0);
    for (int vb = blockIdx.x; vb < vblocks; vb += gridDim.x) {
        double acc[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) acc[d] = 0.0;
        for (int64_t p0 = vb * (int64_t)256 + tid; p0 < np; p0 += J * stride) {
            double2 xv[J][ND], yv[J][ND];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int64_t pj = p0 + j * stride;
                const bool in = pj < np;
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    xv[j][d] = (in && xp[d]) ? ld2(xp[d], pj, n) : zero2;
                    yv[j][d] = (in && yp[d]) ? (yp[d] == xp[d] ? xv[j][d] : ld2(yp[d], pj, n)) : zero2;
                }
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int64_t pj = p0 + j * stride;
                if (pj < np) {
                    double2 w0 = make_double2(1.0, 2 * pj + 1 < n ? 1.0 : 0.0);
                    if (owned != nullptr) {                     // count every node once (on its owner)
                        w0.x = owned[(2 * pj) / nv] ? 1.0 : 0.0;
                        if (w0.y != 0.0) w0.y = owned[(2 * pj + 1) / nv] ? 1.0 : 0.0;
                    }
#pragma unroll
                    for (int d = 0; d < ND; ++d) acc[d] += w0.x * (xv[j][d].x * yv[j][d].x) + w0.y * (xv[j][d].y * yv[j][d].y);
                }
            }
        }
#pragma unroll
        for (int d = 0; d < ND; ++d)
            if (d < nd) {
                const double w = block_sum256(acc[d], s_w);
                if (tid == 0) R.partials[(size_t)d * vblocks + vb] = w;
            }
    }
    reduce_tail(R, nd, s_ops[0], s_ops[1], vblocks);
}

// Light programs: at most two single-term updates y_u = ay y_u + c x_u followed by at most four dots (the first BiCG updates,
// every row of the MR part). k_fused has 2-5 loads per thread in flight for them; here a thread takes J pairs per step and
// issues all their loads first (NU updates, ND dots and J are template parameters so that the register arrays fit). The host checks (light_ok) that no load can observe a store of the same program: the update
// targets are distinct and are read neither by a later update nor -- through a pointer -- by a dot.
template <int NU, int ND, int J>
__global__ void __launch_bounds__(256, 3) k_light(const Fused* __restrict__ Fg, const RedCtx R, int64_t n, const unsigned char* owned,
                                                  int nv, int vblocks) {
    if (stopped(R)) return;
    __shared__ double* s_y[NU];
    __shared__ const double* s_x[NU];
    __shared__ double s_ay[NU], s_c[NU];
    __shared__ const double* s_dx[ND > 0 ? ND : 1];
    __shared__ const double* s_dy[ND > 0 ? ND : 1];
    __shared__ ScOp s_ops[2];
    __shared__ int s_cnt[2];
    __shared__ double s_w[8];
    const int tid = threadIdx.x;
    if (tid < NU && tid < Fg->n_upd) {
        const FusedUpd u = Fg->u[tid];
        const FusedTerm t = Fg->t[u.t0];
        s_y[tid] = u.y;
        s_x[tid] = t.x;
        s_ay[tid] = u.ayp ? u.ay * __ldcg(u.ayp) : u.ay;
        s_c[tid] = t.cp ? t.c * __ldcg(t.cp) : t.c;
    }
    if (tid >= 32 && tid < 32 + ND && tid - 32 < Fg->n_dot) { s_dx[tid - 32] = Fg->dx[tid - 32]; s_dy[tid - 32] = Fg->dy[tid - 32]; }
    if (tid == 64) { s_ops[0] = Fg->ops[0]; s_ops[1] = Fg->ops[1]; s_cnt[0] = Fg->n_upd; s_cnt[1] = Fg->n_dot; }
    __syncthreads();
    const int nu = s_cnt[0], nd = s_cnt[1];
    double* yp[NU];
    const double* xp[NU];
    double ay[NU], c[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const bool live = u < nu;
        yp[u] = live ? s_y[u] : nullptr; xp[u] = live ? s_x[u] : nullptr;
        ay[u] = live ? s_ay[u] : 0.0; c[u] = live ? s_c[u] : 0.0;
    }
    const double* dxp[ND > 0 ? ND : 1];
    const double* dyp[ND > 0 ? ND : 1];
#pragma unroll
    for (int d = 0; d < ND; ++d) { dxp[d] = d < nd ? s_dx[d] : nullptr; dyp[d] = d < nd ? s_dy[d] : nullptr; }
    const int64_t np = (n + 1) >> 1;
    const int64_t stride = (int64_t)vblocks * 256;                    // virtual blocks, as in k_dots
    const double2 zero2 = make_double2(0.0, 0.0);
    for (int vb = blockIdx.x; vb < vblocks; vb += gridDim.x) {
        double acc[ND > 0 ? ND : 1];
#pragma unroll
        for (int d = 0; d < ND; ++d) acc[d] = 0.0;
        for (int64_t p0 = vb * (int64_t)256 + tid; p0 < np; p0 += J * stride) {
            double2 yv[J][NU], xv[J][NU], dxv[J][ND > 0 ? ND : 1], dyv[J][ND > 0 ? ND : 1];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int64_t pj = p0 + j * stride;
                const bool in = pj < np;
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    yv[j][u] = (in && yp[u] && ay[u] != 0.0) ? ld2(yp[u], pj, n) : zero2;
                    xv[j][u] = (in && xp[u]) ? ld2(xp[u], pj, n) : zero2;
                }
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    dxv[j][d] = (in && dxp[d]) ? ld2(dxp[d], pj, n) : zero2;
                    dyv[j][d] = (in && dyp[d] && dyp[d] != dxp[d]) ? ld2(dyp[d], pj, n) : zero2;
                }
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int64_t pj = p0 + j * stride;
                if (pj >= np) continue;
                double2 r0 = zero2;                               // result of the last update (dot operand)
#pragma unroll
                for (int u = 0; u < NU; ++u)
                    if (u < nu) {
                        r0 = make_double2(ay[u] * yv[j][u].x + c[u] * xv[j][u].x, ay[u] * yv[j][u].y + c[u] * xv[j][u].y);
                        st2(yp[u], pj, n, r0);
                    }
                if constexpr (ND > 0) {
                    double2 w0 = make_double2(1.0, 2 * pj + 1 < n ? 1.0 : 0.0);
                    if (owned != nullptr) {                       // count every node once (on its owner)
                        w0.x = owned[(2 * pj) / nv] ? 1.0 : 0.0;
                        if (w0.y != 0.0) w0.y = owned[(2 * pj + 1) / nv] ? 1.0 : 0.0;
                    }
#pragma unroll
                    for (int d = 0; d < ND; ++d)
                        if (d < nd) {
                            const double2 X = dxp[d] ? dxv[j][d] : r0;
                            const double2 Y = dyp[d] ? (dyp[d] == dxp[d] ? X : dyv[j][d]) : r0;
                            acc[d] += w0.x * (X.x * Y.x) + w0.y * (X.y * Y.y);
                        }
                }
            }
        }
        if constexpr (ND > 0) {
#pragma unroll
            for (int d = 0; d < ND; ++d)
                if (d < nd) {
                    const double w = block_sum256(acc[d], s_w);
                    if (tid == 0) R.partials[(size_t)d * vblocks + vb] = w;
                }
        }
    }
    if constexpr (ND > 0) reduce_tail(R, nd, s_ops[0], s_ops[1], vblocks);
}

// ---- SpMV, multi-row streams: one warp walks the CONCATENATED value stream of RW consecutive block rows (they are contiguous
// in memory) with the same software pipeline as k_spmv_bsr, so the dependent chain nodeptr -> (values, column ids) -> x[col]
// is restarted once per RW rows (~ 8 x 4 batches) instead of once per row (~ 4 batches: about one exposed DRAM latency in
// five). Row boundaries inside a batch are resolved with predicated adds on the (rare) slow path; a batch that lies inside one
// row takes the FMA fast path. With DOT the kernel also accumulates sum_rows w[row] . y[row] (the dot product the Krylov
// method takes of the fresh product: r_shadow' A u) and finishes it in its tail: no second pass over y, no extra launch.
// load flavours of the SpMV streams (template parameter LDK): 0 = ld.global.cs (evict-first, still allocates in L1),
// 1 = L1::no_allocate for the value and column streams (they are read once: keep L1 for the x gathers; the L2::evict_first
// qualifier is accepted by ptxas for 256-bit vector loads only),
// 2 = as 1 and the x gathers marked L1::evict_last
template <int LDK> __device__ __forceinline__ double ld_stream_f64(const double* p) {
    if constexpr (LDK == 0) return __ldcs(p);
    double v;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
template <int LDK> __device__ __forceinline__ double ld_stream_val(const double* p) { return ld_stream_f64<LDK>(p); }
template <int LDK> __device__ __forceinline__ float ld_stream_val(const float* p) { return __ldcs(p); }            // FP32-stored factors
template <int LDK> __device__ __forceinline__ int ld_stream_s32(const int* p) {
    if constexpr (LDK == 0) return __ldg(p);
    int v;
    asm volatile("ld.global.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// LDK = 3: the gathered vector is asked to stay in L2 (evict_last policy) while the factor / matrix streams pass through
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <int LDK> __device__ __forceinline__ double ld_gather_f64(const double* p, unsigned long long pol = 0) {
    if constexpr (LDK == 3) {
        double v;
        asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
        return v;
    } else if constexpr (LDK != 2) {
        return __ldg(p);
    } else {
        double v;
        asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p));
        return v;
    }
}
__device__ __forceinline__ void st_hint_f64(double* p, double v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

// the stream loop of one warp (its RW consecutive rows)
// OUT = 0: y[row] = sum (the product). OUT = 1 / 2: one level of a triangular sweep on packed factors (mfb_ilu.cu): the "rows" are
// positions in level order, rowid[] their node ids, x == y == v:  v[id] -= sum  /  v[id] = dinv[id] v[id] - sum (rows pre-scaled)
template <int NV, int UNR, int RW, int LDK, int OUT = 0, typename VT = double>
__device__ __forceinline__ void spmv_mr_rows(const int* __restrict__ nodeptr, const int* __restrict__ nodecol, const VT* __restrict__ K,
                                          const double* x, double* y, int64_t N, int64_t row0,
                                          const int* __restrict__ rowid = nullptr, const double* __restrict__ dinv = nullptr) {
    static_assert(RW <= 31, "lane l holds the row pointer of row l: at most 31 rows per warp");
    constexpr int B = NV * NV, EPW = 32 / B, ACTIVE = EPW * B, W = UNR * EPW;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int le = lane / B, ik = lane - le * B, i = ik / NV, k = ik - i * NV;
    const bool on = lane < ACTIVE;
    unsigned long long pol = 0;
    if constexpr (LDK == 3) pol = l2_policy_evict_last();
    const double* xk = x + k;
    const int nr = (int)min((int64_t)RW, N - row0);
    const int myp = (lane <= nr) ? __ldg(nodeptr + row0 + lane) : 0;     // lane l holds nodeptr[row0 + l]
    // sweeps: the row's own entries of v (and its inverse diagonal block) are fetched now, while the stream runs -- a load at the
    // row end would stall the warp once per (short) row. Lane q < nr * NV holds component q % NV of row q / NV.
    int myid = 0;
    double myv = 0.0;
    if constexpr (OUT != 0) {
        static_assert(OUT == 0 || RW * NV <= 32, "own entries of the warp's rows are held one per lane");
        myid = (lane < nr) ? __ldg(rowid + row0 + lane) : 0;
        const int idq = __shfl_sync(FULL, myid, (lane / NV) & 31);
        const bool has = lane < nr * NV;
        myv = has ? y[(size_t)idq * NV + lane % NV] : 0.0;
        if constexpr (OUT == 2) {                                        // the upper entries are pre-scaled: own value <- U_ii^-1 v_i, once per warp
            double t = 0.0;
#pragma unroll
            for (int m = 0; m < NV; ++m) {
                const double d = has ? __ldg(dinv + (size_t)idq * B + (lane % NV) * NV + m) : 0.0;
                t = fma(d, __shfl_sync(FULL, myv, ((lane / NV) * NV + m) & 31), t);
            }
            myv = t;
        }
    }
    const int s = __shfl_sync(FULL, myp, 0);
    const int degw = __shfl_sync(FULL, myp, nr) - s;                     // length of the whole stream (warp-uniform)
    const int deg = on ? degw : 0;
    int cr = 0, lo = 0, hi = __shfl_sync(FULL, myp, 1) - s;              // current row and its entry range within the stream
    const VT* Kp = K + (size_t)s * B + lane;
    const int* Cp = nodecol + s + le;
    double a[UNR];
    VT v[UNR];                                                             // values stay in their storage type until they are used
    int c[UNR];
    int e = le, base = 0;
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
        a[u] = 0.0;
        const bool ok = e + u * EPW < deg;
        v[u] = ok ? ld_stream_val<LDK>(Kp + u * ACTIVE) : VT(0);
        c[u] = ok ? ld_stream_s32<LDK>(Cp + u * EPW) : 0;
    }
    for (;;) {
        VT vn[UNR];
        int cn[UNR];
        Kp += UNR * ACTIVE;
        Cp += UNR * EPW;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {                                  // streams of the NEXT batch first ...
            const bool ok = e + W + u * EPW < deg;
            vn[u] = ok ? ld_stream_val<LDK>(Kp + u * ACTIVE) : VT(0);
            cn[u] = ok ? ld_stream_s32<LDK>(Cp + u * EPW) : 0;
        }
        double xg[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) xg[u] = ld_gather_f64<LDK>(xk + (size_t)c[u] * NV, pol);   // ... then the dependent gathers of this one
        const int bend = base + W;
        if (bend < hi) {                                                  // the whole batch lies inside row cr
#pragma unroll
            for (int u = 0; u < UNR; ++u) a[u] = fma((double)v[u], xg[u], a[u]);
        } else {                                                          // the batch reaches the end of row cr
            double p[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) p[u] = (double)v[u] * xg[u];
            for (;;) {
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int eu = e + u * EPW;
                    if (eu >= lo && eu < hi) a[u] += p[u];
                }
                double acc = 0.0;
#pragma unroll
                for (int u = 0; u < UNR; ++u) { acc += a[u]; a[u] = 0.0; }
                double t = acc;
#pragma unroll
                for (int d = 1; d < NV; ++d) t += __shfl_down_sync(FULL, acc, d);           // sum over k
                double r = t;
                if constexpr ((B & (B - 1)) == 0) {
#pragma unroll
                    for (int o = 16; o >= B; o >>= 1) r += __shfl_xor_sync(FULL, r, o);
                } else {
#pragma unroll
                    for (int d = 1; d < EPW; ++d) r += __shfl_down_sync(FULL, t, d * B);    // sum over the EPW entries
                }
                if constexpr (OUT == 0) {
                    if (le == 0 && k == 0 && on) y[(size_t)(row0 + cr) * NV + i] = r;
                } else {
                    const int id = __shfl_sync(FULL, myid, cr);
                    const double yi = __shfl_sync(FULL, myv, (cr * NV + i) & 31) - r;
                    if (le == 0 && k == 0 && on) {
                        if constexpr (LDK == 3) st_hint_f64(y + (size_t)id * NV + i, yi, pol);
                        else y[(size_t)id * NV + i] = yi;
                    }
                }
                ++cr;
                lo = hi;
                if (cr >= nr) break;
                hi = __shfl_sync(FULL, myp, cr + 1) - s;
                if (bend < hi) {                                          // row cr goes on in later batches: take its share of this one
#pragma unroll
                    for (int u = 0; u < UNR; ++u)
                        if (e + u * EPW >= lo) a[u] += p[u];
                    break;
                }
            }
            if (cr >= nr) break;
        }
        base = bend;
        e += W;
#pragma unroll
        for (int u = 0; u < UNR; ++u) { v[u] = vn[u]; c[u] = cn[u]; }
    }
}

// ---- SpMV, TMA ring: the value / column streams of a warp's RW rows are pulled into a warp-private shared-memory ring by
// cp.async.bulk (1-D TMA) copies that complete on mbarriers, NSTG stages of E entries ahead of the consumer. The stream no
// longer occupies registers or waits for the gathers: the bytes in flight per SM are set by the ring (2 CTAs x 8 warps x 2
// stages x 4.5 kB = 146 kB for NV = 3) instead of by warps x unroll (43 kB in k_spmv_bsr, which measured at the 5.6 TB/s that
// 43 kB per SM and ~1.1 us of loaded DRAM latency allow). The consumer reads values and column ids from shared memory and
// keeps the x gathers of the NEXT batch in flight while it multiplies the current one.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
                 "r"(bytes), "r"(smem_addr(b))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned long long* b, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_addr(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    for (long long spins = 0; !mbar_try(b, parity); ++spins)
        if (spins > (1ll << 26)) __trap();                 // a copy that never lands must not hang the GPU
}

template <int NV, int UNR, int NBATCH = 4> struct TmaStage {
    static constexpr int B = NV * NV, EPW = 32 / B, W = UNR * EPW;
    static constexpr int NB = NBATCH;                      // batches per stage
    static constexpr int E = NB * W;                       // entries per stage (multiple of 4: 16-byte aligned copies)
    static constexpr int VB = E * B * 8, CB = E * 4, SB = VB + CB;
    static_assert(E % 4 == 0 && VB % 16 == 0 && CB % 16 == 0, "bulk copies need 16-byte granularity");
};

template <int NV, int UNR, int RW, int NSTG, int WARPS, int NBATCH = 4, int MINB = 0>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_spmv_tma(const int* __restrict__ nodeptr, const int* __restrict__ nodecol,
                                                         const double* __restrict__ K, const double* __restrict__ x,
                                                         double* __restrict__ y, int64_t N, const double* __restrict__ sc) {
    if (sc != nullptr && *reinterpret_cast<const volatile double*>(sc + SC_STOP) != 0.0) return;
    using T = TmaStage<NV, UNR, NBATCH>;
    constexpr int B = T::B, EPW = T::EPW, ACTIVE = EPW * B, W = T::W, NB = T::NB, E = T::E, VB = T::VB, SB = T::SB;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char tma_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* ring = tma_smem + (size_t)warp * NSTG * SB;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(tma_smem + (size_t)WARPS * NSTG * SB) + warp * NSTG;
    if (lane == 0)
        for (int q = 0; q < NSTG; ++q) mbar_init(bars + q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int64_t row0 = (blockIdx.x * (int64_t)WARPS + warp) * RW;
    if (row0 >= N) return;
    const int le = lane / B, ik = lane - le * B, i = ik / NV, k = ik - i * NV;
    const bool on = lane < ACTIVE;
    const double* xk = x + k;
    const int nr = (int)min((int64_t)RW, N - row0);
    const int myp = (lane <= nr) ? __ldg(nodeptr + row0 + lane) : 0;     // lane l holds nodeptr[row0 + l]
    const int s = __shfl_sync(FULL, myp, 0);
    const int degw = __shfl_sync(FULL, myp, nr) - s;                     // entries of the warp's stream
    if (degw <= 0) {
        if (lane < nr * NV) y[(size_t)row0 * NV + lane] = 0.0;
        return;
    }
    const int lead = s & 3;                                              // the copies start at a multiple of 4 entries
    const size_t a0 = (size_t)(s - lead);
    const int nstg = (lead + degw + E - 1) / E, nq = nstg * NB;
    auto issue = [&](int g) {                                            // stage g of the stream -> ring slot g % NSTG
        if (lane == 0) {
            const int slot = g % NSTG;
            const size_t e0 = a0 + (size_t)g * E;
            mbar_expect(bars + slot, (unsigned)SB);
            bulk_g2s(ring + (size_t)slot * SB, K + e0 * B, (unsigned)VB, bars + slot);
            bulk_g2s(ring + (size_t)slot * SB + VB, nodecol + e0, (unsigned)T::CB, bars + slot);
        }
    };
    for (int g = 0; g < NSTG && g < nstg; ++g) issue(g);
    int cr = 0, lo = 0, hi = __shfl_sync(FULL, myp, 1) - s;              // current row and its entry range (relative to s)
    double a[UNR], xg[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) a[u] = 0.0;
    // gathers of batch q: entries (relative to s) q * W - lead + le + u * EPW
    auto gather = [&](int q, double (&xo)[UNR]) {
        const int g = q / NB, b = q - g * NB;
        const int* Cs = reinterpret_cast<const int*>(ring + (size_t)(g % NSTG) * SB + VB) + b * W + le;
        const int e = q * W - lead + le;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int eu = e + u * EPW;
            const bool ok = on && eu >= 0 && eu < degw;
            const int c = ok ? Cs[u * EPW] : 0;
            xo[u] = __ldg(xk + (size_t)c * NV);
        }
    };
    mbar_wait(bars, 0);
    gather(0, xg);
    for (int q = 0; q < nq; ++q) {
        const int g = q / NB, b = q - g * NB;
        double xn[UNR];
        if (q + 1 < nq) {
            if (b == NB - 1) mbar_wait(bars + (g + 1) % NSTG, (unsigned)(((g + 1) / NSTG) & 1));
            gather(q + 1, xn);
        }
        const double* Vs = reinterpret_cast<const double*>(ring + (size_t)(g % NSTG) * SB) + (size_t)(b * W) * B + lane;
        const int e = q * W - lead + le;
        double p[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int eu = e + u * EPW;
            const bool ok = on && eu >= 0 && eu < degw;
            p[u] = ok ? Vs[u * ACTIVE] * xg[u] : 0.0;
        }
        const int bend = q * W - lead + W;
        if (bend < hi) {
#pragma unroll
            for (int u = 0; u < UNR; ++u) a[u] += p[u];
        } else {
            for (;;) {
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int eu = e + u * EPW;
                    if (eu >= lo && eu < hi) a[u] += p[u];
                }
                double acc = 0.0;
#pragma unroll
                for (int u = 0; u < UNR; ++u) { acc += a[u]; a[u] = 0.0; }
                double t = acc;
#pragma unroll
                for (int d = 1; d < NV; ++d) t += __shfl_down_sync(FULL, acc, d);
                double r = t;
                if constexpr ((B & (B - 1)) == 0) {
#pragma unroll
                    for (int o = 16; o >= B; o >>= 1) r += __shfl_xor_sync(FULL, r, o);
                } else {
#pragma unroll
                    for (int d = 1; d < EPW; ++d) r += __shfl_down_sync(FULL, t, d * B);
                }
                if (le == 0 && k == 0 && on) y[(size_t)(row0 + cr) * NV + i] = r;
                ++cr;
                lo = hi;
                if (cr >= nr) break;
                hi = __shfl_sync(FULL, myp, cr + 1) - s;
                if (bend < hi) {
#pragma unroll
                    for (int u = 0; u < UNR; ++u)
                        if (e + u * EPW >= lo) a[u] += p[u];
                    break;
                }
            }
        }
        if (b == NB - 1) {                                               // stage g is consumed: refill its slot
            __syncwarp();
            if (g + NSTG < nstg) issue(g + NSTG);
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) xg[u] = xn[u];
        if (cr >= nr) break;
    }
}

template <int NV, int UNR, int RW, bool DOT, int MINB = 0, int LDK = 0>
__global__ void __launch_bounds__(256, MINB) k_spmv_mr(const int* __restrict__ nodeptr, const int* __restrict__ nodecol,
                                                 const double* __restrict__ K, const double* __restrict__ x,
                                                 double* __restrict__ y, int64_t N, const double* __restrict__ wdot,
                                                 double* __restrict__ sc, double* __restrict__ red, double* __restrict__ partials,
                                                 unsigned* __restrict__ counter, const int opcode) {
    if (sc != nullptr && *reinterpret_cast<const volatile double*>(sc + SC_STOP) != 0.0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double dacc = 0.0;
    const int64_t row0 = (blockIdx.x * (int64_t)8 + warp) * RW;
    if (row0 < N) {
        spmv_mr_rows<NV, UNR, RW, LDK>(nodeptr, nodecol, K, x, y, N, row0);
        if constexpr (DOT) {
            // the dot product of the warp's fresh rows: the nr * NV values were written by lanes of this warp a moment ago
            __syncwarp();
            const int m = (int)min((int64_t)RW, N - row0) * NV;
#pragma unroll 1
            for (int q = lane; q < m; q += 32) {
                const size_t at = (size_t)row0 * NV + q;
                dacc += __ldcg(y + at) * __ldg(wdot + at);
            }
        }
    }
    if constexpr (DOT) {
        // lean tail (one GPU only: a partitioned product needs its interface rows completed before any dot): per-block partial,
        // ticket, the last block folds in a fixed order and applies the alpha / beta recurrence (03_BiCGstabl.jl:46-48,54)
        __shared__ double s_w[8];
        __shared__ int s_last;
        const double w = block_sum256(dacc, s_w);
        if (threadIdx.x == 0) {
            partials[blockIdx.x] = w;
            __threadfence();
            s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double v = 0.0;
#pragma unroll 2
            for (int b = threadIdx.x; b < (int)gridDim.x; b += 256) v += __ldcg(partials + b);
            const double tot = block_sum256(v, s_w);
            if (threadIdx.x == 0) {
                red[0] = tot;
                if (opcode == OP_ALPHA) {
                    sc[SC_ALPHA] = sc[SC_RHO0] / tot;
                } else if (opcode == OP_BETA) {
                    sc[SC_BETA] = sc[SC_ALPHA] * tot / sc[SC_RHO0];
                    sc[SC_RHO0] = tot;
                }
                *counter = 0u;
            }
        }
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

// ---- device-resident scalar recurrences of bicgstabl_GS! (03_BiCGstabl.jl:43-93): one thread, enqueued between the
// vector kernels so that an outer iteration needs ONE host synchronisation (the convergence test) instead of 19.
// sc layout: [0] rho0 [1] alpha [2] omega [3] beta [4] res2 [8 + j] sig [24 + j] gamp [40 + j] gam [56 + j] gampp [72 + i*S + j] tau
__global__ void k_sc_init(double* sc) {
    for (int i = threadIdx.x; i < SC_COUNT; i += blockDim.x) sc[i] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0) { sc[SC_RHO0] = 1.0; sc[SC_OMEGA] = 1.0; sc[SC_ALPHA] = 0.0; sc[SC_ITER] = 1.0; }
}
__global__ void k_sc_outer_begin(double* sc) { sc[SC_RHO0] *= -sc[SC_OMEGA]; }                        // rho0 *= -omega (:44)
__global__ void k_sc_beta(double* sc, const double* red) {                                            // (:46-48)
    const double rho1 = red[0];
    sc[SC_BETA] = sc[SC_ALPHA] * rho1 / sc[SC_RHO0];
    sc[SC_RHO0] = rho1;
}
__global__ void k_sc_alpha(double* sc, const double* red) { sc[SC_ALPHA] = sc[SC_RHO0] / red[0]; }   // (:54)
__global__ void k_sc_tau(double* sc, const double* red, int i, int j, int S) { sc[SC_TAU + i * S + j] = red[0] / sc[SC_SIG + i]; }  // (:66)
__global__ void k_sc_sig(double* sc, const double* red, int j) {                                      // (:69-70)
    sc[SC_SIG + j] = red[0];
    sc[SC_GAMP + j] = red[1] / red[0];
}
__global__ void k_sc_gamma(double* sc, int S) {                                                       // (:72-81)
    double* gam = sc + SC_GAM;
    const double* gamp = sc + SC_GAMP;
    const double* tau = sc + SC_TAU;
    gam[S - 1] = gamp[S - 1];
    sc[SC_OMEGA] = gam[S - 1];
    for (int j = S - 2; j >= 0; --j) {
        double d = 0.0;
        for (int i = j + 1; i < S; ++i) d += tau[j * S + i] * gam[i];
        gam[j] = gamp[j] - d;
    }
    for (int j = 0; j < S - 1; ++j) {
        double d = 0.0;
        for (int i = j + 1; i < S - 1; ++i) d += tau[j * S + i] * gam[i + 1];
        sc[SC_GAMPP + j] = gam[j + 1] + d;
    }
}

// ---- device-resident scalars of idrs! (04_IDRs.jl:26-95). Layout: [0] omega [1] beta [2] alpha [8 + i] f [8 + S + i] c
// [8 + 2S + i] -omega*c [8 + 3S + i*S + k] M[i][k]
enum { ID_OMEGA = 0, ID_BETA = 1, ID_ALPHA = 2, ID_F = 8 };
__global__ void k_idr_init(double* sc, int S) {
    for (int i = threadIdx.x; i < 8 + 3 * S + S * S; i += blockDim.x) sc[i] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {
        sc[ID_OMEGA] = 1.0;
        sc[6] = 1.0;                                                           // SC_ITER (the fused path counts on the device)
        for (int i = 0; i < S; ++i) sc[8 + 3 * S + i * S + i] = 1.0;           // M = I (:41)
    }
}
__global__ void k_idr_f(double* sc, const double* red, int S) {                  // f = P' r (:47-49)
    for (int i = threadIdx.x; i < S; i += blockDim.x) sc[ID_F + i] = red[i];
}
__global__ void k_idr_c(double* sc, int k, int S) {                              // c = LowerTriangular(M[k:s,k:s]) \ f[k:s] (:52)
    double* f = sc + ID_F; double* c = sc + ID_F + S; double* oc = sc + ID_F + 2 * S; const double* M = sc + ID_F + 3 * S;
    for (int i = k; i < S; ++i) {
        double v = f[i];
        for (int j = k; j < i; ++j) v -= M[i * S + j] * c[j];
        c[i] = v / M[i * S + i];
        oc[i] = -sc[ID_OMEGA] * c[i];
    }
}
__global__ void k_idr_alpha(double* sc, const double* red, int i, int S) { sc[ID_ALPHA] = red[0] / sc[ID_F + 3 * S + i * S + i]; }   // (:66)
__global__ void k_idr_M(double* sc, const double* red, int k, int S) {           // M[i][k] = P[i]' G[k], i >= k (:71-73); beta; f update
    double* f = sc + ID_F; double* M = sc + ID_F + 3 * S;
    for (int i = k; i < S; ++i) M[i * S + k] = red[i - k];
    const double beta = f[k] / M[k * S + k];
    sc[ID_BETA] = beta;
    for (int i = k + 1; i < S; ++i) f[i] -= beta * M[i * S + k];                  // (:85)
}
__global__ void k_idr_omega(double* sc, const double* red) {                     // modify_Omega (:1-8)
    const double n1 = sqrt(red[0]), n2 = sqrt(red[1]), d = red[2];
    const double angle = 0.70710678118654752440;
    const double rho = fabs(d / (n1 * n2));
    double omega = d / (n1 * n1);
    if (rho < angle) omega = omega * angle / rho;
    sc[ID_OMEGA] = omega;
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
int mfb_spmv_t_internal(mfb_ctx* ctx, const double* K, const double* x, double* y) {
    const int64_t N = ctx->N;
    const unsigned grid = (unsigned)((N * 32 + 255) / 256);
    MFB_CUDA(cudaMemsetAsync(y, 0, N * ctx->n_var * sizeof(double), ctx->stream));
    ProfScope ps(ctx, MFB_T_SPMV);
    switch (ctx->n_var) {
        case 1: LAUNCH(k_spmv_bsr_t<1>, grid, 256, ctx->nodeptr.p, ctx->nodecol.p, K, x, y, N); break;
        case 2: LAUNCH(k_spmv_bsr_t<2>, grid, 256, ctx->nodeptr.p, ctx->nodecol.p, K, x, y, N); break;
        case 3: LAUNCH(k_spmv_bsr_t<3>, grid, 256, ctx->nodeptr.p, ctx->nodecol.p, K, x, y, N); break;
        case 4: LAUNCH(k_spmv_bsr_t<4>, grid, 256, ctx->nodeptr.p, ctx->nodecol.p, K, x, y, N); break;
        case 5: LAUNCH(k_spmv_bsr_t<5>, grid, 256, ctx->nodeptr.p, ctx->nodecol.p, K, x, y, N); break;
        default: ctx->err = "n_var > 5 not supported by the transposed block SpMV"; return MFB_ERR_ARG;
    }
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

int mfb_spmv_kind(mfb_ctx* ctx) {
    (void)ctx;
    static int kind = -1;
    if (kind < 0) {
        const char* e = getenv("MFB_SPMV");
        kind = (e && (e[0] == 'r' || e[0] == '0')) ? 0 : (e && e[0] == 't') ? 2 : 1;   // row: one warp per row (round-1 kernel) | tma: TMA ring | default: multi-row streams
    }
    return kind;
}

namespace {
int spmv_rw() {
    static int rw = -1;
    if (rw < 0) {
        const char* e = getenv("MFB_SPMV_RW");
        rw = e ? atoi(e) : 16;                                     // 16 rows per warp measured best (profiles/spmv_experiments_r2.md)
        if (rw != 4 && rw != 8) rw = 16;
    }
    return rw;
}
template <int NV, int RW, int NSTG, int WARPS, int NBATCH = 4, int MINB = 0>
void launch_tma(mfb_ctx* ctx, const double* K, const double* x, double* y, const double* sc) {
    constexpr int UNR = SpmvUnroll<NV>::value;
    using T = TmaStage<NV, UNR, NBATCH>;
    const int smem = WARPS * NSTG * T::SB + WARPS * NSTG * 8;
    auto kern = k_spmv_tma<NV, UNR, RW, NSTG, WARPS, NBATCH, MINB>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    const int64_t N = ctx->N;
    const unsigned grid = (unsigned)((N + (int64_t)WARPS * RW - 1) / ((int64_t)WARPS * RW));
    kern<<<grid, WARPS * 32, smem, ctx->stream>>>(ctx->nodeptr.p, ctx->nodecol.p, K, x, y, N, sc);
    ctx->launches++;
}
RedCtx null_redctx() {
    RedCtx R;
    memset(&R, 0, sizeof(R));
    return R;
}
template <int NV, int RW>
void launch_spmv_mr(mfb_ctx* ctx, const double* K, const double* x, double* y, const double* w, const RedCtx& R, const ScOp& op) {
    const int64_t N = ctx->N;
    const unsigned grid = (unsigned)((N + 8 * RW - 1) / (8 * RW));
    // the register cap keeps the fused-dot variant at the occupancy of the plain one (its fold / scalar-op tail, executed by one
    // block, may spill instead)
    constexpr int MINB = NV <= 3 ? 4 : (NV == 4 ? 3 : 2);
    if (w) k_spmv_mr<NV, SpmvUnroll<NV>::value, RW, true, MINB><<<grid, 256, 0, ctx->stream>>>(ctx->nodeptr.p, ctx->nodecol.p, K, x, y, N, w, R.sc, R.red, R.partials, R.counter, op.op);
    else k_spmv_mr<NV, SpmvUnroll<NV>::value, RW, false, MINB><<<grid, 256, 0, ctx->stream>>>(ctx->nodeptr.p, ctx->nodecol.p, K, x, y, N, w, R.sc, R.red, R.partials, R.counter, op.op);
    ctx->launches++;
}
template <int NV, int UNR, int RW, int OUT, int MINB, typename VT, int LDK>
__global__ void __launch_bounds__(256, MINB) k_sweep_mr(const int* __restrict__ ptr, const int* __restrict__ col, const VT* __restrict__ val,
                                                        double* v, int w0, int w1, const int* __restrict__ rowid, const double* __restrict__ dinv) {
    const int64_t row0 = w0 + (blockIdx.x * (int64_t)8 + (threadIdx.x >> 5)) * RW;
    if (row0 < w1) spmv_mr_rows<NV, UNR, RW, LDK, OUT, VT>(ptr, col, val, v, v, (int64_t)w1, row0, rowid, dinv);
}
template <int NV, int RW, int UNR, int MINB, int LDK = 0>
void launch_sweep_mr(mfb_ctx* ctx, bool upper, const int* ptr, const int* col, const void* val, bool f32, const int* rowid, const double* dinv,
                     double* v, int w0, int w1) {
    const unsigned grid = (unsigned)((w1 - w0 + 8 * RW - 1) / (8 * RW));
    const float* vf = static_cast<const float*>(val);
    const double* vd = static_cast<const double*>(val);
    if (f32) {
        if (upper) k_sweep_mr<NV, UNR, RW, 2, MINB, float, LDK><<<grid, 256, 0, ctx->stream>>>(ptr, col, vf, v, w0, w1, rowid, dinv);
        else k_sweep_mr<NV, UNR, RW, 1, MINB, float, LDK><<<grid, 256, 0, ctx->stream>>>(ptr, col, vf, v, w0, w1, rowid, dinv);
    } else {
        if (upper) k_sweep_mr<NV, UNR, RW, 2, MINB, double, LDK><<<grid, 256, 0, ctx->stream>>>(ptr, col, vd, v, w0, w1, rowid, dinv);
        else k_sweep_mr<NV, UNR, RW, 1, MINB, double, LDK><<<grid, 256, 0, ctx->stream>>>(ptr, col, vd, v, w0, w1, rowid, dinv);
    }
    ctx->launches++;
}
}  // namespace

// one level [w0, w1) of a triangular sweep on the packed factors of mfb_ilu.cu with the multi-row stream kernel of the SpMV
// (the level's rows are contiguous in the packed arrays and independent of one another); `val` holds doubles or floats.
// false: n_var without this kernel.
bool mfb_sweep_level_mr(mfb_ctx* ctx, bool upper, const int* ptr, const int* col, const void* val, bool f32, const int* rowid,
                        const double* dinv, double* v, int w0, int w1) {
    static const int cfg = [] { const char* e = getenv("MFB_ILU_CFG"); return e ? atoi(e) : 0; }();   // tuning aid (profiles/ilu_r2.md)
    switch (ctx->n_var) {
        case 1: launch_sweep_mr<1, 8, 2, 4, 3>(ctx, upper, ptr, col, val, f32, rowid, dinv, v, w0, w1); return true;
        case 2: launch_sweep_mr<2, 8, 3, 4, 3>(ctx, upper, ptr, col, val, f32, rowid, dinv, v, w0, w1); return true;
        case 3:
            switch (cfg) {
                case 1: launch_sweep_mr<3, 8, 4, 4, 3>(ctx, upper, ptr, col, val, f32, rowid, dinv, v, w0, w1); break;
                case 2: launch_sweep_mr<3, 8, 5, 4, 0>(ctx, upper, ptr, col, val, f32, rowid, dinv, v, w0, w1); break;   // without the L2 hint
                case 3: launch_sweep_mr<3, 8, 5, 3, 3>(ctx, upper, ptr, col, val, f32, rowid, dinv, v, w0, w1); break;
                case 4: launch_sweep_mr<3, 4, 5, 4, 3>(ctx, upper, ptr, col, val, f32, rowid, dinv, v, w0, w1); break;
                default: launch_sweep_mr<3, 8, 5, 4, 3>(ctx, upper, ptr, col, val, f32, rowid, dinv, v, w0, w1); break;
            }
            return true;
        case 4: launch_sweep_mr<4, 8, 5, 3, 3>(ctx, upper, ptr, col, val, f32, rowid, dinv, v, w0, w1); return true;
        default: return false;
    }
}

// y = K x (block rows of this rank). w != nullptr (multi-row kernel only): also red[0] = sum_rows w . y, folded and followed by
// the scalar op in the kernel's tail; R.sc != nullptr: the launch is skipped on the device once the solver's stop flag is up.
static int spmv_launch(mfb_ctx* ctx, const double* K, const double* x, double* y, const double* w, const RedCtx& R, const ScOp& op) {
    const int64_t N = ctx->N;
    ProfScope ps(ctx, MFB_T_SPMV);
    if (mfb_spmv_kind(ctx) == 2 && ctx->n_var == 3 && w == nullptr) {
        launch_tma<3, 8, 3, 8>(ctx, K, x, y, R.sc);
        MFB_CUDA(cudaGetLastError());
        return MFB_OK;
    }
    if (mfb_spmv_kind(ctx) >= 1 && ctx->n_var <= 5) {
        const int rw = spmv_rw();
        switch (ctx->n_var) {
            case 1: launch_spmv_mr<1, 8>(ctx, K, x, y, w, R, op); break;
            case 2: launch_spmv_mr<2, 8>(ctx, K, x, y, w, R, op); break;
            case 3:
                if (rw == 4) launch_spmv_mr<3, 4>(ctx, K, x, y, w, R, op);
                else if (rw == 16) launch_spmv_mr<3, 16>(ctx, K, x, y, w, R, op);
                else launch_spmv_mr<3, 8>(ctx, K, x, y, w, R, op);
                break;
            case 4: launch_spmv_mr<4, 8>(ctx, K, x, y, w, R, op); break;
            case 5: launch_spmv_mr<5, 8>(ctx, K, x, y, w, R, op); break;
        }
        MFB_CUDA(cudaGetLastError());
        return MFB_OK;
    }
    MFB_REQUIRE(w == nullptr, MFB_ERR_STATE, "fused SpMV dot needs the multi-row kernel");
    unsigned grid = (unsigned)((N + SPMV_ROWS - 1) / SPMV_ROWS);
    const unsigned gridf = (unsigned)((N * 32 + 255) / 256);
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

int mfb_spmv_internal(mfb_ctx* ctx, const double* K, const double* x, double* y) {
    return spmv_launch(ctx, K, x, y, nullptr, null_redctx(), ScOp{OP_NONE, 0, 0, 0});
}

namespace {

struct Solver {
    mfb_ctx* ctx;
    int64_t n;
    const double* A;
    int spmv = 0;

    const double* pl = nullptr;   // Pl_Jacobi vector (left preconditioner: b ./= jac_vec after every product), or null
    bool pl_ilu = false;          // Pl_ILU: b <- U^-1 L^-1 b after every product (mfb_ilu.cu)
    bool has_pl() const { return pl != nullptr || pl_ilu; }

    const unsigned char* mask() const { return mfb_is_distributed(ctx) ? ctx->owned.p : nullptr; }
    int Pl(double* v) {
        if (pl) LAUNCH(k_div_inplace, nblk(n), TPB, v, pl, n);
        if (pl_ilu) MFB_TRY(mfb_ilu_apply(ctx, v));
        return MFB_OK;
    }
    // y = Pl(A x): every mul! of the reference's solvers is followed by Pl(.)
    int mul(double* y, const double* x) {
        spmv++;
        MFB_TRY(mfb_spmv_internal(ctx, A, x, y));
        MFB_TRY(mfb_halo_add(ctx, y, ctx->n_var));        // complete the interface rows (no-op on one GPU)
        return Pl(y);
    }
    // y = Pl(A' x)  (tmul!, used by lsqr!)
    int tmul(double* y, const double* x) {
        spmv++;
        MFB_TRY(mfb_spmv_t_internal(ctx, A, x, y));
        MFB_TRY(mfb_halo_add(ctx, y, ctx->n_var));
        return Pl(y);
    }
    // fold the per-block partials of k reductions into ctx->scal[0..k) and sum them over the ranks
    int finish(int k) {
        double* partials = ctx->scal.p + 64;
        const int rc = mfb_reduce_allreduce(ctx, partials, RED_BLOCKS, k, ctx->scal.p);
        if (rc != 1) return rc;
        LAUNCH(k_reduce_partials, k, 256, partials, RED_BLOCKS, ctx->scal.p);
        return mfb_allreduce_sum(ctx, ctx->scal.p, k);
    }
    // dots: results in ctx->h_scal[0..k)
    int dots(int k, const double* const* xs, const double* const* ys) {
        MultiDot M;
        M.n = k;
        for (int i = 0; i < k; ++i) { M.x[i] = xs[i]; M.y[i] = ys[i]; }
        double* partials = ctx->scal.p + 64;
        LAUNCH(k_multidot, RED_BLOCKS, TPB, M, n, partials, mask(), ctx->n_var);
        MFB_TRY(finish(k));
        MFB_CUDA(cudaMemcpyAsync(ctx->h_scal, ctx->scal.p, k * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        return MFB_OK;
    }
    // same reduction, results stay on the device in ctx->scal[0..k) (no host synchronisation)
    int dots_dev(int k, const double* const* xs, const double* const* ys) {
        ProfScope ps(ctx, MFB_T_REDUCE);
        MultiDot M;
        M.n = k;
        for (int i = 0; i < k; ++i) { M.x[i] = xs[i]; M.y[i] = ys[i]; }
        double* partials = ctx->scal.p + 64;
        LAUNCH(k_multidot, RED_BLOCKS, TPB, M, n, partials, mask(), ctx->n_var);
        return finish(k);
    }
    // y = ay*y + sum (c_i * *cp_i) x_i with device-resident factors; optional fused ||y||^2 left in ctx->scal[0]
    int lincomb_dev(double* y, double ay, int k, const double* c, const double* const* cp, const double* const* xs,
                    bool norm2 = false, const double* ayp = nullptr) {
        LinComb L;
        L.y = y; L.ay = ay; L.n = k; L.ayp = ayp;
        for (int i = 0; i < k; ++i) { L.c[i] = c[i]; L.cp[i] = cp[i]; L.x[i] = xs[i]; }
        double* partials = norm2 ? ctx->scal.p + 64 : nullptr;
        LAUNCH(k_lincomb, RED_BLOCKS, TPB, L, n, partials, mask(), ctx->n_var);
        if (norm2) MFB_TRY(finish(1));
        return MFB_OK;
    }
    int axpby_batch_dev(int k, double* const* ys, const double* a, const double* const* ap, const double* b,
                        const double* const* bp, const double* const* xs) {
        AxpbyBatch Bt;
        Bt.n = k;
        for (int i = 0; i < k; ++i) { Bt.y[i] = ys[i]; Bt.a[i] = a[i]; Bt.ap[i] = ap[i]; Bt.b[i] = b[i]; Bt.bp[i] = bp[i]; Bt.x[i] = xs[i]; }
        LAUNCH(k_axpby_batch, RED_BLOCKS, TPB, Bt, n);
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
        for (int i = 0; i < k; ++i) { L.c[i] = c[i]; L.x[i] = xs[i]; L.cp[i] = nullptr; }
        L.ayp = nullptr;
        double* partials = norm2 ? ctx->scal.p + 64 : nullptr;
        LAUNCH(k_lincomb, RED_BLOCKS, TPB, L, n, partials, mask(), ctx->n_var);
        if (norm2) {
            MFB_TRY(finish(1));
            MFB_CUDA(cudaMemcpyAsync(ctx->h_scal, ctx->scal.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            MFB_CUDA(cudaStreamSynchronize(ctx->stream));
            *norm2 = ctx->h_scal[0];
        }
        return MFB_OK;
    }
    int axpby_batch(int k, double* const* ys, const double* a, const double* b, const double* const* xs) {
        AxpbyBatch Bt;
        Bt.n = k;
        for (int i = 0; i < k; ++i) { Bt.y[i] = ys[i]; Bt.a[i] = a[i]; Bt.b[i] = b[i]; Bt.x[i] = xs[i]; Bt.ap[i] = Bt.bp[i] = nullptr; }
        LAUNCH(k_axpby_batch, RED_BLOCKS, TPB, Bt, n);
        return MFB_OK;
    }
    double n_global = 0.0;   // global number of DOFs (sum of owned nodes over ranks)
    double nn(double norm2) const { return std::sqrt(norm2) / std::sqrt(n_global > 0 ? n_global : (double)n); }  // normalized_norm

    // ---- fused reductions (device-resident scalars, no host synchronisation) ----
    double f_tol = 0.0;
    int f_maxiter = 0, f_S = 0;
    std::vector<Fused> progs;          // host copies of the fused programs of this solve
    // reduction context of the NEXT reducing launch (advances the mailbox round on a partitioned mesh)
    RedCtx redctx(bool with_sc = true) {
        RedCtx R;
        memset(&R, 0, sizeof(R));
        R.sc = with_sc ? ctx->ksc.p : nullptr;
        R.red = ctx->scal.p;
        R.partials = ctx->scal.p + 64;
        R.counter = ctx->red_counter.p;
        R.tol = f_tol;
        R.inv_sqrt_n = 1.0 / std::sqrt(n_global > 0 ? n_global : (double)n);
        R.maxiter = f_maxiter;
        R.S = f_S;
        R.mode = 0;
        if (mfb_is_distributed(ctx)) {
            P2PInfo I;
            if (mfb_p2p_next(ctx, &I)) {
                R.mode = 1; R.rank = I.rank; R.n_ranks = I.n_ranks; R.seq = I.seq; R.err = I.err; R.peers = I.peers;
            } else {
                R.mode = 2;
            }
        }
        return R;
    }
    RedCtx stopctx() {                   // for launches that only honour the stop flag
        RedCtx R;
        memset(&R, 0, sizeof(R));
        R.sc = ctx->ksc.p;
        return R;
    }
    // NCCL fallback of a reducing launch: sum red[0..nd) over the ranks, then the scalar ops in a one-thread kernel
    int after_reduce(const RedCtx& R, int nd, const ScOp& o0, const ScOp& o1) {
        if (R.mode != 2) return MFB_OK;
        MFB_TRY(mfb_allreduce_sum(ctx, ctx->scal.p, nd));
        if (R.sc && (o0.op != OP_NONE || o1.op != OP_NONE)) LAUNCH(k_apply_ops, 1, 1, R, o0, o1);
        return MFB_OK;
    }
    int upload_programs() {
        MFB_CUDA(ctx->kprog.alloc(progs.size() * sizeof(Fused)));
        MFB_CUDA(cudaMemcpyAsync(ctx->kprog.p, progs.data(), progs.size() * sizeof(Fused), cudaMemcpyHostToDevice, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));     // progs may be rebuilt by the next pass
        return MFB_OK;
    }
    // k_light's conditions: 1-2 single-term updates, <= 4 dots, and no load of the program can see one of its stores
    static bool light_ok(const Fused& F) {
        static const bool off = [] { const char* e = getenv("MFB_NO_LIGHT"); return e && e[0] == '1'; }();
        if (off || F.n_upd < 1 || F.n_upd > 2 || F.n_dot > 4) return false;
        for (int u = 0; u < F.n_upd; ++u) {
            if (F.u[u].nt != 1) return false;
            for (int v = 0; v < F.n_upd; ++v) {
                if (v != u && F.u[v].y == F.u[u].y) return false;
                if (F.t[F.u[v].t0].x == F.u[u].y) return false;          // also x == own y: (ay + c) y is legal but not worth a case
            }
            for (int d = 0; d < F.n_dot; ++d)
                if (F.dx[d] == F.u[u].y || F.dy[d] == F.u[u].y) return false;
        }
        return true;
    }
    int run(int idx) {
        const Fused& F = progs[idx];
        const Fused* dev = reinterpret_cast<const Fused*>(ctx->kprog.p) + idx;
        const int nd = F.n_dot;
        const bool light = light_ok(F);
        // k_dots / k_light hold 3 CTAs per SM (launch bounds): ONE wave of 3 x 148 CTAs runs the RED_BLOCKS virtual blocks, so that
        // the per-CTA prologue is paid once and no partial wave trails (RED_BLOCKS = 8 x 148 CTAs would be 2.67 waves)
        constexpr int WAVE3 = 3 * 148;
        if (nd == 0) {
            const RedCtx R0 = stopctx();
            if (light && F.n_upd == 1) LAUNCH((k_light<1, 0, 4>), WAVE3, TPB, dev, R0, n, (const unsigned char*)nullptr, ctx->n_var, RED_BLOCKS);
            else if (light) LAUNCH((k_light<2, 0, 2>), WAVE3, TPB, dev, R0, n, (const unsigned char*)nullptr, ctx->n_var, RED_BLOCKS);
            else LAUNCH((k_fused<0>), RED_BLOCKS, TPB, dev, R0, n, (const unsigned char*)nullptr, ctx->n_var);
            return MFB_OK;
        }
        ProfScope ps(ctx, MFB_T_REDUCE);
        const RedCtx R = redctx();
        if (light) {
            if (F.n_upd == 1) {
                if (nd == 1) LAUNCH((k_light<1, 1, 2>), WAVE3, TPB, dev, R, n, mask(), ctx->n_var, RED_BLOCKS);
                else if (nd == 2) LAUNCH((k_light<1, 2, 2>), WAVE3, TPB, dev, R, n, mask(), ctx->n_var, RED_BLOCKS);
                else LAUNCH((k_light<1, 4, 1>), WAVE3, TPB, dev, R, n, mask(), ctx->n_var, RED_BLOCKS);
            } else {
                if (nd <= 2) LAUNCH((k_light<2, 2, 1>), WAVE3, TPB, dev, R, n, mask(), ctx->n_var, RED_BLOCKS);
                else LAUNCH((k_light<2, 4, 1>), WAVE3, TPB, dev, R, n, mask(), ctx->n_var, RED_BLOCKS);
            }
            return after_reduce(R, nd, F.ops[0], F.ops[1]);
        }
        if (F.n_upd == 0 && nd <= 2) {                     // one or two plain dots: more loads in flight per thread
            if (nd == 1) LAUNCH((k_dots<1, 4>), WAVE3, TPB, dev, R, n, mask(), ctx->n_var, RED_BLOCKS);
            else LAUNCH((k_dots<2, 2>), WAVE3, TPB, dev, R, n, mask(), ctx->n_var, RED_BLOCKS);
            return after_reduce(R, nd, F.ops[0], F.ops[1]);
        }
        if (nd <= 2) LAUNCH((k_fused<2>), RED_BLOCKS, TPB, dev, R, n, mask(), ctx->n_var);
        else if (nd <= 4) LAUNCH((k_fused<4>), RED_BLOCKS, TPB, dev, R, n, mask(), ctx->n_var);
        else if (nd <= 8) LAUNCH((k_fused<8>), RED_BLOCKS, TPB, dev, R, n, mask(), ctx->n_var);
        else LAUNCH((k_fused<FU_MAXD>), RED_BLOCKS, TPB, dev, R, n, mask(), ctx->n_var);
        return after_reduce(R, nd, F.ops[0], F.ops[1]);
    }
    // y = A x with the dot w'y (+ scalar op) fused into the SpMV tail when the product needs no completion by other ranks
    // and no left preconditioner; otherwise SpMV (+ interface exchange, + Pl) followed by the program `dot_prog`
    // (off by default: inlined into the stream loop's kernel the fold tail makes ptxas spill inside the loop under the 4-CTA
    // register cap -- 3.3 ms instead of 2.2 ms per SpMV measured; MFB_SPMV_DOT=1 re-enables it for experiments)
    bool can_fuse_spmv_dot() const {
        static const bool want = [] { const char* e = getenv("MFB_SPMV_DOT"); return e && e[0] == '1'; }();
        return want && mfb_spmv_kind(ctx) == 1 && ctx->n_var <= 5 && !has_pl() && !mfb_is_distributed(ctx);
    }
    int mul_dot(double* y, const double* x, const double* w, const ScOp& op, int dot_prog) {
        spmv++;
        if (w && can_fuse_spmv_dot()) {
            const RedCtx R = redctx();
            return spmv_launch(ctx, A, x, y, w, R, op);
        }
        MFB_TRY(spmv_launch(ctx, A, x, y, nullptr, stopctx(), ScOp{OP_NONE, 0, 0, 0}));
        MFB_TRY(mfb_halo_add(ctx, y, ctx->n_var));
        MFB_TRY(Pl(y));
        if (w) MFB_TRY(run(dot_prog));
        return MFB_OK;
    }
};

// builder of one fused program
struct FB {
    Fused F;
    FB() { memset(&F, 0, sizeof(F)); }
    FB& upd(double* y, double ay, const double* ayp = nullptr) {
        FusedUpd& U = F.u[F.n_upd++];
        U.y = y; U.ay = ay; U.ayp = ayp; U.t0 = F.n_term; U.nt = 0;
        return *this;
    }
    FB& term(double c, const double* cp, const double* x) {
        FusedTerm& T = F.t[F.n_term++];
        T.c = c; T.cp = cp; T.x = x;
        F.u[F.n_upd - 1].nt++;
        return *this;
    }
    FB& dot(const double* x, const double* y) {      // nullptr operand: the result of the last update
        F.dx[F.n_dot] = x; F.dy[F.n_dot] = y; F.n_dot++;
        return *this;
    }
    FB& op(int which, int code, int i, int j, int off) {
        F.ops[which] = ScOp{code, i, j, off};
        return *this;
    }
};

// r = Pl(b - A x) (the opening lines of every solver) or, with left == false, the plain b - A x of iterative_Solve!;
// returns the normalized norm
int true_residual(Solver& S, double* r, const double* b, const double* x, double* res, bool left = true) {
    const double* keep = S.pl;
    const bool keep_ilu = S.pl_ilu;
    S.pl = nullptr; S.pl_ilu = false;
    int rc = S.mul(r, x);
    S.pl = keep; S.pl_ilu = keep_ilu;
    MFB_TRY(rc);
    double c[1] = {1.0};
    const double* xs[1] = {b};
    double n2;
    if (left && S.has_pl()) {
        MFB_TRY(S.lincomb(r, -1.0, 1, c, xs));
        MFB_TRY(S.Pl(r));
        double d;
        MFB_TRY(S.dot1(r, r, &d));
        n2 = d;
    } else {
        MFB_TRY(S.lincomb(r, -1.0, 1, c, xs, &n2));
    }
    *res = S.nn(n2);
    return MFB_OK;
}

// idrs!  (04_IDRs.jl:26-95). Vectors: P[s], U[s], G[s], Ar. Same operations in the same order as the reference; M, f, c,
// omega, alpha, beta live on the device (one-thread kernels between the vector kernels): one host synchronisation per
// inner step -- the convergence test the reference makes there -- instead of one per dot product (k + 3 of them).
int idrs_legacy(Solver& S, double* x, const double* b, double* r, double tol, int maxiter, int s, uint64_t seed, int pass,
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
    MFB_CUDA(ctx->ksc.alloc(SC_COUNT > 8 + 3 * MAXD + MAXD * MAXD ? SC_COUNT : 8 + 3 * MAXD + MAXD * MAXD));
    double* sc = ctx->ksc.p;
    const double* red = ctx->scal.p;
    const double* cdev = sc + ID_F + s;          // c[i]
    const double* ocdev = sc + ID_F + 2 * s;     // -omega * c[i]
    LAUNCH(k_idr_init, 1, 128, sc, s);
    std::vector<const double*> xs(2 * s + 2), ys(s + 2), cp(2 * s + 2);
    std::vector<double> cf(2 * s + 2);
    auto check = [&](double* resout) -> int {    // ||r||^2 is in ctx->scal[0]
        MFB_CUDA(cudaMemcpyAsync(ctx->h_scal, ctx->scal.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        *resout = S.nn(ctx->h_scal[0]);
        return MFB_OK;
    };
    while (true) {
        for (int i = 0; i < s; ++i) { xs[i] = P[i]; ys[i] = r; }
        MFB_TRY(S.dots_dev(s, xs.data(), ys.data()));
        LAUNCH(k_idr_f, 1, 32, sc, red, s);
        for (int k = 0; k < s; ++k) {
            LAUNCH(k_idr_c, 1, 1, sc, k, s);
            // U[k] = c_k U[k] + sum_{i>k} c_i U[i] + omega*(r - sum_{i>=k} c_i G[i])
            int nt = 0;
            for (int i = k; i < s; ++i) {
                if (i != k) { cf[nt] = 1.0; cp[nt] = cdev + i; xs[nt++] = U[i]; }
                cf[nt] = 1.0; cp[nt] = ocdev + i; xs[nt++] = G[i];
            }
            cf[nt] = 1.0; cp[nt] = sc + ID_OMEGA; xs[nt++] = r;
            MFB_TRY(S.lincomb_dev(U[k], 1.0, nt, cf.data(), cp.data(), xs.data(), false, cdev + k));
            MFB_TRY(S.mul(G[k], U[k]));
            for (int i = 0; i < k; ++i) {
                const double* dx1[1] = {P[i]};
                const double* dy1[1] = {G[k]};
                MFB_TRY(S.dots_dev(1, dx1, dy1));
                LAUNCH(k_idr_alpha, 1, 1, sc, red, i, s);
                double* yy[2] = {G[k], U[k]};
                const double a[2] = {1.0, 1.0}, bb[2] = {-1.0, -1.0};
                const double* ap[2] = {nullptr, nullptr};
                const double* bp[2] = {sc + ID_ALPHA, sc + ID_ALPHA};
                const double* xx[2] = {G[i], U[i]};
                MFB_TRY(S.axpby_batch_dev(2, yy, a, ap, bb, bp, xx));
            }
            for (int i = k; i < s; ++i) { xs[i - k] = P[i]; ys[i - k] = G[k]; }
            MFB_TRY(S.dots_dev(s - k, xs.data(), ys.data()));
            LAUNCH(k_idr_M, 1, 1, sc, red, k, s);
            {
                const double one = 1.0, mone = -1.0;
                const double* bptr[1] = {sc + ID_BETA};
                const double* ux[1] = {U[k]};
                const double* gx[1] = {G[k]};
                MFB_TRY(S.lincomb_dev(x, 1.0, 1, &one, bptr, ux));
                MFB_TRY(S.lincomb_dev(r, 1.0, 1, &mone, bptr, gx, true));
                MFB_TRY(check(&res));
            }
            if (!(res > tol) || iter >= maxiter) { *iters = iter; return MFB_OK; }
            iter++;
        }
        MFB_TRY(S.mul(Ar, r));
        {
            const double* a3[3] = {Ar, r, Ar};
            const double* b3[3] = {Ar, r, r};
            MFB_TRY(S.dots_dev(3, a3, b3));
            LAUNCH(k_idr_omega, 1, 1, sc, red);
            const double one = 1.0, mone = -1.0;
            const double* optr[1] = {sc + ID_OMEGA};
            const double* rx[1] = {r};
            const double* ax[1] = {Ar};
            MFB_TRY(S.lincomb_dev(x, 1.0, 1, &one, optr, rx));
            MFB_TRY(S.lincomb_dev(r, 1.0, 1, &mone, optr, ax, true));
            MFB_TRY(check(&res));
        }
        if (!(res > tol) || iter >= maxiter) { *iters = iter; return MFB_OK; }
        iter++;
    }
}

// bicgstabl_GS!  (03_BiCGstabl.jl:18-96). Vectors: R[1..s] (R[0] = r), U[0..s], r_shadow.
// Same operations in the same order as the reference; the scalars (rho, alpha, beta, tau, sigma, gamma...) live on the
// device and are updated by one-thread kernels between the vector kernels, so the stream never drains inside an outer
// iteration: one host synchronisation per 2 s SpMVs (the convergence test) instead of one per dot product.
int bicgstabl_gs_legacy(Solver& S, double* x, const double* b, double* r, double tol, int maxiter, int s, uint64_t seed,
                        int pass, std::vector<double*>& W, int* iters) {
    mfb_ctx* ctx = S.ctx;
    const int64_t n = S.n;
    double res;
    MFB_TRY(true_residual(S, r, b, x, &res));
    if (res <= tol) { *iters = 0; return MFB_OK; }
    int iter = 1;
    if (s > 16) { ctx->err = "bicgstabl_GS: s > 16"; return MFB_ERR_ARG; }
    std::vector<double*> R(s + 1), U(s + 1);
    R[0] = r;
    for (int i = 1; i <= s; ++i) R[i] = W[i - 1];
    for (int i = 0; i <= s; ++i) U[i] = W[s + i];
    double* r_shadow = W[2 * s + 1];
    LAUNCH(k_rand, RED_BLOCKS, TPB, r_shadow, n, (unsigned long long)seed, (unsigned long long)(pass * 64 + 63), ctx->gid.p, ctx->n_var);
    for (int i = 1; i <= s; ++i) MFB_CUDA(cudaMemsetAsync(R[i], 0, n * sizeof(double), ctx->stream));
    for (int i = 0; i <= s; ++i) MFB_CUDA(cudaMemsetAsync(U[i], 0, n * sizeof(double), ctx->stream));
    MFB_CUDA(ctx->ksc.alloc(SC_COUNT));
    double* sc = ctx->ksc.p;
    const double* red = ctx->scal.p;
    LAUNCH(k_sc_init, 1, 128, sc);
    std::vector<double*> yy(s + 2);
    std::vector<const double*> xx(2 * s + 2), pa(2 * s + 2), pb(2 * s + 2);
    std::vector<double> ca(2 * s + 2), cb(2 * s + 2);
    while (true) {
        LAUNCH(k_sc_outer_begin, 1, 1, sc);
        for (int j = 0; j < s; ++j) {
            const double* d1x[1] = {r_shadow};
            const double* d1y[1] = {R[j]};
            MFB_TRY(S.dots_dev(1, d1x, d1y));
            LAUNCH(k_sc_beta, 1, 1, sc, red);
            for (int i = 0; i <= j; ++i) { yy[i] = U[i]; ca[i] = -1.0; pa[i] = sc + SC_BETA; cb[i] = 1.0; pb[i] = nullptr; xx[i] = R[i]; }
            MFB_TRY(S.axpby_batch_dev(j + 1, yy.data(), ca.data(), pa.data(), cb.data(), pb.data(), xx.data()));   // U[i] = R[i] - beta U[i]
            MFB_TRY(S.mul(U[j + 1], U[j]));
            d1y[0] = U[j + 1];
            MFB_TRY(S.dots_dev(1, d1x, d1y));
            LAUNCH(k_sc_alpha, 1, 1, sc, red);
            for (int i = 0; i <= j; ++i) { yy[i] = R[i]; ca[i] = 1.0; pa[i] = nullptr; cb[i] = -1.0; pb[i] = sc + SC_ALPHA; xx[i] = U[i + 1]; }
            yy[j + 1] = x; ca[j + 1] = 1.0; pa[j + 1] = nullptr; cb[j + 1] = 1.0; pb[j + 1] = sc + SC_ALPHA; xx[j + 1] = U[0];
            MFB_TRY(S.axpby_batch_dev(j + 2, yy.data(), ca.data(), pa.data(), cb.data(), pb.data(), xx.data()));   // R[i] -= alpha U[i+1]; x += alpha U[0]
            MFB_TRY(S.mul(R[j + 1], R[j]));
        }
        for (int j = 0; j < s; ++j) {
            for (int i = 0; i < j; ++i) {
                const double* ax[1] = {R[i + 1]};
                const double* ay[1] = {R[j + 1]};
                MFB_TRY(S.dots_dev(1, ax, ay));
                LAUNCH(k_sc_tau, 1, 1, sc, red, i, j, s);
                const double m1 = -1.0;
                const double* cp1[1] = {sc + SC_TAU + i * s + j};
                MFB_TRY(S.lincomb_dev(R[j + 1], 1.0, 1, &m1, cp1, ax));                                          // R[j+1] -= tau_ij R[i+1]
            }
            const double* a2[2] = {R[j + 1], R[0]};
            const double* b2[2] = {R[j + 1], R[j + 1]};
            MFB_TRY(S.dots_dev(2, a2, b2));
            LAUNCH(k_sc_sig, 1, 1, sc, red, j);
        }
        LAUNCH(k_sc_gamma, 1, 1, sc, s);
        // x += gam[0]*R[0] + sum gampp[j]*R[j+1]      (uses R[0] before its update, as the reference does)
        int nt = 0;
        ca[nt] = 1.0; pa[nt] = sc + SC_GAM; xx[nt++] = R[0];
        for (int j = 0; j < s - 1; ++j) { ca[nt] = 1.0; pa[nt] = sc + SC_GAMPP + j; xx[nt++] = R[j + 1]; }
        MFB_TRY(S.lincomb_dev(x, 1.0, nt, ca.data(), pa.data(), xx.data()));
        // U[0] -= gam[s-1]*U[s] + sum gam[j]*U[j+1]
        nt = 0;
        ca[nt] = -1.0; pa[nt] = sc + SC_GAM + s - 1; xx[nt++] = U[s];
        for (int j = 0; j < s - 1; ++j) { ca[nt] = -1.0; pa[nt] = sc + SC_GAM + j; xx[nt++] = U[j + 1]; }
        MFB_TRY(S.lincomb_dev(U[0], 1.0, nt, ca.data(), pa.data(), xx.data()));
        // R[0] -= gamp[s-1]*R[s] + sum gamp[j]*R[j+1]   (+ fused norm)
        nt = 0;
        ca[nt] = -1.0; pa[nt] = sc + SC_GAMP + s - 1; xx[nt++] = R[s];
        for (int j = 0; j < s - 1; ++j) { ca[nt] = -1.0; pa[nt] = sc + SC_GAMP + j; xx[nt++] = R[j + 1]; }
        MFB_TRY(S.lincomb_dev(R[0], 1.0, nt, ca.data(), pa.data(), xx.data(), true));
        MFB_CUDA(cudaMemcpyAsync(ctx->h_scal, ctx->scal.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));          // the one synchronisation of the outer iteration
        const double n2 = ctx->h_scal[0];
        iter += s;
        if (!(S.nn(n2) > tol) || iter >= maxiter) { *iters = iter; return MFB_OK; }   // also leaves on NaN (breakdown)
    }
}

// bicgstabl_GS!  (03_BiCGstabl.jl:18-96), fused form. Same operations in the same order as the reference, but
//  * every dot product rides in the kernel that produces its operand: r_shadow'(A u) in the SpMV tail, the Gram-Schmidt
//    coefficients of the MR part in the update that precedes them, ||r||^2 and the next rho in the closing update;
//  * the scalar recurrences run in the tail of those kernels (reduce_tail / sc_apply): no one-thread kernels;
//  * the convergence test sets a device-side stop flag that every later launch honours, and the host reads it back one outer
//    iteration late: the stream never drains (round 1: one synchronisation per outer iteration, 82 launches per 2 s SpMVs;
//    now 3 s + 12 for s = 4, i.e. 3 per SpMV on one GPU).
int bicgstabl_gs(Solver& S, double* x, const double* b, double* r, double tol, int maxiter, int s, uint64_t seed,
                 int pass, std::vector<double*>& W, int* iters) {
    mfb_ctx* ctx = S.ctx;
    const int64_t n = S.n;
    double res;
    MFB_TRY(true_residual(S, r, b, x, &res));
    if (res <= tol) { *iters = 0; return MFB_OK; }
    if (s > 16) { ctx->err = "bicgstabl_GS: s > 16"; return MFB_ERR_ARG; }
    std::vector<double*> R(s + 1), U(s + 1);
    R[0] = r;
    for (int i = 1; i <= s; ++i) R[i] = W[i - 1];
    for (int i = 0; i <= s; ++i) U[i] = W[s + i];
    double* r_shadow = W[2 * s + 1];
    LAUNCH(k_rand, RED_BLOCKS, TPB, r_shadow, n, (unsigned long long)seed, (unsigned long long)(pass * 64 + 63), ctx->gid.p, ctx->n_var);
    for (int i = 1; i <= s; ++i) MFB_CUDA(cudaMemsetAsync(R[i], 0, n * sizeof(double), ctx->stream));
    for (int i = 0; i <= s; ++i) MFB_CUDA(cudaMemsetAsync(U[i], 0, n * sizeof(double), ctx->stream));
    MFB_CUDA(ctx->ksc.alloc(SC_COUNT));
    double* sc = ctx->ksc.p;
    LAUNCH(k_sc_init, 1, 128, sc);
    S.f_tol = tol; S.f_maxiter = maxiter; S.f_S = s;
    const bool fuse = S.can_fuse_spmv_dot();
    // ---- programs of one outer iteration ----
    S.progs.clear();
    auto add = [&](const FB& fb) { S.progs.push_back(fb.F); return (int)S.progs.size() - 1; };
    const int p_rho0 = add(FB().dot(r_shadow, R[0]).op(0, OP_BETA0, 0, 0, 0));          // opening of the first outer iteration
    std::vector<int> pa(s), pd(s), pdotU(s, -1), pdotR(s, -1);
    for (int j = 0; j < s; ++j) {
        FB a;                                                                             // U[i] = R[i] - beta U[i], i <= j  (:49-51)
        for (int i = 0; i <= j; ++i) a.upd(U[i], -1.0, sc + SC_BETA).term(1.0, nullptr, R[i]);
        pa[j] = add(a);
        FB d;                                                                             // R[i] -= alpha U[i+1]; x += alpha U[0]  (:55-58)
        for (int i = 0; i <= j; ++i) d.upd(R[i], 1.0).term(-1.0, sc + SC_ALPHA, U[i + 1]);
        d.upd(x, 1.0).term(1.0, sc + SC_ALPHA, U[0]);
        pd[j] = add(d);
        if (!fuse) {
            pdotU[j] = add(FB().dot(r_shadow, U[j + 1]).op(0, OP_ALPHA, 0, 0, 0));
            if (j < s - 1) pdotR[j] = add(FB().dot(r_shadow, R[j + 1]).op(0, OP_BETA, 0, 0, 0));
        }
    }
    // MR part by modified Gram-Schmidt (:62-71): row j needs tau_ij = <R[i+1], R[j+1]> / sig_i one after the other
    std::vector<int> pmr;
    {
        FB m0;                                                                            // j = 0: sig_0, gamp_0 (+ tau_01 for the next row)
        m0.dot(R[1], R[1]).dot(R[0], R[1]).op(0, OP_SIG, 0, 0, 0);
        if (s > 1) m0.dot(R[1], R[2]).op(1, OP_TAU, 0, 1, 2);
        pmr.push_back(add(m0));
        for (int j = 1; j < s; ++j)
            for (int i = 0; i < j; ++i) {
                FB m;
                m.upd(R[j + 1], 1.0).term(-1.0, sc + SC_TAU + i * s + j, R[i + 1]);       // R[j+1] -= tau_ij R[i+1]
                if (i < j - 1) {
                    m.dot(R[i + 2], nullptr).op(0, OP_TAU, i + 1, j, 0);
                } else {
                    m.dot(nullptr, nullptr).dot(R[0], nullptr).op(0, OP_SIG, 0, j, 0);    // sig_j, gamp_j (and the gammas after the last row)
                    if (j < s - 1) m.dot(R[1], R[j + 2]).op(1, OP_TAU, 0, j + 1, 2);
                }
                pmr.push_back(add(m));
            }
    }
    FB fin;                                                                               // (:83-91), then ||r||^2 and the next rho
    fin.upd(x, 1.0).term(1.0, sc + SC_GAM, R[0]);
    for (int j = 0; j < s - 1; ++j) fin.term(1.0, sc + SC_GAMPP + j, R[j + 1]);
    fin.upd(U[0], 1.0).term(-1.0, sc + SC_GAM + s - 1, U[s]);
    for (int j = 0; j < s - 1; ++j) fin.term(-1.0, sc + SC_GAM + j, U[j + 1]);
    fin.upd(R[0], 1.0).term(-1.0, sc + SC_GAMP + s - 1, R[s]);
    for (int j = 0; j < s - 1; ++j) fin.term(-1.0, sc + SC_GAMP + j, R[j + 1]);
    fin.dot(nullptr, nullptr).dot(r_shadow, nullptr).op(0, OP_FINAL, 0, 0, 0);
    const int p_fin = add(fin);
    MFB_TRY(S.upload_programs());
    // ---- run: the host enqueues outer iteration k + 1 before it knows the outcome of k ----
    for (int q = 0; q < 2; ++q)
        if (!ctx->lag_ev[q]) MFB_CUDA(cudaEventCreateWithFlags(&ctx->lag_ev[q], cudaEventDisableTiming));
    double* hslot = ctx->h_scal + 32;                 // two read-back slots of 8 doubles: sc[0..8)
    MFB_TRY(S.run(p_rho0));
    const ScOp none{OP_NONE, 0, 0, 0};
    // the launches of an outer iteration that was enqueued after the stop flag went up do nothing: their profile events and
    // SpMV counts are dropped again (marks of the last two iterations)
    struct Mark { size_t ev[MFB_T_COUNT][2]; int spmv; int64_t launches; } marks[2];
    auto take_mark = [&](Mark& m) {
        for (int q = 0; q < MFB_T_COUNT; ++q) { m.ev[q][0] = ctx->prof[q].start.size(); m.ev[q][1] = ctx->prof[q].stop.size(); }
        m.spmv = S.spmv; m.launches = ctx->launches;
    };
    auto drop_after = [&](const Mark& m) {
        for (int q = 0; q < MFB_T_COUNT; ++q) {
            ProfEvents& P = ctx->prof[q];
            if (q == MFB_T_SOLVE) continue;                         // the enclosing solve scope is still open
            while (P.start.size() > m.ev[q][0]) { ctx->event_pool.push_back(P.start.back()); P.start.pop_back(); }
            while (P.stop.size() > m.ev[q][1]) { ctx->event_pool.push_back(P.stop.back()); P.stop.pop_back(); }
        }
        S.spmv = m.spmv; ctx->launches = m.launches;
    };
    for (int k = 0;; ++k) {
        take_mark(marks[k & 1]);
        for (int j = 0; j < s; ++j) {
            MFB_TRY(S.run(pa[j]));
            MFB_TRY(S.mul_dot(U[j + 1], U[j], r_shadow, ScOp{OP_ALPHA, 0, 0, 0}, pdotU[j]));
            MFB_TRY(S.run(pd[j]));
            if (j < s - 1) MFB_TRY(S.mul_dot(R[j + 1], R[j], r_shadow, ScOp{OP_BETA, 0, 0, 0}, pdotR[j]));
            else MFB_TRY(S.mul_dot(R[j + 1], R[j], nullptr, none, -1));
        }
        for (int pm : pmr) MFB_TRY(S.run(pm));
        MFB_TRY(S.run(p_fin));
        MFB_CUDA(cudaMemcpyAsync(hslot + 8 * (k & 1), sc, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaEventRecord(ctx->lag_ev[k & 1], ctx->stream));
        if (k >= 1) {
            MFB_CUDA(cudaEventSynchronize(ctx->lag_ev[(k - 1) & 1]));
            if (hslot[8 * ((k - 1) & 1) + SC_STOP] != 0.0) {       // iteration k - 1 was the last: iteration k ran empty
                MFB_CUDA(cudaStreamSynchronize(ctx->stream));       // its events must have completed before they are recycled
                drop_after(marks[k & 1]);
                break;
            }
        }
    }
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    // the launches enqueued after the stop flag went up did nothing: the newest slot holds the final scalars
    const double* fs = hslot[8 + SC_ITER] >= hslot[SC_ITER] ? hslot + 8 : hslot;
    *iters = (int)fs[SC_ITER];
    return MFB_OK;
}

// idrs!  (04_IDRs.jl:26-95), fused form: same operations in the same order; every reduction rides in the kernel that produces its
// operand, the small recurrences (f, c, M, alpha, beta, omega, the convergence test) run in the tails of those kernels, and the
// host reads the stop flag back once per cycle of s + 1 inner steps, one cycle late. Launches per cycle: 4 s + s (s - 1) / 2 + 3
// (63 for s = 8, i.e. 7 per SpMV; round 1: ~21 per SpMV and one host synchronisation per inner step).
int idrs(Solver& S, double* x, const double* b, double* r, double tol, int maxiter, int s, uint64_t seed, int pass,
         std::vector<double*>& W, int* iters) {
    mfb_ctx* ctx = S.ctx;
    const int64_t n = S.n;
    double res;
    MFB_TRY(true_residual(S, r, b, x, &res));
    if (res <= tol) { *iters = 0; return MFB_OK; }
    double** P = &W[0];
    double** U = &W[s];
    double** G = &W[2 * s];
    double* Ar = W[3 * s];
    for (int k = 0; k < s; ++k) {
        LAUNCH(k_rand, RED_BLOCKS, TPB, P[k], n, (unsigned long long)seed, (unsigned long long)(pass * 64 + k), ctx->gid.p, ctx->n_var);
        MFB_CUDA(cudaMemsetAsync(U[k], 0, n * sizeof(double), ctx->stream));
        MFB_CUDA(cudaMemsetAsync(G[k], 0, n * sizeof(double), ctx->stream));
    }
    MFB_CUDA(ctx->ksc.alloc(SC_COUNT > 8 + 3 * MAXD + MAXD * MAXD ? SC_COUNT : 8 + 3 * MAXD + MAXD * MAXD));
    double* sc = ctx->ksc.p;
    LAUNCH(k_idr_init, 1, 128, sc, s);                      // omega = 1, M = I, STOP = 0, ITER = 1
    S.f_tol = tol; S.f_maxiter = maxiter; S.f_S = s;
    const double* cdev = sc + IDS_F + s;                    // c[i]
    const double* ocdev = sc + IDS_F + 2 * s;               // -omega c[i]
    S.progs.clear();
    auto add = [&](const FB& fb) { S.progs.push_back(fb.F); return (int)S.progs.size() - 1; };
    FB f0;                                                  // f = P' r, c of k = 0 (opening of the first cycle)
    for (int i = 0; i < s; ++i) f0.dot(P[i], r);
    f0.op(0, OP_IDR_FC, 1, 0, 0);
    const int p_f0 = add(f0);
    std::vector<int> pB(s), pD(s), pE(s);
    std::vector<std::vector<int>> pC(s);
    for (int k = 0; k < s; ++k) {
        FB B;                                               // U[k] = c_k U[k] + sum_{i>k} c_i U[i] + omega (r - sum_{i>=k} c_i G[i])  (:53-63)
        B.upd(U[k], 1.0, cdev + k);
        for (int i = k; i < s; ++i) {
            if (i != k) B.term(1.0, cdev + i, U[i]);
            B.term(1.0, ocdev + i, G[i]);
        }
        B.term(1.0, sc + IDS_OMEGA, r);
        pB[k] = add(B);
        FB D;                                               // after G[k] = A U[k]: alpha_0 (k > 0) or the whole column M[., 0]
        if (k == 0) { for (int i = 0; i < s; ++i) D.dot(P[i], G[0]); D.op(0, OP_IDR_M, 0, 0, 0); }
        else D.dot(P[0], G[k]).op(0, OP_IDR_ALPHA, 0, 0, 0);
        pD[k] = add(D);
        for (int i = 0; i < k; ++i) {                       // G[k] -= alpha G[i]; U[k] -= alpha U[i]  (:65-69), next alpha or M[., k]
            FB Cp;
            Cp.upd(U[k], 1.0).term(-1.0, sc + IDS_ALPHA, U[i]);
            Cp.upd(G[k], 1.0).term(-1.0, sc + IDS_ALPHA, G[i]);          // last update: the dots below see the new G[k]
            if (i + 1 < k) Cp.dot(P[i + 1], nullptr).op(0, OP_IDR_ALPHA, i + 1, 0, 0);
            else { for (int q = k; q < s; ++q) Cp.dot(P[q], nullptr); Cp.op(0, OP_IDR_M, 0, k, 0); }
            pC[k].push_back(add(Cp));
        }
        FB E;                                               // x += beta U[k]; r -= beta G[k]; ||r||^2  (:77-80)
        E.upd(x, 1.0).term(1.0, sc + IDS_BETA, U[k]);
        E.upd(r, 1.0).term(-1.0, sc + IDS_BETA, G[k]);
        E.dot(nullptr, nullptr).op(0, OP_IDR_RES, 0, k + 1 < s ? k + 1 : -1, 0);
        pE[k] = add(E);
    }
    FB Fo;                                                  // omega from Ar = A r  (:87-88)
    Fo.dot(Ar, Ar).dot(r, r).dot(Ar, r).op(0, OP_IDR_OMEGA, 0, 0, 0);
    const int p_om = add(Fo);
    FB Gp;                                                  // x += omega r; r -= omega Ar; ||r||^2; f = P' r and c of the next cycle  (:89-93)
    Gp.upd(x, 1.0).term(1.0, sc + IDS_OMEGA, r);
    Gp.upd(r, 1.0).term(-1.0, sc + IDS_OMEGA, Ar);
    Gp.dot(nullptr, nullptr);
    for (int i = 0; i < s; ++i) Gp.dot(P[i], nullptr);
    Gp.op(0, OP_IDR_RES, 0, -1, 0).op(1, OP_IDR_FC, 1, 0, 1);
    const int p_end = add(Gp);
    MFB_TRY(S.upload_programs());
    for (int q = 0; q < 2; ++q)
        if (!ctx->lag_ev[q]) MFB_CUDA(cudaEventCreateWithFlags(&ctx->lag_ev[q], cudaEventDisableTiming));
    double* hslot = ctx->h_scal + 32;
    MFB_TRY(S.run(p_f0));
    struct Mark { size_t ev[MFB_T_COUNT][2]; int spmv; int64_t launches; } marks[2];
    auto take_mark = [&](Mark& m) {
        for (int q = 0; q < MFB_T_COUNT; ++q) { m.ev[q][0] = ctx->prof[q].start.size(); m.ev[q][1] = ctx->prof[q].stop.size(); }
        m.spmv = S.spmv; m.launches = ctx->launches;
    };
    auto drop_after = [&](const Mark& m) {
        for (int q = 0; q < MFB_T_COUNT; ++q) {
            if (q == MFB_T_SOLVE) continue;
            ProfEvents& Pe = ctx->prof[q];
            while (Pe.start.size() > m.ev[q][0]) { ctx->event_pool.push_back(Pe.start.back()); Pe.start.pop_back(); }
            while (Pe.stop.size() > m.ev[q][1]) { ctx->event_pool.push_back(Pe.stop.back()); Pe.stop.pop_back(); }
        }
        S.spmv = m.spmv; ctx->launches = m.launches;
    };
    const ScOp none{OP_NONE, 0, 0, 0};
    for (int cyc = 0;; ++cyc) {
        take_mark(marks[cyc & 1]);
        for (int k = 0; k < s; ++k) {
            MFB_TRY(S.run(pB[k]));
            MFB_TRY(S.mul_dot(G[k], U[k], nullptr, none, -1));
            MFB_TRY(S.run(pD[k]));
            for (int pc : pC[k]) MFB_TRY(S.run(pc));
            MFB_TRY(S.run(pE[k]));
        }
        MFB_TRY(S.mul_dot(Ar, r, nullptr, none, -1));
        MFB_TRY(S.run(p_om));
        MFB_TRY(S.run(p_end));
        MFB_CUDA(cudaMemcpyAsync(hslot + 8 * (cyc & 1), sc, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaEventRecord(ctx->lag_ev[cyc & 1], ctx->stream));
        if (cyc >= 1) {
            MFB_CUDA(cudaEventSynchronize(ctx->lag_ev[(cyc - 1) & 1]));
            if (hslot[8 * ((cyc - 1) & 1) + SC_STOP] != 0.0) {
                MFB_CUDA(cudaStreamSynchronize(ctx->stream));
                drop_after(marks[cyc & 1]);
                break;
            }
        }
    }
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    const double* fs = hslot[8 + SC_ITER] >= hslot[SC_ITER] ? hslot + 8 : hslot;
    *iters = (int)fs[SC_ITER];
    return MFB_OK;
}

int ensure_scalars(mfb_ctx* ctx) {
    // partial sums: MAXD dots x RED_BLOCKS blocks for the vector kernels, one per CTA for the SpMV with a fused dot
    const size_t need = 64 + std::max((size_t)MAXD * RED_BLOCKS, (size_t)(ctx->N / 32 + 2));
    if (!ctx->scal.p || ctx->scal.n < need) MFB_CUDA(ctx->scal.alloc(need));
    if (!ctx->h_scal) MFB_CUDA(cudaMallocHost((void**)&ctx->h_scal, 64 * sizeof(double)));
    if (!ctx->red_counter.p) {
        MFB_CUDA(ctx->red_counter.alloc(4));
        MFB_CUDA(cudaMemsetAsync(ctx->red_counter.p, 0, 4 * sizeof(unsigned), ctx->stream));
    }
    return MFB_OK;
}

int ensure_work(mfb_ctx* ctx, int count, int64_t n) {
    if ((int)ctx->work.size() < count) ctx->work.resize(count);
    for (int i = 0; i < count; ++i) MFB_CUDA(ctx->work[i].alloc(n));
    return MFB_OK;
}


// small helpers for the remaining methods -----------------------------------------------------------------------------
struct Vec {
    Solver& S;
    int set(double* y, std::initializer_list<double> c, std::initializer_list<const double*> x, double* norm2 = nullptr) {
        return S.lincomb(y, 0.0, (int)c.size(), c.begin(), x.begin(), norm2);                 // y = sum c_i x_i
    }
    int add(double* y, std::initializer_list<double> c, std::initializer_list<const double*> x, double* norm2 = nullptr) {
        return S.lincomb(y, 1.0, (int)c.size(), c.begin(), x.begin(), norm2);                 // y += sum c_i x_i
    }
    int dot(const double* x, const double* y, double* out) { return S.dot1(x, y, out); }
};

// bicgstabl!  (03_BiCGstabl.jl:98-162): BiCG part as in bicgstabl_GS!, MR part by the normal equations M gamma = M[:,1]
// solved with a (host) LU, as the reference does with lu!.
int bicgstabl_lu(Solver& S, double* x, const double* b, double* r, double tol, int maxiter, int s, uint64_t seed, int pass,
                 std::vector<double*>& W, int* iters) {
    mfb_ctx* ctx = S.ctx;
    const int64_t n = S.n;
    Vec V{S};
    double res;
    MFB_TRY(true_residual(S, r, b, x, &res));
    // the reference tests the UN-normalised norm here (`r_norm <= tol`, :103-104)
    if (res * std::sqrt(S.n_global > 0 ? S.n_global : (double)n) <= tol) { *iters = 0; return MFB_OK; }
    int iter = 1;
    std::vector<double*> R(s + 1), U(s + 1);
    R[0] = r;
    for (int i = 1; i <= s; ++i) R[i] = W[i - 1];
    for (int i = 0; i <= s; ++i) U[i] = W[s + i];
    double* r_shadow = W[2 * s + 1];
    LAUNCH(k_rand, RED_BLOCKS, TPB, r_shadow, n, (unsigned long long)seed, (unsigned long long)(pass * 64 + 63), ctx->gid.p, ctx->n_var);
    for (int i = 1; i <= s; ++i) MFB_CUDA(cudaMemsetAsync(R[i], 0, n * sizeof(double), ctx->stream));
    for (int i = 0; i <= s; ++i) MFB_CUDA(cudaMemsetAsync(U[i], 0, n * sizeof(double), ctx->stream));
    std::vector<double> gam(s), M((s + 1) * (s + 1));
    double omega = 1.0, rho0 = 1.0, alpha = 0.0;
    std::vector<double*> yy(s + 1);
    std::vector<double> aa(s + 1), bb(s + 1);
    std::vector<const double*> xx(2 * s + 2), ys(2 * s + 2);
    std::vector<double> cf(2 * s + 2);
    while (true) {
        rho0 *= -omega;
        for (int j = 0; j < s; ++j) {
            double rho1;
            MFB_TRY(V.dot(r_shadow, R[j], &rho1));
            const double beta = alpha * rho1 / rho0;
            rho0 = rho1;
            for (int i = 0; i <= j; ++i) { yy[i] = U[i]; aa[i] = -beta; bb[i] = 1.0; xx[i] = R[i]; }
            MFB_TRY(S.axpby_batch(j + 1, yy.data(), aa.data(), bb.data(), xx.data()));      // U[i] = R[i] - beta U[i]
            MFB_TRY(S.mul(U[j + 1], U[j]));
            double d;
            MFB_TRY(V.dot(r_shadow, U[j + 1], &d));
            alpha = rho0 / d;
            for (int i = 0; i <= j; ++i) { yy[i] = R[i]; aa[i] = 1.0; bb[i] = -alpha; xx[i] = U[i + 1]; }
            MFB_TRY(S.axpby_batch(j + 1, yy.data(), aa.data(), bb.data(), xx.data()));      // R[i] -= alpha U[i+1]
            MFB_TRY(S.mul(R[j + 1], R[j]));
            MFB_TRY(V.add(x, {alpha}, {U[0]}));
        }
        // MR part: Gram matrix of R[0..s] in fused multi-dot launches
        int nd = 0;
        std::vector<const double*> da, db;
        for (int j = 0; j <= s; ++j)
            for (int i = 0; i <= j; ++i) { da.push_back(R[i]); db.push_back(R[j]); }
        std::vector<double> g(da.size());
        for (size_t o = 0; o < da.size(); o += MAXD) {
            const int k = (int)std::min<size_t>(MAXD, da.size() - o);
            MFB_TRY(S.dots(k, da.data() + o, db.data() + o));
            for (int i = 0; i < k; ++i) g[o + i] = ctx->h_scal[i];
        }
        for (int j = 0; j <= s; ++j)
            for (int i = 0; i <= j; ++i) { M[i * (s + 1) + j] = M[j * (s + 1) + i] = g[nd++]; }
        // gamma = M[L,L] \ M[L,1], L = 2..s+1: Gaussian elimination with partial pivoting (lu!)
        {
            std::vector<double> Am(s * s), rhs(s);
            for (int i = 0; i < s; ++i) {
                rhs[i] = M[(i + 1) * (s + 1)];
                for (int j = 0; j < s; ++j) Am[i * s + j] = M[(i + 1) * (s + 1) + j + 1];
            }
            for (int c = 0; c < s; ++c) {
                int pv = c;
                for (int i = c + 1; i < s; ++i) if (std::fabs(Am[i * s + c]) > std::fabs(Am[pv * s + c])) pv = i;
                if (pv != c) { for (int j = 0; j < s; ++j) std::swap(Am[c * s + j], Am[pv * s + j]); std::swap(rhs[c], rhs[pv]); }
                for (int i = c + 1; i < s; ++i) {
                    const double f = Am[i * s + c] / Am[c * s + c];
                    for (int j = c; j < s; ++j) Am[i * s + j] -= f * Am[c * s + j];
                    rhs[i] -= f * rhs[c];
                }
            }
            for (int i = s - 1; i >= 0; --i) {
                double v = rhs[i];
                for (int j = i + 1; j < s; ++j) v -= Am[i * s + j] * gam[j];
                gam[i] = v / Am[i * s + i];
            }
        }
        int nt = 0;
        for (int i = 0; i < s; ++i) { cf[nt] = -gam[i]; xx[nt++] = U[i + 1]; }
        MFB_TRY(S.lincomb(U[0], 1.0, nt, cf.data(), xx.data()));
        nt = 0;
        for (int i = 0; i < s; ++i) { cf[nt] = gam[i]; xx[nt++] = R[i]; }
        MFB_TRY(S.lincomb(x, 1.0, nt, cf.data(), xx.data()));
        nt = 0;
        for (int i = 0; i < s; ++i) { cf[nt] = -gam[i]; xx[nt++] = R[i + 1]; }
        double n2;
        MFB_TRY(S.lincomb(R[0], 1.0, nt, cf.data(), xx.data(), &n2));
        omega = gam[s - 1];
        iter += s;
        if (S.nn(n2) <= tol || iter >= maxiter) { *iters = iter; return MFB_OK; }
    }
}

// x = A \ b for a small dense system on the host: Gaussian elimination with partial pivoting (what Julia's `\` does for a square
// matrix: lu!). A is s x s row-major and is overwritten.
void dense_solve(std::vector<double>& Am, std::vector<double>& rhs, std::vector<double>& sol, int s) {
    for (int c = 0; c < s; ++c) {
        int pv = c;
        for (int i = c + 1; i < s; ++i) if (std::fabs(Am[i * s + c]) > std::fabs(Am[pv * s + c])) pv = i;
        if (pv != c) { for (int j = 0; j < s; ++j) std::swap(Am[c * s + j], Am[pv * s + j]); std::swap(rhs[c], rhs[pv]); }
        for (int i = c + 1; i < s; ++i) {
            const double f = Am[i * s + c] / Am[c * s + c];
            for (int j = c; j < s; ++j) Am[i * s + j] -= f * Am[c * s + j];
            rhs[i] -= f * rhs[c];
        }
    }
    for (int i = s - 1; i >= 0; --i) {
        double v = rhs[i];
        for (int j = i + 1; j < s; ++j) v -= Am[i * s + j] * sol[j];
        sol[i] = v / Am[i * s + i];
    }
}

// modify_Omega (04_IDRs.jl:1-8) on the host from the three reductions |v1|^2, |v2|^2, v1'v2
double modify_omega_host(double n1sq, double n2sq, double d) {
    const double angle = 0.70710678118654752440;
    const double n1 = std::sqrt(n1sq), n2 = std::sqrt(n2sq);
    const double rho = std::fabs(d / (n1 * n2));
    double omega = d / (n1 * n1);
    if (rho < angle) omega = omega * angle / rho;
    return omega;
}

// idrs_original!  (04_IDRs.jl:97-169; "not used, not exploiting orthogonality"). Restated operation by operation, INCLUDING the
// k == 0 branch that overwrites r with Pl(A Q) instead of subtracting it (:143-148): the method is exported, so it is provided
// with the reference's behaviour, not repaired. Vectors: P[s], U[s], G[s], Q, V, Ar. Scalars on the host (M \ f is a dense LU).
int idrs_original(Solver& S, double* x, const double* b, double* r, double tol, int maxiter, int s, uint64_t seed, int pass,
                  std::vector<double*>& W, int* iters) {
    mfb_ctx* ctx = S.ctx;
    const int64_t n = S.n;
    Vec V_{S};
    double res;
    MFB_TRY(true_residual(S, r, b, x, &res));
    if (res <= tol) { *iters = 0; return MFB_OK; }
    int iter = 1;
    double** P = &W[0];
    double** U = &W[s];
    double** G = &W[2 * s];
    double *Q = W[3 * s], *V = W[3 * s + 1], *Ar = W[3 * s + 2];
    for (int k = 0; k < s; ++k)
        LAUNCH(k_rand, RED_BLOCKS, TPB, P[k], n, (unsigned long long)seed, (unsigned long long)(pass * 64 + k), ctx->gid.p, ctx->n_var);
    std::vector<double> M(s * s, 0.0), f(s, 0.0), c(s, 0.0);
    std::vector<const double*> xs(s + 1), ys(s + 1);
    std::vector<double> cf(s + 1);
    double omega = 1.0, n2 = 0.0;
    auto fill_M_col = [&](int k) -> int {                   // M[i, k] = P[i]' G[k]
        for (int i = 0; i < s; ++i) { xs[i] = P[i]; ys[i] = G[k]; }
        MFB_TRY(S.dots(s, xs.data(), ys.data()));
        for (int i = 0; i < s; ++i) M[i * s + k] = ctx->h_scal[i];
        return MFB_OK;
    };
    for (int k = 0; k < s; ++k) {                           // (:113-124)
        MFB_CUDA(cudaMemcpyAsync(U[k], r, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        MFB_TRY(S.mul(G[k], r));
        const double* a3[3] = {G[k], r, G[k]};
        const double* b3[3] = {G[k], r, r};
        MFB_TRY(S.dots(3, a3, b3));
        omega = modify_omega_host(ctx->h_scal[0], ctx->h_scal[1], ctx->h_scal[2]);
        MFB_TRY(V_.add(x, {omega}, {U[k]}));
        MFB_TRY(V_.add(r, {-omega}, {G[k]}, &n2));
        MFB_TRY(fill_M_col(k));
    }
    while (true) {
        if (S.nn(n2) <= tol || iter >= maxiter) { *iters = iter; return MFB_OK; }      // (:128)
        iter++;
        for (int k = 0; k <= s; ++k) {
            for (int i = 0; i < s; ++i) { xs[i] = P[i]; ys[i] = r; }
            MFB_TRY(S.dots(s, xs.data(), ys.data()));
            for (int i = 0; i < s; ++i) f[i] = ctx->h_scal[i];
            {
                std::vector<double> Am(M), rhs(f);
                dense_solve(Am, rhs, c, s);                 // c = M \ f (:135)
            }
            // V = r - sum c_i G[i];  Q = sum c_i U[i]   (:137-144)
            for (int i = 0; i < s; ++i) { cf[i] = -c[i]; xs[i] = G[i]; }
            cf[s] = 1.0; xs[s] = r;
            MFB_TRY(S.lincomb(V, 0.0, s + 1, cf.data(), xs.data()));
            for (int i = 0; i < s; ++i) { cf[i] = c[i]; xs[i] = U[i]; }
            MFB_TRY(S.lincomb(Q, 0.0, s, cf.data(), xs.data()));
            if (k == 0) {
                MFB_TRY(S.mul(Ar, V));
                const double* a3[3] = {Ar, V, Ar};
                const double* b3[3] = {Ar, V, V};
                MFB_TRY(S.dots(3, a3, b3));
                omega = modify_omega_host(ctx->h_scal[0], ctx->h_scal[1], ctx->h_scal[2]);
                MFB_TRY(V_.add(Q, {omega}, {V}));
                MFB_TRY(V_.add(x, {1.0}, {Q}));
                MFB_TRY(S.mul(r, Q));                       // r = Pl(A Q): as the reference has it (:147-148)
                MFB_TRY(V_.dot(r, r, &n2));
            } else {
                MFB_TRY(V_.set(U[k - 1], {1.0, omega}, {Q, V}));
                MFB_TRY(S.mul(G[k - 1], U[k - 1]));
                MFB_TRY(V_.add(x, {1.0}, {U[k - 1]}));
                MFB_TRY(V_.add(r, {-1.0}, {G[k - 1]}, &n2));
                MFB_TRY(fill_M_col(k - 1));
            }
        }
    }
}

// Hessenberg least squares by Givens rotations (05_GMRES.jl:7-37): H is (w+1) x w column-major with leading dimension ld
void hessenberg_solve(std::vector<double>& H, int ld, int w, std::vector<double>& rhs) {
    for (int i = 0; i < w; ++i) {
        const double f = H[i * ld + i], g = H[i * ld + i + 1];
        double c, sn;
        if (g == 0.0) { c = 1.0; sn = 0.0; }
        else if (f == 0.0) { c = 0.0; sn = 1.0; }
        else { const double rr = std::copysign(std::hypot(f, g), f); c = f / rr; sn = g / rr; }
        H[i * ld + i] = c * f + sn * g;
        for (int j = i + 1; j < w; ++j) {
            const double tmp = -sn * H[j * ld + i] + c * H[j * ld + i + 1];
            H[j * ld + i] = c * H[j * ld + i] + sn * H[j * ld + i + 1];
            H[j * ld + i + 1] = tmp;
        }
        const double tmp = -sn * rhs[i] + c * rhs[i + 1];
        rhs[i] = c * rhs[i] + sn * rhs[i + 1];
        rhs[i + 1] = tmp;
    }
    for (int i = w - 1; i >= 0; --i) {
        double v = rhs[i];
        for (int j = i + 1; j < w; ++j) v -= H[j * ld + i] * rhs[j];
        rhs[i] = v / H[i * ld + i];
    }
}

// gmres!  (05_GMRES.jl:46-101): restarted every s iterations, modified Gram-Schmidt. Vectors: Q[0..s].
int gmres(Solver& S, double* x, const double* b, double* r, double tol, int maxiter, int s, std::vector<double*>& W, int* iters) {
    Vec V{S};
    double res;
    MFB_TRY(true_residual(S, r, b, x, &res));
    if (res <= tol) { *iters = 0; return MFB_OK; }
    int iter = 1;
    std::vector<double*> Q(s + 1);
    for (int i = 0; i <= s; ++i) Q[i] = W[i];
    const int ld = s + 1;
    std::vector<double> H((size_t)ld * s, 0.0), y(s + 1, 0.0);
    double d;
    MFB_TRY(V.dot(r, r, &d));
    double r_norm = std::sqrt(d);
    y[0] = r_norm;
    std::vector<double> cf(s + 1);
    std::vector<const double*> xs(s + 1);
    while (true) {
        MFB_TRY(V.set(Q[0], {1.0 / r_norm}, {r}));
        for (int i = 1; i <= s; ++i) {
            MFB_TRY(S.mul(Q[i], Q[i - 1]));
            for (int j = 0; j < i; ++j) {
                double h;
                MFB_TRY(V.dot(Q[j], Q[i], &h));
                H[(i - 1) * ld + j] = h;
                MFB_TRY(V.add(Q[i], {-h}, {Q[j]}));
            }
            double n2;
            MFB_TRY(V.dot(Q[i], Q[i], &n2));
            const double hn = std::sqrt(n2);
            H[(i - 1) * ld + i] = hn;
            if (hn == 0.0) {   // exact solve inside the cycle (:72-79)
                const int w = i - 1;
                if (w > 0) {
                    hessenberg_solve(H, ld, w, y);
                    for (int j = 0; j < w; ++j) { cf[j] = y[j]; xs[j] = Q[j]; }
                    MFB_TRY(S.lincomb(x, 1.0, w, cf.data(), xs.data()));
                }
                *iters = iter + i - 1;
                return MFB_OK;
            }
            MFB_TRY(V.set(Q[i], {1.0 / hn}, {Q[i]}));
        }
        hessenberg_solve(H, ld, s, y);
        for (int j = 0; j < s; ++j) { cf[j] = y[j]; xs[j] = Q[j]; }
        MFB_TRY(S.lincomb(x, 1.0, s, cf.data(), xs.data()));
        iter += s;
        MFB_TRY(true_residual(S, r, b, x, &res));
        if (res <= tol || iter > maxiter) { *iters = iter; return MFB_OK; }
        std::fill(y.begin(), y.end(), 0.0);
        std::fill(H.begin(), H.end(), 0.0);
        MFB_TRY(V.dot(r, r, &d));
        r_norm = std::sqrt(d);
        y[0] = r_norm;
    }
}

// cgs!  (07_CGS.jl:10-50). Vectors: r0, u, p, s, v. (The reference's FEM_buffer work vectors are uninitialised; zero here.)
int cgs(Solver& S, double* x, const double* b, double* r, double tol, int maxiter, std::vector<double*>& W, int* iters) {
    mfb_ctx* ctx = S.ctx;
    const int64_t n = S.n;
    Vec V{S};
    double res;
    MFB_TRY(true_residual(S, r, b, x, &res));
    if (res <= tol) { *iters = 0; return MFB_OK; }
    int iter = 1;
    double *r0 = W[0], *u = W[1], *p = W[2], *sv = W[3], *v = W[4];
    MFB_CUDA(cudaMemcpyAsync(r0, r, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    for (double* w : {u, p, sv, v}) MFB_CUDA(cudaMemsetAsync(w, 0, n * sizeof(double), ctx->stream));
    double rho = 1.0, rhobar, alpha, beta;
    while (true) {
        rhobar = rho;
        MFB_TRY(V.dot(r, r0, &rho));
        beta = rho / rhobar;
        MFB_TRY(V.set(sv, {1.0, beta}, {r, p}));                          // s = r + beta p
        MFB_TRY(S.lincomb(u, beta * beta, 2, std::initializer_list<double>{1.0, beta}.begin(),
                          std::initializer_list<const double*>{sv, p}.begin()));   // u = s + beta (p + beta u)
        MFB_TRY(S.mul(v, u));
        double d;
        MFB_TRY(V.dot(v, r0, &d));
        alpha = rho / d;
        MFB_TRY(V.set(p, {1.0, -alpha}, {sv, v}));                        // p = s - alpha v
        MFB_TRY(V.add(x, {alpha, alpha}, {p, sv}));                       // x += alpha (p + s)
        MFB_TRY(true_residual(S, r, b, x, &res));
        iter++;
        if (res <= tol || iter > maxiter) { *iters = iter; return MFB_OK; }
    }
}

// cgs2!  (07_CGS.jl:52-105). Vectors: r0, s0 (random), u, w, s, v, t, c.
int cgs2(Solver& S, double* x, const double* b, double* r, double tol, int maxiter, uint64_t seed, int pass,
         std::vector<double*>& W, int* iters) {
    mfb_ctx* ctx = S.ctx;
    const int64_t n = S.n;
    Vec V{S};
    double res;
    MFB_TRY(true_residual(S, r, b, x, &res));
    if (res <= tol) { *iters = 0; return MFB_OK; }
    int iter = 1;
    double *r0 = W[0], *s0 = W[1], *u = W[2], *w = W[3], *sv = W[4], *v = W[5], *t = W[6], *c = W[7];
    MFB_CUDA(cudaMemcpyAsync(r0, r, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    LAUNCH(k_rand, RED_BLOCKS, TPB, s0, n, (unsigned long long)seed, (unsigned long long)(pass * 64 + 62), ctx->gid.p, ctx->n_var);
    for (double* z : {u, w, sv, v, t, c}) MFB_CUDA(cudaMemsetAsync(z, 0, n * sizeof(double), ctx->stream));
    double alpha = 1, alphabar = 1, beta, betabar, rho, rhobar, sigma = 1, sigmabar = 1;
    while (true) {
        const double* da[2] = {r, r};
        const double* db[2] = {r0, s0};
        MFB_TRY(S.dots(2, da, db));
        rho = ctx->h_scal[0];
        rhobar = ctx->h_scal[1];
        beta = 1 / alphabar * rho / sigma;
        MFB_TRY(V.set(v, {1.0, beta}, {r, u}));                           // v = r + beta u
        betabar = 1 / alpha * rhobar / sigmabar;
        MFB_TRY(V.set(t, {1.0, betabar}, {r, sv}));                       // t = r + betabar s
        MFB_TRY(S.lincomb(w, beta * betabar, 2, std::initializer_list<double>{1.0, beta}.begin(),
                          std::initializer_list<const double*>{t, u}.begin()));    // w = t + beta (u + betabar w)
        MFB_TRY(S.mul(c, w));
        const double* ea[2] = {c, c};
        MFB_TRY(S.dots(2, ea, db));
        sigma = ctx->h_scal[0];
        sigmabar = ctx->h_scal[1];
        alpha = rho / sigma;
        alphabar = rhobar / sigmabar;
        MFB_TRY(V.set(sv, {1.0, -alpha}, {t, c}));                        // s = t - alpha c
        MFB_TRY(V.set(u, {1.0, -alphabar}, {v, c}));                      // u = v - alphabar c
        MFB_TRY(V.add(x, {alpha, alphabar}, {v, sv}));                    // x += alpha v + alphabar s
        MFB_TRY(true_residual(S, r, b, x, &res));
        iter++;
        if (res <= tol || iter > maxiter) { *iters = iter; return MFB_OK; }
    }
}

// tfqmr!  (08_QMR.jl:3-76). Vectors: r0, r_cgs, p, q, u, v, d, tmp.
int tfqmr(Solver& S, double* x, const double* b, double* r, double tol, int maxiter, int checkiter, std::vector<double*>& W,
          int* iters) {
    mfb_ctx* ctx = S.ctx;
    const int64_t n = S.n;
    Vec V{S};
    double res;
    MFB_TRY(true_residual(S, r, b, x, &res));
    if (res <= tol) { *iters = 0; return MFB_OK; }
    int iter = 1;
    double *r0 = W[0], *rc = W[1], *p = W[2], *q = W[3], *u = W[4], *v = W[5], *d = W[6], *tmp = W[7];
    for (double* z : {r0, rc, p, u}) MFB_CUDA(cudaMemcpyAsync(z, r, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    for (double* z : {q, d, tmp}) MFB_CUDA(cudaMemsetAsync(z, 0, n * sizeof(double), ctx->stream));
    MFB_TRY(S.mul(v, p));
    double rr;
    MFB_TRY(V.dot(r, r, &rr));
    double r_norm = std::sqrt(rr), r_norm_old, tau = r_norm, rho = rr, rhobar, theta = 0.0, eta = 0.0, alpha, beta, c;
    while (true) {
        double dv;
        MFB_TRY(V.dot(v, r0, &dv));
        alpha = rho / dv;
        MFB_TRY(V.set(q, {1.0, -alpha}, {u, v}));                         // q = u - alpha v
        MFB_TRY(V.set(v, {1.0, 1.0}, {u, q}));                            // v = u + q
        MFB_TRY(S.mul(tmp, v));
        double n2;
        MFB_TRY(V.add(rc, {-alpha}, {tmp}, &n2));                         // r_cgs -= alpha tmp
        r_norm_old = r_norm;
        r_norm = std::sqrt(n2);
        MFB_TRY(S.lincomb(d, theta * theta * eta / alpha, 1, std::initializer_list<double>{1.0}.begin(),
                          std::initializer_list<const double*>{u}.begin()));       // d = u + (theta^2 eta / alpha) d
        theta = r_norm_old / tau;
        c = 1 / std::sqrt(1 + theta * theta);
        tau *= theta * c;
        eta = c * c * alpha;
        MFB_TRY(V.add(x, {eta}, {d}));
        MFB_TRY(S.lincomb(d, theta * theta * eta / alpha, 1, std::initializer_list<double>{1.0}.begin(),
                          std::initializer_list<const double*>{q}.begin()));       // d = q + (theta^2 eta / alpha) d
        theta = std::sqrt(r_norm * r_norm_old) / tau;
        c = 1 / std::sqrt(1 + theta * theta);
        tau *= theta * c;
        eta = c * c * alpha;
        MFB_TRY(V.add(x, {eta}, {d}));
        rhobar = rho;
        MFB_TRY(V.dot(rc, r0, &rho));
        beta = rho / rhobar;
        MFB_TRY(V.set(u, {1.0, beta}, {rc, q}));                          // u = r_cgs + beta q
        MFB_TRY(S.lincomb(p, beta * beta, 2, std::initializer_list<double>{1.0, beta}.begin(),
                          std::initializer_list<const double*>{u, q}.begin()));    // p = u + beta (q + beta p)
        MFB_TRY(S.mul(v, p));
        iter++;
        if (iter > maxiter) { *iters = iter; return MFB_OK; }
        if (iter % checkiter == 0) {
            MFB_TRY(true_residual(S, r, b, x, &res));
            if (res <= tol) { *iters = iter; return MFB_OK; }
        }
    }
}

// lsqr!  (06_LSQR.jl:10-73). Vectors: u, v, w, tmp.
int lsqr(Solver& S, double* x, const double* b, double* r, double tol, int maxiter, std::vector<double*>& W, int* iters) {
    mfb_ctx* ctx = S.ctx;
    const int64_t n = S.n;
    Vec V{S};
    double res;
    MFB_TRY(true_residual(S, r, b, x, &res));
    if (res <= tol) { *iters = 0; return MFB_OK; }
    int iter = 1;
    double *u = W[0], *v = W[1], *w = W[2], *tmp = W[3];
    double n2;
    MFB_TRY(V.dot(r, r, &n2));
    double beta = std::sqrt(n2);
    MFB_TRY(V.set(u, {1.0 / beta}, {r}));
    MFB_TRY(S.tmul(v, u));
    MFB_TRY(V.dot(v, v, &n2));
    double alpha = std::sqrt(n2);
    if (alpha != 0) MFB_TRY(V.set(v, {1.0 / alpha}, {v}));
    MFB_CUDA(cudaMemcpyAsync(w, v, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    double phibar = beta, rhobar = alpha;
    while (true) {
        MFB_TRY(S.mul(tmp, v));
        MFB_TRY(S.lincomb(u, -alpha, 1, std::initializer_list<double>{1.0}.begin(),
                          std::initializer_list<const double*>{tmp}.begin(), &n2));           // u = Pl(A v) - alpha u
        beta = std::sqrt(n2);
        if (beta != 0) {
            MFB_TRY(V.set(u, {1.0 / beta}, {u}));
            MFB_TRY(S.tmul(tmp, u));
            MFB_TRY(S.lincomb(v, -beta, 1, std::initializer_list<double>{1.0}.begin(),
                              std::initializer_list<const double*>{tmp}.begin(), &n2));       // v = Pl(A' u) - beta v
            alpha = std::sqrt(n2);
            if (alpha != 0) MFB_TRY(V.set(v, {1.0 / alpha}, {v}));
        }
        const double rho = std::sqrt(rhobar * rhobar + beta * beta);
        const double c = rhobar / rho, sn = beta / rho, theta = sn * alpha;
        rhobar = -c * alpha;
        const double phi = c * phibar;
        phibar = sn * phibar;
        MFB_TRY(V.add(x, {phi / rho}, {w}));
        MFB_TRY(S.lincomb(w, -theta / rho, 1, std::initializer_list<double>{1.0}.begin(),
                          std::initializer_list<const double*>{v}.begin()));                  // w = v - (theta / rho) w
        iter++;
        MFB_TRY(true_residual(S, r, b, x, &res));
        if (res <= tol || iter > maxiter) { *iters = iter; return MFB_OK; }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
extern "C" int mfb_krylov_solve_ex(mfb_ctx* ctx, int method, int s, int maxiter, int max_pass, double tol, uint64_t seed,
                                   int pr_mode, int pl_mode, int checkiter, double* delta_out, mfb_solve_info* info) {
    if (!ctx) return MFB_ERR_ARG;
    MFB_REQUIRE(ctx->U > 0 && ctx->K_total.p, MFB_ERR_STATE, "mfb_krylov_solve: pattern/matrix not built");
    MFB_REQUIRE(method >= MFB_IDRS && method <= MFB_IDRS_ORIGINAL, MFB_ERR_ARG, "unknown Krylov method");
    MFB_REQUIRE(pr_mode >= MFB_PR_JACOBI && pr_mode <= MFB_PR_IDENTITY && pl_mode >= MFB_PL_IDENTITY && pl_mode <= MFB_PL_ILU,
                MFB_ERR_ARG, "unknown preconditioner mode");
    MFB_REQUIRE(s >= 1 && (method == MFB_GMRES ? s + 1 <= MAXT : (2 * s + 2 <= MAXT && s + 2 <= MAXB && s <= MAXD)), MFB_ERR_ARG,
                "s out of range");
    MFB_REQUIRE(!(mfb_is_distributed(ctx) && (pr_mode == MFB_PR_JACOBI_COLUMN || pl_mode == MFB_PL_JACOBI_ROW || pl_mode == MFB_PL_ILU)), MFB_ERR_ARG,
                "row/column-norm Jacobi and ILU need assembled rows: not available on a partitioned mesh");
    MFB_CUDA(cudaSetDevice(ctx->device));
    MFB_TRY(ensure_scalars(ctx));
    ProfScope ps(ctx, MFB_T_SOLVE);
    const int nv = ctx->n_var;
    const int64_t n = ctx->N * nv;
    const int64_t nval = ctx->U * nv * nv;
    // work vectors of the method (beyond x and r)
    int nw = 0;
    switch (method) {
        case MFB_IDRS: nw = 3 * s + 1; break;
        case MFB_BICGSTABL_GS: case MFB_BICGSTABL: nw = 2 * s + 2; break;
        case MFB_GMRES: nw = s + 1; break;
        case MFB_CGS: nw = 5; break;
        case MFB_CGS2: case MFB_TFQMR: nw = 8; break;
        case MFB_LSQR: nw = 4; break;
        case MFB_IDRS_ORIGINAL: nw = 3 * s + 3; break;
    }
    const int nvec = nw + 3;  // + x, r, (spare)
    MFB_TRY(ensure_work(ctx, nvec, n));
    DevBuf<double>& Ks = ctx->Ks;  // scaled matrix copy: allocated once and kept across solves (the reference allocates K_vals per solve)
    DevBuf<double> plv; // Pl_Jacobi vector
    MFB_CUDA(Ks.alloc(nval + (size_t)MFB_STREAM_PAD * nv * nv));
    MFB_CUDA(ctx->jac.alloc(n));
    MFB_CUDA(ctx->delta.alloc(n));
    std::vector<int> diag_ok(nv);
    for (int v = 0; v < nv; ++v) diag_ok[v] = ctx->block_of[v * nv + v] >= 0;
    DevBuf<int> d_ok;
    MFB_CUDA(d_ok.alloc(nv));
    MFB_CUDA(cudaMemcpyAsync(d_ok.p, diag_ok.data(), nv * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    const unsigned char* own = mfb_is_distributed(ctx) ? ctx->owned.p : (const unsigned char*)nullptr;
    // ---- right preconditioner: Pr_Jacobi! (02_Preconditioner.jl:103-120) or Identity ----
    if (pr_mode == MFB_PR_JACOBI) {
        LAUNCH(k_jacobi_diag, nblk(n), TPB, ctx->nodeptr.p, ctx->nodecol.p, ctx->K_total.p, d_ok.p, ctx->N, nv, own, ctx->jac.p);
        MFB_TRY(mfb_halo_add(ctx, ctx->jac.p, nv));
        LAUNCH(k_jacobi_abs, nblk(n), TPB, ctx->jac.p, n);
    } else if (pr_mode == MFB_PR_JACOBI_COLUMN) {
        MFB_CUDA(cudaMemsetAsync(ctx->jac.p, 0, n * sizeof(double), ctx->stream));
        LAUNCH(k_jacobi_sq, (unsigned)ctx->N, 32, ctx->nodeptr.p, ctx->nodecol.p, ctx->K_total.p, ctx->N, nv, 0, ctx->jac.p);
        LAUNCH(k_sqrt, nblk(n), TPB, ctx->jac.p, n);
    } else {
        LAUNCH(k_fill1, nblk(n), TPB, ctx->jac.p, n);
    }
    LAUNCH(k_scale_copy, nblk(nval), TPB, ctx->nodecol.p, ctx->K_total.p, ctx->jac.p, nval, nv, Ks.p);
    // ---- left preconditioner: Pl_Jacobi (:150-166), computed from the already right-scaled matrix (:38-40) ----
    bool use_ilu = false;
    if (pl_mode == MFB_PL_ILU) {                                 // Pl_ILU(A) on the right-scaled matrix (:38-40, 179-194)
        MFB_TRY(mfb_ilu_factor(ctx, Ks.p, nullptr));
        use_ilu = true;
    } else if (pl_mode != MFB_PL_IDENTITY) {
        MFB_CUDA(plv.alloc(n));
        if (pl_mode == MFB_PL_JACOBI) {
            LAUNCH(k_jacobi_diag, nblk(n), TPB, ctx->nodeptr.p, ctx->nodecol.p, Ks.p, d_ok.p, ctx->N, nv, own, plv.p);
            MFB_TRY(mfb_halo_add(ctx, plv.p, nv));
            LAUNCH(k_jacobi_abs, nblk(n), TPB, plv.p, n);
        } else {
            MFB_CUDA(cudaMemsetAsync(plv.p, 0, n * sizeof(double), ctx->stream));
            LAUNCH(k_jacobi_sq, (unsigned)ctx->N, 32, ctx->nodeptr.p, ctx->nodecol.p, Ks.p, ctx->N, nv, 1, plv.p);
            LAUNCH(k_sqrt, nblk(n), TPB, plv.p, n);
        }
    }
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    d_ok.release();
    std::vector<double*> W(nvec - 2);
    for (int i = 0; i < nvec - 2; ++i) W[i] = ctx->work[i].p;
    double* x = ctx->work[nvec - 2].p;
    double* r = ctx->work[nvec - 1].p;
    const double* b = ctx->residue.p;
    MFB_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double), ctx->stream));
    MFB_CUDA(cudaMemcpyAsync(r, b, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    Solver S{ctx, n, Ks.p};
    S.n_global = ctx->n_global_nodes * nv;
    S.pl = plv.p;
    S.pl_ilu = use_ilu;
    mfb_solve_info inf;
    memset(&inf, 0, sizeof(inf));
    {
        double d;
        MFB_TRY(S.dot1(b, b, &d));
        inf.initial_residual = S.nn(d);
    }
    int pass = 1;
    double res = inf.initial_residual, tol_factor = 1.0;
    const char* lenv = getenv("MFB_KRYLOV_LEGACY");
    const bool legacy = lenv && lenv[0] == '1';       // round-1 launch structure (A/B measurements)
    while (true) {
        int it = 0;
        const double ptol = tol_factor * tol;
        switch (method) {
            case MFB_IDRS:
                if (legacy) MFB_TRY(idrs_legacy(S, x, b, r, ptol, maxiter, s, seed, pass, W, &it));
                else MFB_TRY(idrs(S, x, b, r, ptol, maxiter, s, seed, pass, W, &it));
                break;
            case MFB_BICGSTABL_GS:
                if (legacy) MFB_TRY(bicgstabl_gs_legacy(S, x, b, r, ptol, maxiter, s, seed, pass, W, &it));
                else MFB_TRY(bicgstabl_gs(S, x, b, r, ptol, maxiter, s, seed, pass, W, &it));
                break;
            case MFB_BICGSTABL: MFB_TRY(bicgstabl_lu(S, x, b, r, ptol, maxiter, s, seed, pass, W, &it)); break;
            case MFB_GMRES: MFB_TRY(gmres(S, x, b, r, ptol, maxiter, s, W, &it)); break;
            case MFB_CGS: MFB_TRY(cgs(S, x, b, r, ptol, maxiter, W, &it)); break;
            case MFB_CGS2: MFB_TRY(cgs2(S, x, b, r, ptol, maxiter, seed, pass, W, &it)); break;
            case MFB_TFQMR: MFB_TRY(tfqmr(S, x, b, r, ptol, maxiter, checkiter > 0 ? checkiter : 200, W, &it)); break;
            case MFB_LSQR: MFB_TRY(lsqr(S, x, b, r, ptol, maxiter, W, &it)); break;
            case MFB_IDRS_ORIGINAL: MFB_TRY(idrs_original(S, x, b, r, ptol, maxiter, s, seed, pass, W, &it)); break;
        }
        inf.iterations += it;
        MFB_TRY(true_residual(S, r, b, x, &res, false));        // the plain b - A x (:45-48)
        if (S.has_pl()) {                                        // left preconditioned: rescale the next pass's tolerance (:50-53)
            MFB_TRY(S.Pl(r));
            double d;
            MFB_TRY(S.dot1(r, r, &d));
            const double pres = S.nn(d);
            tol_factor = std::min(pres / res, 1.0);
        }
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
    plv.release();
    MFB_TRY(mfb_p2p_check(ctx));
    if (info) *info = inf;
    return inf.converged ? MFB_OK : MFB_NOT_CONVERGED;
}

extern "C" int mfb_krylov_solve(mfb_ctx* ctx, int method, int s, int maxiter, int max_pass, double tol, uint64_t seed,
                                double* delta_out, mfb_solve_info* info) {
    return mfb_krylov_solve_ex(ctx, method, s, maxiter, max_pass, tol, seed, MFB_PR_JACOBI, MFB_PL_IDENTITY, 200, delta_out, info);
}

// Development aid (bench.py --spmv-sweep): times `reps` launches of one (unroll, min-blocks, rows-per-CTA) variant of the
// NV = 3 block SpMV on K_total with CUDA events; variant 0 is the production kernel. Returns ms per launch.
extern "C" int mfb_spmv_variant_bench(mfb_ctx* ctx, int variant, int reps, double* ms_out, double* max_abs_diff) {
    if (!ctx || !ms_out) return MFB_ERR_ARG;
    MFB_REQUIRE(ctx->U > 0 && ctx->n_var == 3, MFB_ERR_STATE, "mfb_spmv_variant_bench: needs a built 3-variable system");
    MFB_CUDA(cudaSetDevice(ctx->device));
    const int64_t N = ctx->N, n = N * 3;
    MFB_TRY(ensure_work(ctx, 2, n));
    double *x = ctx->work[0].p, *y = ctx->work[1].p;
    LAUNCH(k_rand, RED_BLOCKS, TPB, x, n, 7ull, 1ull, ctx->gid.p, 3);
    cudaEvent_t e0, e1;
    MFB_CUDA(cudaEventCreate(&e0));
    MFB_CUDA(cudaEventCreate(&e1));
    const int* np = ctx->nodeptr.p; const int* nc = ctx->nodecol.p; const double* K = ctx->K_total.p;
    auto launch = [&](int v) {
        auto g = [&](int rows) { return (unsigned)((N + rows - 1) / rows); };
        switch (v) {
            case 0: LAUNCH((k_spmv_bsr<3>), g(64), 256, np, nc, K, x, y, N); break;
            case 1: LAUNCH((k_spmv_bsr<3, 5, 6, 64>), g(64), 256, np, nc, K, x, y, N); break;
            case 2: LAUNCH((k_spmv_bsr<3, 4, 6, 64>), g(64), 256, np, nc, K, x, y, N); break;
            case 3: LAUNCH((k_spmv_bsr<3, 4, 8, 64>), g(64), 256, np, nc, K, x, y, N); break;
            case 4: LAUNCH((k_spmv_bsr<3, 3, 8, 64>), g(64), 256, np, nc, K, x, y, N); break;
            case 5: LAUNCH((k_spmv_bsr<3, 6, 5, 64>), g(64), 256, np, nc, K, x, y, N); break;
            case 6: LAUNCH((k_spmv_bsr<3, 5, 5, 128>), g(128), 256, np, nc, K, x, y, N); break;
            case 7: LAUNCH((k_spmv_bsr<3, 5, 5, 32>), g(32), 256, np, nc, K, x, y, N); break;
            case 8: LAUNCH((k_spmv_bsr<3, 4, 6, 128>), g(128), 256, np, nc, K, x, y, N); break;
            case 9: LAUNCH((k_spmv_bsr<3, 2, 8, 64>), g(64), 256, np, nc, K, x, y, N); break;
            case 10: LAUNCH((k_spmv_bsr_x<3, 5, 64>), g(64), 256, np, nc, K, x, y, N); break;
            case 11: LAUNCH((k_spmv_bsr_x<3, 4, 64>), g(64), 256, np, nc, K, x, y, N); break;
            case 12: LAUNCH((k_spmv_bsr_x<3, 3, 64>), g(64), 256, np, nc, K, x, y, N); break;
            case 13: LAUNCH((k_spmv_bsr_x<3, 5, 128>), g(128), 256, np, nc, K, x, y, N); break;
            // multi-row streams: RW rows per warp (8 warps per CTA), register cap for 4 / 5 CTAs per SM, unroll 5 / 4 / 6
            case 14: LAUNCH((k_spmv_mr<3, 5, 8, false, 4>), g(64), 256, np, nc, K, x, y, N, (const double*)nullptr, (double*)nullptr, (double*)nullptr, (double*)nullptr, (unsigned*)nullptr, 0); break;
            case 15: LAUNCH((k_spmv_mr<3, 5, 4, false, 4>), g(32), 256, np, nc, K, x, y, N, (const double*)nullptr, (double*)nullptr, (double*)nullptr, (double*)nullptr, (unsigned*)nullptr, 0); break;
            case 16: LAUNCH((k_spmv_mr<3, 5, 16, false, 4>), g(128), 256, np, nc, K, x, y, N, (const double*)nullptr, (double*)nullptr, (double*)nullptr, (double*)nullptr, (unsigned*)nullptr, 0); break;
            case 17: LAUNCH((k_spmv_mr<3, 5, 8, false, 5>), g(64), 256, np, nc, K, x, y, N, (const double*)nullptr, (double*)nullptr, (double*)nullptr, (double*)nullptr, (unsigned*)nullptr, 0); break;
            case 18: LAUNCH((k_spmv_mr<3, 4, 8, false, 5>), g(64), 256, np, nc, K, x, y, N, (const double*)nullptr, (double*)nullptr, (double*)nullptr, (double*)nullptr, (unsigned*)nullptr, 0); break;
            case 19: LAUNCH((k_spmv_mr<3, 6, 8, false, 4>), g(64), 256, np, nc, K, x, y, N, (const double*)nullptr, (double*)nullptr, (double*)nullptr, (double*)nullptr, (unsigned*)nullptr, 0); break;
            case 20: launch_tma<3, 8, 3, 8>(ctx, K, x, y, nullptr); break;
            case 21: launch_tma<3, 8, 2, 8>(ctx, K, x, y, nullptr); break;
            case 22: launch_tma<3, 16, 3, 8>(ctx, K, x, y, nullptr); break;
            case 23: launch_tma<3, 8, 4, 4>(ctx, K, x, y, nullptr); break;
            case 24: launch_tma<3, 16, 2, 8>(ctx, K, x, y, nullptr); break;
            case 25: launch_tma<3, 16, 2, 4, 4, 4>(ctx, K, x, y, nullptr); break;      // 4 warps, 2 stages: 36 kB per CTA, up to 4 CTAs
            case 26: launch_tma<3, 16, 3, 4, 4, 4>(ctx, K, x, y, nullptr); break;      // 4 warps, 3 stages: 55 kB per CTA
            // load-policy flavours of the multi-row kernel (streams not allocated in L1, evict-first in L2)
            case 27: LAUNCH((k_spmv_mr<3, 5, 8, false, 4, 1>), g(64), 256, np, nc, K, x, y, N, (const double*)nullptr, (double*)nullptr, (double*)nullptr, (double*)nullptr, (unsigned*)nullptr, 0); break;
            case 28: LAUNCH((k_spmv_mr<3, 5, 8, false, 4, 2>), g(64), 256, np, nc, K, x, y, N, (const double*)nullptr, (double*)nullptr, (double*)nullptr, (double*)nullptr, (unsigned*)nullptr, 0); break;
            case 29: LAUNCH((k_spmv_mr<3, 5, 16, false, 4, 1>), g(128), 256, np, nc, K, x, y, N, (const double*)nullptr, (double*)nullptr, (double*)nullptr, (double*)nullptr, (unsigned*)nullptr, 0); break;
            case 30: LAUNCH((k_spmv_mr<3, 5, 16, false, 4, 2>), g(128), 256, np, nc, K, x, y, N, (const double*)nullptr, (double*)nullptr, (double*)nullptr, (double*)nullptr, (unsigned*)nullptr, 0); break;
            case 31: LAUNCH((k_spmv_mr<3, 5, 24, false, 4, 0>), g(192), 256, np, nc, K, x, y, N, (const double*)nullptr, (double*)nullptr, (double*)nullptr, (double*)nullptr, (unsigned*)nullptr, 0); break;
            default: break;
        }
    };
    MFB_REQUIRE(variant >= 0 && variant <= 31, MFB_ERR_ARG, "unknown SpMV variant");
    for (int i = 0; i < 3; ++i) launch(variant);
    MFB_CUDA(cudaEventRecord(e0, ctx->stream));
    for (int i = 0; i < reps; ++i) launch(variant);
    MFB_CUDA(cudaEventRecord(e1, ctx->stream));
    MFB_CUDA(cudaEventSynchronize(e1));
    MFB_CUDA(cudaGetLastError());
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms_out = ms / reps;
    if (max_abs_diff) {                                   // compare with the production kernel on the same x
        std::vector<double> yv(n), y0(n);
        MFB_CUDA(cudaMemcpyAsync(yv.data(), y, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        launch(0);
        MFB_CUDA(cudaMemcpyAsync(y0.data(), y, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        double d = 0.0;
        for (int64_t q = 0; q < n; ++q) d = std::max(d, std::fabs(yv[q] - y0[q]));
        *max_abs_diff = d;
    }
    return MFB_OK;
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
    MFB_TRY(mfb_halo_add(ctx, yi.p, ctx->n_var));            // partitioned mesh: complete the interface rows
    MFB_TRY(mfb_to_reference(ctx, yi.p, yr.p, 1));
    MFB_TRY(mfb_stage_out(ctx, yr.p, n * sizeof(double), y));
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    xr.release(); xi.release(); yi.release(); yr.release();
    return mfb_p2p_check(ctx);
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
    return mfb_p2p_check(ctx);
}
