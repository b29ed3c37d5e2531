// mfb_ilu.cu -- Pl_ILU: incomplete LU left preconditioner on the block-CSR matrix.
//
// Reference: Pl_ILU(A) = ilu02!(copy(A)) and two cuSPARSE triangular solves per application
// (src/solver/linear_solver/02_Preconditioner.jl:179-194), on the right-Jacobi-scaled matrix (:38-40). cuSPARSE factorises in
// the matrix's own row order and finds its parallelism by level analysis; on a 3-D FEM graph in a locality-preserving order
// that gives O(10^3) dependency levels, i.e. thousands of tiny launches per application. Here the SAME incomplete
// factorisation -- zero fill, (L U)_ij = A_ij on the pattern -- is taken in an elimination order chosen for the machine: the
// nodes are coloured greedily in the priority order of a hash of their global id, and eliminated colour class by colour class
// (ties by the hash). The dependency levels are then bounded by the number of colours (32 at 88^3 hex20; a plain hash order has
// 127, MFB_ILU_ORDER=hash), each level is one launch, and the iteration counts were lower than with the plain hash order. The
// blocks are the n_var x n_var node blocks of the library's matrix format: block ILU(0) without pivoting, unit lower factor
// (L_ik = A_ik U_kk^-1), the inverted diagonal blocks kept aside.
//   setup  (per pattern)  colouring, elimination rank, lower/upper flags per entry, levels, rows by level, lower entries in
//                         elimination order, index arrays of the packed factors
//   factor (per solve)    level by level, one warp per row: for every lower entry k in order, L_ik and the update of row i by the
//                         upper part of row k (positions found by binary search in row i); then the inverse of U_ii. Then the
//                         factors are PACKED for the sweeps: per direction, rows in level order with their entries contiguous,
//                         values rounded to FP32 (see ilu_use_f32), upper rows pre-multiplied by U_ii^-1
//   apply  (per product)  forward sweep over the levels (unit L), backward sweep, in place, each level on the multi-row stream
//                         loop of the SpMV (mfb_sweep_level_mr in mfb_krylov.cu); one-warp-per-row kernels remain for n_var > 4
//                         and as A/B switches (MFB_ILU_SWEEP=row, MFB_ILU_UNPACKED=1)
// Everything a row receives is written by its own warp: the factorisation and the sweeps are bit-reproducible.
// Measurements and what bounds the sweeps: profiles/ilu_r2.md.
#include <cstdio>
#include <cstring>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>

#include "mfb_internal.h"

namespace {
constexpr int TPB = 256;
inline unsigned nblk(int64_t n) { return (unsigned)((n + TPB - 1) / TPB); }
typedef unsigned long long u64;

__device__ __forceinline__ u64 mix64(u64 z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void k_prio(u64* key, int64_t N, const long long* gid) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < N) key[i] = mix64((u64)gid[i] + 0x5bd1e995ull);      // hash of the GLOBAL node id: the order does not depend on the internal numbering
}
// one relaxation round of the greedy colouring in priority order: colour[i] = smallest colour that no neighbour of HIGHER priority
// carries (fixed point = the sequential greedy colouring by descending priority; reached after as many rounds as the priority
// order has dependency levels). At most 128 colours (the node graph's degree bound is 81).
__global__ void k_color_round(const int* nodeptr, const int* nodecol, const u64* key, int64_t N, const int* cin, int* cout, int* changed) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const u64 ki = key[i];
    u64 m0 = 0, m1 = 0;
    for (int p = nodeptr[i]; p < nodeptr[i + 1]; ++p) {
        const int j = nodecol[p];
        if (j == (int)i) continue;
        const u64 kj = key[j];
        if (kj > ki || (kj == ki && j > (int)i)) {
            const int c = cin[j];
            if (c < 64) m0 |= 1ull << c; else m1 |= 1ull << (c - 64 < 63 ? c - 64 : 63);
        }
    }
    const int c = ~m0 ? __ffsll((long long)~m0) - 1 : 64 + __ffsll((long long)~m1) - 1;
    cout[i] = c;
    if (c != cin[i]) *changed = 1;
}
__global__ void k_color_key(u64* key, const int* color, int64_t N) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < N) key[i] = ((u64)color[i] << 56) | (key[i] >> 8);
}
__global__ void k_invert_perm(const int* order, int64_t N, int* pos) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r < N) pos[order[r]] = (int)r;
}
// kind[p] of entry p = (i, col): 0 lower (col eliminated before i), 1 diagonal, 2 upper
__global__ void k_entry_kind(const int* nodeptr, const int* nodecol, const int* pos, int64_t N, unsigned char* kind) {
    int64_t i = blockIdx.x;
    if (i >= N) return;
    const int pi = pos[i];
    for (int p = nodeptr[i] + threadIdx.x; p < nodeptr[i + 1]; p += blockDim.x) {
        const int c = nodecol[p];
        kind[p] = c == (int)i ? 1 : (pos[c] < pi ? 0 : 2);
    }
}
// one relaxation round of level[i] = 1 + max over lower neighbours
__global__ void k_level_round(const int* nodeptr, const int* nodecol, const unsigned char* kind, int64_t N, const int* lin, int* lout, int* changed) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    int l = 0;
    for (int p = nodeptr[i]; p < nodeptr[i + 1]; ++p)
        if (kind[p] == 0) { const int lk = lin[nodecol[p]] + 1; l = lk > l ? lk : l; }
    lout[i] = l;
    if (l != lin[i]) *changed = 1;
}
// lperm[row segment]: positions of the row's LOWER entries in elimination order (ascending pos of the column); nlow[i] = their number
__global__ void k_lower_order(const int* nodeptr, const int* nodecol, const unsigned char* kind, const int* pos, int64_t N, int* lperm, int* nlow) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int s = nodeptr[i], t = nodeptr[i + 1];
    int n = 0;
    for (int p = s; p < t; ++p)
        if (kind[p] == 0) {                                   // insertion sort by pos of the column
            const int key = pos[nodecol[p]];
            int q = n++;
            while (q > 0 && pos[nodecol[lperm[s + q - 1]]] > key) { lperm[s + q] = lperm[s + q - 1]; --q; }
            lperm[s + q] = p;
        }
    nlow[i] = n;
}

template <int NV> __device__ __forceinline__ void ld_block(const double* p, double (&a)[NV * NV]) {
#pragma unroll
    for (int q = 0; q < NV * NV; ++q) a[q] = __ldcg(p + q);
}
// c = a * b (NV x NV, row-major)
template <int NV> __device__ __forceinline__ void mul_block(const double (&a)[NV * NV], const double (&b)[NV * NV], double (&c)[NV * NV]) {
#pragma unroll
    for (int r = 0; r < NV; ++r)
#pragma unroll
        for (int s = 0; s < NV; ++s) {
            double v = 0.0;
#pragma unroll
            for (int m = 0; m < NV; ++m) v += a[r * NV + m] * b[m * NV + s];
            c[r * NV + s] = v;
        }
}
// in-place inverse by Gauss-Jordan without pivoting (block ILU(0) takes the diagonal blocks as they come)
template <int NV> __device__ __forceinline__ void inv_block(double (&a)[NV * NV]) {
    double b[NV * NV];
#pragma unroll
    for (int q = 0; q < NV * NV; ++q) b[q] = (q / NV == q % NV) ? 1.0 : 0.0;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        const double d = 1.0 / a[c * NV + c];
#pragma unroll
        for (int s = 0; s < NV; ++s) { a[c * NV + s] *= d; b[c * NV + s] *= d; }
#pragma unroll
        for (int r = 0; r < NV; ++r)
            if (r != c) {
                const double f = a[r * NV + c];
#pragma unroll
                for (int s = 0; s < NV; ++s) { a[r * NV + s] -= f * a[c * NV + s]; b[r * NV + s] -= f * b[c * NV + s]; }
            }
    }
#pragma unroll
    for (int q = 0; q < NV * NV; ++q) a[q] = b[q];
}

// factorisation of the rows of one level: one warp per row
template <int NV>
__global__ void __launch_bounds__(256) k_ilu_factor(const int* __restrict__ rows, int n_rows, const int* __restrict__ nodeptr,
                                                    const int* __restrict__ nodecol, const unsigned char* __restrict__ kind,
                                                    const int* __restrict__ lperm, const int* __restrict__ nlow, double* F, double* dinv) {
    constexpr int B = NV * NV;
    const int w = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= n_rows) return;
    const int i = rows[w];
    const int s = nodeptr[i], t = nodeptr[i + 1];
    const int nl = nlow[i];
    for (int q = 0; q < nl; ++q) {
        const int p = lperm[s + q];
        const int k = nodecol[p];
        double a[B], d[B], L[B];
        ld_block<NV>(F + (size_t)p * B, a);
        ld_block<NV>(dinv + (size_t)k * B, d);
        mul_block<NV>(a, d, L);                                  // L_ik = A_ik U_kk^-1 (every lane computes it, lane 0 stores it)
        if (lane == 0) {
#pragma unroll
            for (int m = 0; m < B; ++m) __stcg(F + (size_t)p * B + m, L[m]);
        }
        // row i -= L_ik * (upper part of row k), where the pattern of row i has the column
        const int ks = nodeptr[k], kt = nodeptr[k + 1];
        for (int r = ks + lane; r < kt; r += 32) {
            if (kind[r] != 2) continue;
            const int j = nodecol[r];
            int lo = s, hi = t - 1;                              // binary search of column j in row i (columns ascending)
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (nodecol[mid] < j) lo = mid + 1; else hi = mid; }
            if (nodecol[lo] != j) continue;                      // zero fill: outside the pattern the update is dropped
            double u[B], x[B], pr[B];
            ld_block<NV>(F + (size_t)r * B, u);
            ld_block<NV>(F + (size_t)lo * B, x);
            mul_block<NV>(L, u, pr);
#pragma unroll
            for (int m = 0; m < B; ++m) __stcg(F + (size_t)lo * B + m, x[m] - pr[m]);
        }
        __syncwarp();                                            // the next lower entry may have been updated by another lane
        __threadfence_block();
    }
    if (lane == 0) {                                             // U_ii^-1
        int lo = s, hi = t - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (nodecol[mid] < i) lo = mid + 1; else hi = mid; }
        double a[B];
        ld_block<NV>(F + (size_t)lo * B, a);
        inv_block<NV>(a);
#pragma unroll
        for (int m = 0; m < B; ++m) __stcg(dinv + (size_t)i * B + m, a[m]);
    }
}

// one level of a triangular sweep, one warp per row. LOWER: v_i -= sum_{lower} L_ik v_k; else v_i = U_ii^-1 (v_i - sum_{upper} U_ij v_j)
template <int NV, bool LOWER>
__global__ void __launch_bounds__(256) k_ilu_sweep(const int* __restrict__ rows, int n_rows, const int* __restrict__ nodeptr,
                                                   const int* __restrict__ nodecol, const unsigned char* __restrict__ kind,
                                                   const double* __restrict__ F, const double* __restrict__ dinv, double* v) {
    constexpr int B = NV * NV;
    const int w = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= n_rows) return;
    const int i = rows[w];
    double acc[NV];
#pragma unroll
    for (int r = 0; r < NV; ++r) acc[r] = 0.0;
    for (int p = nodeptr[i] + lane; p < nodeptr[i + 1]; p += 32) {
        if (kind[p] != (LOWER ? 0 : 2)) continue;
        const double* f = F + (size_t)p * B;
        const double* x = v + (size_t)nodecol[p] * NV;
        double xv[NV];
#pragma unroll
        for (int m = 0; m < NV; ++m) xv[m] = __ldcg(x + m);     // written by an earlier launch of this sweep
#pragma unroll
        for (int r = 0; r < NV; ++r)
#pragma unroll
            for (int m = 0; m < NV; ++m) acc[r] += __ldg(f + r * NV + m) * xv[m];
    }
#pragma unroll
    for (int r = 0; r < NV; ++r)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
    if (lane == 0) {
        double y[NV];
#pragma unroll
        for (int r = 0; r < NV; ++r) y[r] = v[(size_t)i * NV + r] - acc[r];
        if (LOWER) {
#pragma unroll
            for (int r = 0; r < NV; ++r) v[(size_t)i * NV + r] = y[r];
        } else {
#pragma unroll
            for (int r = 0; r < NV; ++r) {
                double z = 0.0;
#pragma unroll
                for (int m = 0; m < NV; ++m) z += __ldg(dinv + (size_t)i * B + r * NV + m) * y[m];
                v[(size_t)i * NV + r] = z;
            }
        }
    }
}

// ---- packed sweeps ----
__global__ void k_pack_counts(const int* rows, const int* nodeptr, const int* nlow, int64_t N, int* cl, int* cu) {
    int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (w >= N) return;
    const int i = rows[w];
    cl[w] = nlow[i];
    cu[w] = nodeptr[i + 1] - nodeptr[i] - nlow[i] - 1;
}
__global__ void k_pack_index(const int* rows, const int* nodeptr, const int* nodecol, const unsigned char* kind, const int* lperm,
                             const int* nlow, const int* Lptr, const int* Uptr, int64_t N, int* Lcol, int* Lsrc, int* Ucol, int* Usrc) {
    int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (w >= N) return;
    const int i = rows[w];
    const int s = nodeptr[i], t = nodeptr[i + 1];
    for (int q = 0; q < nlow[i]; ++q) { const int p = lperm[s + q]; Lcol[Lptr[w] + q] = nodecol[p]; Lsrc[Lptr[w] + q] = p; }
    int u = Uptr[w];
    for (int p = s; p < t; ++p)
        if (kind[p] == 2) { Ucol[u] = nodecol[p]; Usrc[u] = p; ++u; }
}
template <typename VT>
__global__ void k_pack_values(const int* src, int64_t n, int B, const double* F, VT* out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n * B) return;
    out[t] = (VT)__ldcg(F + (size_t)src[t / B] * B + t % B);
}
// packed upper entries of the stream-kernel sweeps: U'_ij = U_ii^-1 U_ij, so that the backward sweep is z_i = (U_ii^-1 y_i) - sum U'_ij z_j
// and its row epilogue needs no block product (one warp per row)
template <int NV, typename VT>
__global__ void k_pack_upper_scaled(const int* __restrict__ rows, const int* __restrict__ ptr, const int* __restrict__ src, int64_t n_rows,
                                    const double* __restrict__ F, const double* __restrict__ dinv, VT* __restrict__ out) {
    constexpr int B = NV * NV;
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_rows) return;
    double d[B], a[B], c[B];
    ld_block<NV>(dinv + (size_t)rows[w] * B, d);
    for (int j = ptr[w] + lane; j < ptr[w + 1]; j += 32) {
        ld_block<NV>(F + (size_t)src[j] * B, a);
        mul_block<NV>(d, a, c);
#pragma unroll
        for (int q = 0; q < B; ++q) out[(size_t)j * B + q] = (VT)c[q];
    }
}

// one level of a sweep on the packed factors: one warp per row, lane <-> fixed block position (i, k) like the SpMV, the row's
// entries as one coalesced stream. LOWER: v_i -= sum L_ik v_k;  else v_i = U_ii^-1 (v_i - sum U_ij v_j)
template <int NV, bool LOWER>
__global__ void __launch_bounds__(256) k_ilu_sweep_packed(const int* __restrict__ rows, int w0, int n_rows, const int* __restrict__ ptr,
                                                          const int* __restrict__ col, const double* __restrict__ val,
                                                          const double* __restrict__ dinv, double* v) {
    constexpr int B = NV * NV, EPW = 32 / B, ACTIVE = EPW * B;
    const int wl = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (wl >= n_rows) return;
    const int w = w0 + wl;
    const int row = rows[w];
    const int le = lane / B, ik = lane - le * B, i = ik / NV, k = ik - i * NV;
    const bool on = lane < ACTIVE;
    const int s = ptr[w], deg = on ? ptr[w + 1] - s : 0;
    const double* Vp = val + (size_t)s * B + lane;
    const int* Cp = col + s + le;
    double a0 = 0.0, a1 = 0.0;
    int e = le;
    for (; e + EPW < deg; e += 2 * EPW) {
        const double v0 = __ldcs(Vp + (size_t)(e - le) * B), v1 = __ldcs(Vp + (size_t)(e - le + EPW) * B);
        const int c0 = __ldg(Cp + (e - le)), c1 = __ldg(Cp + (e - le + EPW));
        a0 += v0 * __ldcg(v + (size_t)c0 * NV + k);
        a1 += v1 * __ldcg(v + (size_t)c1 * NV + k);
    }
    if (e < deg) a0 += __ldcs(Vp + (size_t)(e - le) * B) * __ldcg(v + (size_t)__ldg(Cp + (e - le)) * NV + k);
    const double acc = a0 + a1;
    double t = acc;
#pragma unroll
    for (int d = 1; d < NV; ++d) t += __shfl_down_sync(0xffffffffu, acc, d);                 // sum over k
    double r = t;
    if constexpr ((B & (B - 1)) == 0) {
#pragma unroll
        for (int o = 16; o >= B; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    } else {
#pragma unroll
        for (int d = 1; d < EPW; ++d) r += __shfl_down_sync(0xffffffffu, t, d * B);          // sum over the EPW entries
    }
    // lanes ik = i * NV (le == 0, k == 0) hold component i of the row's sum
    const double yi = (le == 0 && k == 0 && on) ? v[(size_t)row * NV + i] - r : 0.0;
    if (LOWER) {
        if (le == 0 && k == 0 && on) v[(size_t)row * NV + i] = yi;
    } else {
        double z = 0.0;                                                                        // z_i = sum_m dinv[i][m] y_m
#pragma unroll
        for (int m = 0; m < NV; ++m) {
            const double ym = __shfl_sync(0xffffffffu, yi, m * NV);
            if (le == 0 && k == 0 && on) z += __ldg(dinv + (size_t)row * B + i * NV + m) * ym;
        }
        if (le == 0 && k == 0 && on) v[(size_t)row * NV + i] = z;
    }
}

// defect of the defining property (L U)_ij = A_ij on the pattern: one warp per row, max |.| into out[0], max |A| into out[1]
template <int NV>
__global__ void __launch_bounds__(256) k_ilu_defect(const int* __restrict__ nodeptr, const int* __restrict__ nodecol,
                                                    const unsigned char* __restrict__ kind, const int* __restrict__ pos, const double* F,
                                                    const double* dinv, const double* A, int64_t N, double* out) {
    constexpr int B = NV * NV;
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= N) return;
    const int s = nodeptr[i], t = nodeptr[i + 1];
    double worst = 0.0, amax = 0.0;
    for (int p = s + lane; p < t; p += 32) {                   // entry (i, j)
        const int j = nodecol[p];
        double sum[B];
#pragma unroll
        for (int m = 0; m < B; ++m) sum[m] = 0.0;
        // sum over k eliminated before both i and j, k adjacent to i (lower entry of row i) and j in the upper part of row k
        for (int q = s; q < t; ++q) {
            if (kind[q] != 0) continue;
            const int k = nodecol[q];
            if (!(pos[k] < pos[j])) continue;
            int lo = nodeptr[k], hi = nodeptr[k + 1] - 1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (nodecol[mid] < j) lo = mid + 1; else hi = mid; }
            if (nodecol[lo] != j) continue;
            double L[B], u[B], pr[B];
            ld_block<NV>(F + (size_t)q * B, L);
            ld_block<NV>(F + (size_t)lo * B, u);
            mul_block<NV>(L, u, pr);
#pragma unroll
            for (int m = 0; m < B; ++m) sum[m] += pr[m];
        }
        double own[B], a[B];
        ld_block<NV>(F + (size_t)p * B, own);
        ld_block<NV>(A + (size_t)p * B, a);
        if (kind[p] == 0) {                                     // lower entry: (L U)_ij = sum + L_ij U_jj
            double ujj[B], dj[B], pr[B];
            ld_block<NV>(dinv + (size_t)j * B, dj);
#pragma unroll
            for (int m = 0; m < B; ++m) ujj[m] = dj[m];
            inv_block<NV>(ujj);
            mul_block<NV>(own, ujj, pr);
#pragma unroll
            for (int m = 0; m < B; ++m) sum[m] += pr[m];
        } else {                                                // diagonal / upper entry: (L U)_ij = sum + U_ij (unit L_ii)
#pragma unroll
            for (int m = 0; m < B; ++m) sum[m] += own[m];
        }
#pragma unroll
        for (int m = 0; m < B; ++m) { worst = fmax(worst, fabs(sum[m] - a[m])); amax = fmax(amax, fabs(a[m])); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
        amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    }
    if (lane == 0) {
        atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(worst));     // non-negative doubles order like integers
        atomicMax(reinterpret_cast<unsigned long long*>(out + 1), (unsigned long long)__double_as_longlong(amax));
    }
}
}  // namespace

#define LAUNCH(kernel, grid, block, ...)                          \
    do {                                                          \
        kernel<<<(grid), (block), 0, ctx->stream>>>(__VA_ARGS__); \
        ctx->launches++;                                          \
    } while (0)

struct IluPlan {
    int64_t N = 0, U = 0;
    int nv = 0;
    DevBuf<int> pos, level, rows, lperm, nlow;
    DevBuf<unsigned char> kind;
    std::vector<int> level_ptr;          // rows of level l: rows[level_ptr[l] .. level_ptr[l+1])
    DevBuf<double> F, dinv;              // factors in the matrix's block layout, inverted diagonal blocks
    // packed factors for the sweeps: rows in LEVEL order, the lower (upper) entries of a row contiguous, so that a sweep level is one
    // contiguous range of rows and entries read as a coalesced stream (the layout of the matrix interleaves lower and upper entries)
    DevBuf<int> Lptr, Uptr, Lcol, Ucol, Lsrc, Usrc;   // [N+1], [N+1], [UL], [UU], source position of every packed entry in F
    DevBuf<double> Lval, Uval;
    DevBuf<float> Lval32, Uval32;        // the packed factors rounded to FP32 (default for the sweeps, see ilu_use_f32)
    bool f32 = false;
    bool scaledU = false;                // packed upper entries pre-multiplied by the inverse diagonal block of their row
    int64_t UL = 0, UU = 0;
    bool factored = false;
};
static std::map<mfb_ctx*, IluPlan*>& ilu_table() {
    static std::map<mfb_ctx*, IluPlan*> t;
    return t;
}
void mfb_ilu_free(mfb_ctx* ctx) {
    auto it = ilu_table().find(ctx);
    if (it != ilu_table().end()) { delete it->second; ilu_table().erase(it); }
}

static int ilu_setup(mfb_ctx* ctx, IluPlan*& P) {
    IluPlan*& slot = ilu_table()[ctx];
    if (slot && slot->N == ctx->N && slot->U == ctx->U && slot->nv == ctx->n_var) { P = slot; return MFB_OK; }
    delete slot;
    P = slot = new IluPlan();
    const int64_t N = ctx->N, U = ctx->U;
    P->N = N; P->U = U; P->nv = ctx->n_var;
    auto pol = thrust::cuda::par.on(ctx->stream);
    DevBuf<u64> key;
    DevBuf<int> order, lev2, changed;
    MFB_CUDA(key.alloc(N)); MFB_CUDA(order.alloc(N)); MFB_CUDA(P->pos.alloc(N));
    LAUNCH(k_prio, nblk(N), TPB, key.p, N, ctx->gid.p);
    // elimination by colour classes of a greedy colouring (ties inside a class by the hash): the dependency levels are bounded by
    // the number of colours (32 at 88^3) instead of the depth of the plain hash order (127; MFB_ILU_ORDER=hash, round-2 first version)
    static const bool by_color = [] { const char* e = getenv("MFB_ILU_ORDER"); return !(e && e[0] == 'h'); }();
    if (by_color) {
        DevBuf<int> ca, cb, chg;
        MFB_CUDA(ca.alloc(N)); MFB_CUDA(cb.alloc(N)); MFB_CUDA(chg.alloc(1));
        MFB_CUDA(cudaMemsetAsync(ca.p, 0, N * sizeof(int), ctx->stream));
        int *a = ca.p, *b = cb.p;
        for (int round = 0; round < 100000; ++round) {
            MFB_CUDA(cudaMemsetAsync(chg.p, 0, sizeof(int), ctx->stream));
            LAUNCH(k_color_round, nblk(N), TPB, ctx->nodeptr.p, ctx->nodecol.p, key.p, N, a, b, chg.p);
            int h = 0;
            MFB_CUDA(cudaMemcpyAsync(&h, chg.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            MFB_CUDA(cudaStreamSynchronize(ctx->stream));
            std::swap(a, b);
            if (!h) break;
        }
        LAUNCH(k_color_key, nblk(N), TPB, key.p, a, N);
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    thrust::device_ptr<int> op(order.p);
    thrust::device_ptr<u64> kp(key.p);
    thrust::sequence(pol, op, op + N);
    thrust::stable_sort_by_key(pol, kp, kp + N, op);                  // elimination order: ascending hash (ties by index)
    LAUNCH(k_invert_perm, nblk(N), TPB, order.p, N, P->pos.p);
    MFB_CUDA(P->kind.alloc(U));
    LAUNCH(k_entry_kind, (unsigned)N, 64, ctx->nodeptr.p, ctx->nodecol.p, P->pos.p, N, P->kind.p);
    // dependency levels by relaxation
    MFB_CUDA(P->level.alloc(N)); MFB_CUDA(lev2.alloc(N)); MFB_CUDA(changed.alloc(1));
    MFB_CUDA(cudaMemsetAsync(P->level.p, 0, N * sizeof(int), ctx->stream));
    int *la = P->level.p, *lb = lev2.p;
    for (int round = 0; round < 100000; ++round) {
        MFB_CUDA(cudaMemsetAsync(changed.p, 0, sizeof(int), ctx->stream));
        LAUNCH(k_level_round, nblk(N), TPB, ctx->nodeptr.p, ctx->nodecol.p, P->kind.p, N, la, lb, changed.p);
        int h = 0;
        MFB_CUDA(cudaMemcpyAsync(&h, changed.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        std::swap(la, lb);
        if (!h) break;
    }
    if (la != P->level.p) MFB_CUDA(cudaMemcpyAsync(P->level.p, la, N * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    // rows by level
    MFB_CUDA(P->rows.alloc(N));
    MFB_CUDA(cudaMemcpyAsync(lev2.p, P->level.p, N * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    thrust::device_ptr<int> rp(P->rows.p), lp(lev2.p);
    thrust::sequence(pol, rp, rp + N);
    thrust::stable_sort_by_key(pol, lp, lp + N, rp);
    std::vector<int> hl(N);
    MFB_CUDA(cudaMemcpyAsync(hl.data(), lev2.p, N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    const int n_levels = hl.empty() ? 0 : hl.back() + 1;
    P->level_ptr.assign(n_levels + 1, 0);
    for (int64_t r = 0; r < N; ++r) P->level_ptr[hl[r] + 1]++;
    for (int l = 0; l < n_levels; ++l) P->level_ptr[l + 1] += P->level_ptr[l];
    if (getenv("MFB_ILU_VERBOSE")) {
        fprintf(stderr, "[mfb ilu] %d levels, rows per level:", n_levels);
        for (int l = 0; l < n_levels; ++l) fprintf(stderr, " %d", P->level_ptr[l + 1] - P->level_ptr[l]);
        fprintf(stderr, "\n");
    }
    MFB_CUDA(P->lperm.alloc(U)); MFB_CUDA(P->nlow.alloc(N));
    LAUNCH(k_lower_order, nblk(N), TPB, ctx->nodeptr.p, ctx->nodecol.p, P->kind.p, P->pos.p, N, P->lperm.p, P->nlow.p);
    // packed layout of the sweeps: rows in level order, lower / upper entries contiguous
    MFB_CUDA(P->Lptr.alloc(N + 1)); MFB_CUDA(P->Uptr.alloc(N + 1));
    MFB_CUDA(cudaMemsetAsync(P->Lptr.p, 0, (N + 1) * sizeof(int), ctx->stream));
    MFB_CUDA(cudaMemsetAsync(P->Uptr.p, 0, (N + 1) * sizeof(int), ctx->stream));
    LAUNCH(k_pack_counts, nblk(N), TPB, P->rows.p, ctx->nodeptr.p, P->nlow.p, N, P->Lptr.p, P->Uptr.p);
    {
        thrust::device_ptr<int> a(P->Lptr.p), b(P->Uptr.p);
        thrust::exclusive_scan(pol, a, a + (N + 1), a);
        thrust::exclusive_scan(pol, b, b + (N + 1), b);
        int ul = 0, uu = 0;
        MFB_CUDA(cudaMemcpyAsync(&ul, P->Lptr.p + N, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaMemcpyAsync(&uu, P->Uptr.p + N, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
        P->UL = ul; P->UU = uu;
    }
    MFB_CUDA(P->Lcol.alloc(P->UL + MFB_STREAM_PAD)); MFB_CUDA(P->Lsrc.alloc(P->UL > 0 ? P->UL : 1));
    MFB_CUDA(P->Ucol.alloc(P->UU + MFB_STREAM_PAD)); MFB_CUDA(P->Usrc.alloc(P->UU > 0 ? P->UU : 1));
    LAUNCH(k_pack_index, nblk(N), TPB, P->rows.p, ctx->nodeptr.p, ctx->nodecol.p, P->kind.p, P->lperm.p, P->nlow.p, P->Lptr.p, P->Uptr.p, N,
           P->Lcol.p, P->Lsrc.p, P->Ucol.p, P->Usrc.p);
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

template <int NV>
static int ilu_factor_nv(mfb_ctx* ctx, IluPlan* P) {
    for (size_t l = 0; l + 1 < P->level_ptr.size(); ++l) {
        const int off = P->level_ptr[l], cnt = P->level_ptr[l + 1] - off;
        if (cnt > 0)
            LAUNCH((k_ilu_factor<NV>), nblk((int64_t)cnt * 32), TPB, P->rows.p + off, cnt, ctx->nodeptr.p, ctx->nodecol.p, P->kind.p, P->lperm.p,
                   P->nlow.p, P->F.p, P->dinv.p);
    }
    return MFB_OK;
}
static bool ilu_stream_sweeps() {
    static const bool on = [] { const char* e = getenv("MFB_ILU_SWEEP"); return !(e && e[0] == 'r'); }();   // "row": one warp per row
    return on;
}
// The packed factors are kept in FP32 (accumulation in FP64): the incomplete factorisation is an approximation of A to a few
// per cent, rounding its entries to 2^-24 does not change what it does to the spectrum, and the sweeps -- pure streaming of the
// factors -- move half the bytes. The preconditioner stays one fixed linear operator, so the Krylov methods converge to the same
// tolerance on the TRUE residual (which never sees the factors). MFB_ILU_FP64=1 keeps doubles.
static bool ilu_unpacked() {
    static const bool on = [] { const char* e = getenv("MFB_ILU_UNPACKED"); return e && e[0] == '1'; }();   // round-2 first version (A/B)
    return on;
}
static bool ilu_use_f32(mfb_ctx* ctx) {
    static const bool f64 = [] { const char* e = getenv("MFB_ILU_FP64"); return e && e[0] == '1'; }();
    return !f64 && !ilu_unpacked() && ilu_stream_sweeps() && ctx->n_var <= 4;
}
template <typename VT>
static void pack_upper_scaled(mfb_ctx* ctx, IluPlan* P, VT* out) {
    const unsigned grid = nblk(P->N * 32);
    switch (ctx->n_var) {
        case 1: LAUNCH((k_pack_upper_scaled<1, VT>), grid, TPB, P->rows.p, P->Uptr.p, P->Usrc.p, P->N, P->F.p, P->dinv.p, out); break;
        case 2: LAUNCH((k_pack_upper_scaled<2, VT>), grid, TPB, P->rows.p, P->Uptr.p, P->Usrc.p, P->N, P->F.p, P->dinv.p, out); break;
        case 3: LAUNCH((k_pack_upper_scaled<3, VT>), grid, TPB, P->rows.p, P->Uptr.p, P->Usrc.p, P->N, P->F.p, P->dinv.p, out); break;
        default: LAUNCH((k_pack_upper_scaled<4, VT>), grid, TPB, P->rows.p, P->Uptr.p, P->Usrc.p, P->N, P->F.p, P->dinv.p, out); break;
    }
}

template <int NV>
static int ilu_apply_nv(mfb_ctx* ctx, IluPlan* P, double* v) {
    const int nl = (int)P->level_ptr.size() - 1;
    const bool unpacked = ilu_unpacked();
    const bool stream_kernel = P->scaledU;
    for (int l = 1; l < nl; ++l) {                               // level 0 has no lower entries
        const int off = P->level_ptr[l], cnt = P->level_ptr[l + 1] - off;
        if (cnt <= 0) continue;
        if (unpacked)
            LAUNCH((k_ilu_sweep<NV, true>), nblk((int64_t)cnt * 32), TPB, P->rows.p + off, cnt, ctx->nodeptr.p, ctx->nodecol.p, P->kind.p, P->F.p,
                   P->dinv.p, v);
        else if (!(stream_kernel && mfb_sweep_level_mr(ctx, false, P->Lptr.p, P->Lcol.p, P->f32 ? (const void*)P->Lval32.p : (const void*)P->Lval.p,
                                                       P->f32, P->rows.p, P->dinv.p, v, off, off + cnt)))
            LAUNCH((k_ilu_sweep_packed<NV, true>), nblk((int64_t)cnt * 32), TPB, P->rows.p, off, cnt, P->Lptr.p, P->Lcol.p, P->Lval.p, P->dinv.p, v);
    }
    for (int l = nl - 1; l >= 0; --l) {
        const int off = P->level_ptr[l], cnt = P->level_ptr[l + 1] - off;
        if (cnt <= 0) continue;
        if (unpacked)
            LAUNCH((k_ilu_sweep<NV, false>), nblk((int64_t)cnt * 32), TPB, P->rows.p + off, cnt, ctx->nodeptr.p, ctx->nodecol.p, P->kind.p, P->F.p,
                   P->dinv.p, v);
        else if (!(stream_kernel && mfb_sweep_level_mr(ctx, true, P->Uptr.p, P->Ucol.p, P->f32 ? (const void*)P->Uval32.p : (const void*)P->Uval.p,
                                                       P->f32, P->rows.p, P->dinv.p, v, off, off + cnt)))
            LAUNCH((k_ilu_sweep_packed<NV, false>), nblk((int64_t)cnt * 32), TPB, P->rows.p, off, cnt, P->Uptr.p, P->Ucol.p, P->Uval.p, P->dinv.p, v);
    }
    return MFB_OK;
}

// factorise A ([U][nv*nv], the matrix of the running solve) into the context's ILU storage
int mfb_ilu_factor(mfb_ctx* ctx, const double* A, int* n_levels) {
    MFB_REQUIRE(!mfb_is_distributed(ctx), MFB_ERR_ARG, "Pl_ILU needs assembled rows: not available on a partitioned mesh");
    MFB_REQUIRE(ctx->n_var >= 1 && ctx->n_var <= 4, MFB_ERR_ARG, "Pl_ILU: n_var <= 4");
    IluPlan* P = nullptr;
    MFB_TRY(ilu_setup(ctx, P));
    const int B = ctx->n_var * ctx->n_var;
    MFB_CUDA(P->F.alloc((size_t)ctx->U * B));
    MFB_CUDA(P->dinv.alloc((size_t)ctx->N * B));
    MFB_CUDA(cudaMemcpyAsync(P->F.p, A, (size_t)ctx->U * B * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    switch (ctx->n_var) {
        case 1: MFB_TRY(ilu_factor_nv<1>(ctx, P)); break;
        case 2: MFB_TRY(ilu_factor_nv<2>(ctx, P)); break;
        case 3: MFB_TRY(ilu_factor_nv<3>(ctx, P)); break;
        default: MFB_TRY(ilu_factor_nv<4>(ctx, P)); break;
    }
    P->f32 = ilu_use_f32(ctx);
    P->scaledU = ilu_stream_sweeps() && ctx->n_var <= 4 && !ilu_unpacked();
    if (P->f32) {
        MFB_CUDA(P->Lval32.alloc((size_t)(P->UL + MFB_STREAM_PAD) * B)); MFB_CUDA(P->Uval32.alloc((size_t)(P->UU + MFB_STREAM_PAD) * B));
        if (P->UL > 0) LAUNCH(k_pack_values<float>, nblk(P->UL * B), TPB, P->Lsrc.p, P->UL, B, P->F.p, P->Lval32.p);
        if (P->scaledU) pack_upper_scaled<float>(ctx, P, P->Uval32.p);
        else if (P->UU > 0) LAUNCH(k_pack_values<float>, nblk(P->UU * B), TPB, P->Usrc.p, P->UU, B, P->F.p, P->Uval32.p);
    } else {
        MFB_CUDA(P->Lval.alloc((size_t)(P->UL + MFB_STREAM_PAD) * B)); MFB_CUDA(P->Uval.alloc((size_t)(P->UU + MFB_STREAM_PAD) * B));
        if (P->UL > 0) LAUNCH(k_pack_values<double>, nblk(P->UL * B), TPB, P->Lsrc.p, P->UL, B, P->F.p, P->Lval.p);
        if (P->scaledU) pack_upper_scaled<double>(ctx, P, P->Uval.p);
        else if (P->UU > 0) LAUNCH(k_pack_values<double>, nblk(P->UU * B), TPB, P->Usrc.p, P->UU, B, P->F.p, P->Uval.p);
    }
    MFB_CUDA(cudaGetLastError());
    P->factored = true;
    if (n_levels) *n_levels = (int)P->level_ptr.size() - 1;
    return MFB_OK;
}

// v <- U^-1 L^-1 v (the Pl(b) of _Pl_ILU, 02_Preconditioner.jl:189-193)
int mfb_ilu_apply(mfb_ctx* ctx, double* v) {
    auto it = ilu_table().find(ctx);
    MFB_REQUIRE(it != ilu_table().end() && it->second->factored, MFB_ERR_STATE, "Pl_ILU applied before it was factorised");
    IluPlan* P = it->second;
    ProfScope ps(ctx, MFB_T_PRECOND);
    switch (ctx->n_var) {
        case 1: MFB_TRY(ilu_apply_nv<1>(ctx, P, v)); break;
        case 2: MFB_TRY(ilu_apply_nv<2>(ctx, P, v)); break;
        case 3: MFB_TRY(ilu_apply_nv<3>(ctx, P, v)); break;
        default: MFB_TRY(ilu_apply_nv<4>(ctx, P, v)); break;
    }
    MFB_CUDA(cudaGetLastError());
    return MFB_OK;
}

// Test / diagnosis entry: factorises K_total as it is (no Jacobi scaling) and returns max |(L U - A)_ij| over the pattern divided
// by max |A_ij| (the defining property of an incomplete factorisation with zero fill), the number of dependency levels, and
// applies the preconditioner to `v_inout` (reference layout, may be NULL).
extern "C" int mfb_ilu_selftest(mfb_ctx* ctx, double* rel_defect, int32_t* n_levels, double* v_inout, int64_t n) {
    if (!ctx) return MFB_ERR_ARG;
    MFB_REQUIRE(ctx->U > 0 && ctx->K_total.p, MFB_ERR_STATE, "mfb_ilu_selftest: pattern/matrix not built");
    MFB_CUDA(cudaSetDevice(ctx->device));
    int nl = 0;
    MFB_TRY(mfb_ilu_factor(ctx, ctx->K_total.p, &nl));
    IluPlan* P = ilu_table()[ctx];
    DevBuf<double> out;
    MFB_CUDA(out.alloc(2));
    MFB_CUDA(cudaMemsetAsync(out.p, 0, 2 * sizeof(double), ctx->stream));
    const unsigned grid = nblk(ctx->N * 32);
    switch (ctx->n_var) {
        case 1: LAUNCH((k_ilu_defect<1>), grid, TPB, ctx->nodeptr.p, ctx->nodecol.p, P->kind.p, P->pos.p, P->F.p, P->dinv.p, ctx->K_total.p, ctx->N, out.p); break;
        case 2: LAUNCH((k_ilu_defect<2>), grid, TPB, ctx->nodeptr.p, ctx->nodecol.p, P->kind.p, P->pos.p, P->F.p, P->dinv.p, ctx->K_total.p, ctx->N, out.p); break;
        case 3: LAUNCH((k_ilu_defect<3>), grid, TPB, ctx->nodeptr.p, ctx->nodecol.p, P->kind.p, P->pos.p, P->F.p, P->dinv.p, ctx->K_total.p, ctx->N, out.p); break;
        default: LAUNCH((k_ilu_defect<4>), grid, TPB, ctx->nodeptr.p, ctx->nodecol.p, P->kind.p, P->pos.p, P->F.p, P->dinv.p, ctx->K_total.p, ctx->N, out.p); break;
    }
    double h[2] = {0, 0};
    MFB_CUDA(cudaMemcpyAsync(h, out.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (rel_defect) *rel_defect = h[1] > 0 ? h[0] / h[1] : h[0];
    if (n_levels) *n_levels = nl;
    if (v_inout) {
        MFB_REQUIRE(n == ctx->N * ctx->n_var, MFB_ERR_ARG, "mfb_ilu_selftest: wrong vector length");
        DevBuf<double> vr, vi;
        MFB_CUDA(vr.alloc(n)); MFB_CUDA(vi.alloc(n));
        MFB_TRY(mfb_stage_in(ctx, v_inout, n * sizeof(double), vr.p));
        MFB_TRY(mfb_to_internal(ctx, vr.p, vi.p, 1));
        MFB_TRY(mfb_ilu_apply(ctx, vi.p));
        MFB_TRY(mfb_to_reference(ctx, vi.p, vr.p, 1));
        MFB_TRY(mfb_stage_out(ctx, vr.p, n * sizeof(double), v_inout));
        MFB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return MFB_OK;
}
