// spmv_exp2.cu -- SpMV variants on the REAL 88^3 hex20 node graph (dumped with MFB_DUMP_PATTERN) to study the x gather.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
constexpr int NV = 3, B = 9, EPW = 3, ACTIVE = 27;

template <int MODE> __device__ __forceinline__ double ldv(const double* p) { return MODE == 0 ? __ldcs(p) : MODE == 1 ? __ldcg(p) : __ldg(p); }
template <int MODE> __device__ __forceinline__ int ldi(const int* p) { return MODE == 0 ? __ldg(p) : MODE == 1 ? __ldcg(p) : __ldcs(p); }

// ROWS_PER_CTA == 0: one warp per row, rows interleaved over the grid; otherwise each CTA owns a contiguous chunk of rows
template <int UNR, int VMODE, int CMODE, int ROWS_PER_CTA, bool SHFL>
__global__ void __launch_bounds__(256) k_v(const int* __restrict__ nodeptr, const int* __restrict__ nodecol,
                                            const double* __restrict__ K, const double* __restrict__ x, double* __restrict__ y, int64_t N) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int le = lane / B, ik = lane - le * B, i = ik / NV, k = ik - i * NV;
    const double* xk = x + k;
    int64_t row0, rstride, rend;
    if (ROWS_PER_CTA == 0) { row0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; rstride = 1; rend = min((int64_t)row0 + 1, N); }
    else { row0 = blockIdx.x * (int64_t)ROWS_PER_CTA + warp; rstride = 8; rend = min((blockIdx.x + 1) * (int64_t)ROWS_PER_CTA, N); }
    for (int64_t row = row0; row < rend; row += rstride) {
        const int s = nodeptr[row];
        const int deg = (lane < ACTIVE) ? nodeptr[row + 1] - s : 0;
        const double* Kp = K + (size_t)s * B + lane;
        const int* Cp = nodecol + s + le;
        double a[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) a[u] = 0.0;
        int e = le;
        for (; e + (UNR - 1) * EPW < deg; e += UNR * EPW) {
            double v[UNR]; int c[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) { v[u] = ldv<VMODE>(Kp + u * ACTIVE); c[u] = ldi<CMODE>(Cp + u * EPW); }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                double xv;
                if (SHFL) {   // only the i == 0 lanes gather, the other two block rows get it by shuffle
                    double g = (i == 0) ? __ldg(xk + (size_t)c[u] * NV) : 0.0;
                    xv = __shfl_sync(0xffffffffu, g, le * B + k);
                } else xv = __ldg(xk + (size_t)c[u] * NV);
                a[u] += v[u] * xv;
            }
            Kp += UNR * ACTIVE; Cp += UNR * EPW;
        }
        for (; e < deg; e += EPW) { a[0] += ldv<VMODE>(Kp) * __ldg(xk + (size_t)ldi<CMODE>(Cp) * NV); Kp += ACTIVE; Cp += EPW; }
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < UNR; ++u) acc += a[u];
        double t = acc;
        for (int d = 1; d < NV; ++d) t += __shfl_down_sync(0xffffffffu, acc, d);
        double r = t;
        for (int d = 1; d < EPW; ++d) r += __shfl_down_sync(0xffffffffu, t, d * B);
        if (le == 0 && k == 0 && lane < ACTIVE) y[(size_t)row * NV + i] = r;
    }
}

// P: software-pipelined -- the streaming loads (values, column ids) of batch j+1 are issued BEFORE the dependent x gathers of
// batch j are consumed, and (chunked mode) the next row's pointers are fetched one row ahead.
template <int UNR, int ROWS_PER_CTA>
__global__ void __launch_bounds__(256) k_p(const int* __restrict__ nodeptr, const int* __restrict__ nodecol,
                                            const double* __restrict__ K, const double* __restrict__ x, double* __restrict__ y, int64_t N) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int le = lane / B, ik = lane - le * B, i = ik / NV, k = ik - i * NV;
    const double* xk = x + k;
    int64_t row0, rstride, rend;
    if (ROWS_PER_CTA == 0) { row0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; rstride = 1; rend = min((int64_t)row0 + 1, N); }
    else { row0 = blockIdx.x * (int64_t)ROWS_PER_CTA + warp; rstride = 8; rend = min((blockIdx.x + 1) * (int64_t)ROWS_PER_CTA, N); }
    if (row0 >= rend) return;
    int ps = nodeptr[row0], pt = nodeptr[row0 + 1];
    for (int64_t row = row0; row < rend; row += rstride) {
        const int s = ps, deg = (lane < ACTIVE) ? pt - ps : 0;
        const int64_t nrow = row + rstride;
        if (nrow < rend) { ps = __ldg(nodeptr + nrow); pt = __ldg(nodeptr + nrow + 1); }
        const double* Kp = K + (size_t)s * B + lane;
        const int* Cp = nodecol + s + le;
        double a[UNR], v[UNR]; int c[UNR];
        int e = le;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            a[u] = 0.0;
            const bool ok = e + u * EPW < deg;
            v[u] = ok ? __ldcs(Kp + u * ACTIVE) : 0.0;
            c[u] = ok ? __ldg(Cp + u * EPW) : 0;
        }
        while (e < deg) {
            double vn[UNR]; int cn[UNR];
            e += UNR * EPW; Kp += UNR * ACTIVE; Cp += UNR * EPW;
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
        for (int d = 1; d < NV; ++d) t += __shfl_down_sync(0xffffffffu, acc, d);
        double r = t;
        for (int d = 1; d < EPW; ++d) r += __shfl_down_sync(0xffffffffu, t, d * B);
        if (le == 0 && k == 0 && lane < ACTIVE) y[(size_t)row * NV + i] = r;
    }
}

// PP: like P but two batches of streaming loads in flight ahead of the gathers
template <int UNR, int ROWS_PER_CTA>
__global__ void __launch_bounds__(256) k_pp(const int* __restrict__ nodeptr, const int* __restrict__ nodecol,
                                             const double* __restrict__ K, const double* __restrict__ x, double* __restrict__ y, int64_t N) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int le = lane / B, ik = lane - le * B, i = ik / NV, k = ik - i * NV;
    const double* xk = x + k;
    int64_t row0 = blockIdx.x * (int64_t)ROWS_PER_CTA + warp, rstride = 8, rend = min((blockIdx.x + 1) * (int64_t)ROWS_PER_CTA, N);
    if (row0 >= rend) return;
    int ps = nodeptr[row0], pt = nodeptr[row0 + 1];
    for (int64_t row = row0; row < rend; row += rstride) {
        const int s = ps, deg = (lane < ACTIVE) ? pt - ps : 0;
        const int64_t nrow = row + rstride;
        if (nrow < rend) { ps = __ldg(nodeptr + nrow); pt = __ldg(nodeptr + nrow + 1); }
        const double* Kp = K + (size_t)s * B + lane;
        const int* Cp = nodecol + s + le;
        double a[UNR], v0[UNR], v1[UNR]; int c0[UNR], c1[UNR];
        int e = le;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            a[u] = 0.0;
            bool ok = e + u * EPW < deg;
            v0[u] = ok ? __ldcs(Kp + u * ACTIVE) : 0.0; c0[u] = ok ? __ldg(Cp + u * EPW) : 0;
            ok = e + (UNR + u) * EPW < deg;
            v1[u] = ok ? __ldcs(Kp + (UNR + u) * ACTIVE) : 0.0; c1[u] = ok ? __ldg(Cp + (UNR + u) * EPW) : 0;
        }
        while (e < deg) {
            double vn[UNR]; int cn[UNR];
            e += UNR * EPW; Kp += UNR * ACTIVE; Cp += UNR * EPW;
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const bool ok = e + (UNR + u) * EPW < deg;
                vn[u] = ok ? __ldcs(Kp + (UNR + u) * ACTIVE) : 0.0;
                cn[u] = ok ? __ldg(Cp + (UNR + u) * EPW) : 0;
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) a[u] += v0[u] * __ldg(xk + (size_t)c0[u] * NV);
#pragma unroll
            for (int u = 0; u < UNR; ++u) { v0[u] = v1[u]; c0[u] = c1[u]; v1[u] = vn[u]; c1[u] = cn[u]; }
        }
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < UNR; ++u) acc += a[u];
        double t = acc;
        for (int d = 1; d < NV; ++d) t += __shfl_down_sync(0xffffffffu, acc, d);
        double r = t;
        for (int d = 1; d < EPW; ++d) r += __shfl_down_sync(0xffffffffu, t, d * B);
        if (le == 0 && k == 0 && lane < ACTIVE) y[(size_t)row * NV + i] = r;
    }
}

int main(int argc, char** argv) {
    FILE* f = fopen(argc > 1 ? argv[1] : "/tmp/pattern.bin", "rb");
    if (!f) { printf("no pattern file\n"); return 1; }
    int64_t hdr[2]; if (fread(hdr, 8, 2, f) != 2) return 1;
    const int64_t N = hdr[0], U = hdr[1];
    std::vector<int> ptr(N + 1), col(U);
    if (fread(ptr.data(), 4, N + 1, f) != (size_t)N + 1 || fread(col.data(), 4, U, f) != (size_t)U) return 1;
    fclose(f);
    printf("N=%ld U=%ld values %.2f GB avg deg %.1f\n", (long)N, (long)U, U * 72.0 / 1e9, (double)U / N);
    {   // locality statistics of the node graph in the library's internal numbering
        double sumspan = 0; for (int64_t r = 0; r < N; r += 97) sumspan += col[ptr[r + 1] - 1] - col[ptr[r]];
        printf("mean column span of a row: %.0f nodes\n", sumspan / ((N + 96) / 97));
    }
    int *dptr, *dcol; double *dK, *dx, *dy, *dy2;
    CK(cudaMalloc(&dptr, (N + 1) * 4)); CK(cudaMalloc(&dcol, U * 4)); CK(cudaMalloc(&dK, U * 72)); CK(cudaMalloc(&dx, N * 24)); CK(cudaMalloc(&dy, N * 24)); CK(cudaMalloc(&dy2, N * 24));
    CK(cudaMemcpy(dptr, ptr.data(), (N + 1) * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dcol, col.data(), U * 4, cudaMemcpyHostToDevice));
    uint64_t rng = 88172645463325252ull;
    auto rnd = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    {
        std::vector<double> hx(N * 3); for (auto& v : hx) v = (double)(rnd() % 1000) / 1000.0;
        CK(cudaMemcpy(dx, hx.data(), N * 24, cudaMemcpyHostToDevice));
        std::vector<double> blk(1 << 20); for (auto& v : blk) v = (double)(rnd() % 1000) / 1000.0 - 0.5;
        for (size_t off = 0; off < (size_t)U * 9; off += blk.size()) CK(cudaMemcpy(dK + off, blk.data(), std::min(blk.size(), (size_t)U * 9 - off) * 8, cudaMemcpyHostToDevice));
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = U * 76.0 + N * 52.0;
    bool first = true;
    auto timeit = [&](const char* name, auto launch) {
        for (int w = 0; w < 2; ++w) launch();
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        const int reps = 8;
        for (int w = 0; w < reps; ++w) launch();
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
        double md = 0;
        if (!first) {
            std::vector<double> h1(N * 3), h2(N * 3);
            CK(cudaMemcpy(h1.data(), dy, N * 24, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2.data(), dy2, N * 24, cudaMemcpyDeviceToHost));
            for (size_t q = 0; q < h1.size(); ++q) md = std::max(md, std::abs(h1[q] - h2[q]));
        }
        first = false;
        printf("%-44s %8.3f ms  %7.1f GB/s actual  maxdiff %.1e\n", name, ms, bytes / ms / 1e6, md);
    };
    unsigned grid = (unsigned)((N * 32 + 255) / 256);
#define RUNW(name, U_, VM, CM, SH) timeit(name, [&] { k_v<U_, VM, CM, 0, SH><<<grid, 256>>>(dptr, dcol, dK, dx, first ? dy : dy2, N); })
#define RUNC(name, U_, VM, CM, R, SH) timeit(name, [&] { k_v<U_, VM, CM, R, SH><<<(unsigned)((N + R - 1) / R), 256>>>(dptr, dcol, dK, dx, dy2, N); })
    RUNW("A4  ldcs/ldg unroll4 (round-1 kernel)", 4, 0, 0, false);
#define RUNP(name, U_, R) timeit(name, [&] { k_p<U_, R><<<(unsigned)(R == 0 ? grid : (N + (R ? R : 1) - 1) / (R ? R : 1)), 256>>>(dptr, dcol, dK, dx, dy2, N); })
    RUNP("P3/64  pipelined unroll3, chunk 64", 3, 64);
    RUNP("P3/32  pipelined unroll3, chunk 32", 3, 32);
    RUNP("P3/128 pipelined unroll3, chunk 128", 3, 128);
    RUNP("P3/256 pipelined unroll3, chunk 256", 3, 256);
    RUNP("P5/64  pipelined unroll5, chunk 64", 5, 64);
    RUNP("P6/64  pipelined unroll6, chunk 64", 6, 64);
    RUNP("P9/64  pipelined unroll9, chunk 64", 9, 64);
#define RUNPP(name, U_, R) timeit(name, [&] { k_pp<U_, R><<<(unsigned)((N + R - 1) / R), 256>>>(dptr, dcol, dK, dx, dy2, N); })
    RUNPP("PP2/64 depth-2 unroll2, chunk 64", 2, 64);
    RUNPP("PP3/64 depth-2 unroll3, chunk 64", 3, 64);
    RUNPP("PP3/128 depth-2 unroll3, chunk 128", 3, 128);
    RUNPP("PP1/64 depth-2 unroll1, chunk 64", 1, 64);
    return 0;
}
