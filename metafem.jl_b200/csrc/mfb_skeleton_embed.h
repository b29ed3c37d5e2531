// generated from mfb_skeleton.cuh by build.py -- do not edit
static const char mfb_skeleton_src[] = R"MFBSKEL(// mfb_skeleton.cuh -- fused element-assembly skeleton for sm_100a (FP64, no tensor cores).
//
// This header is compiled twice: by nvcc into libmetafem_b200.so (host-visible MfbArgs layout)
// and by NVRTC together with the CUDA C the term emitter generates for one weak form
// (mfb_kernel_compile). It replaces, in ONE kernel per block of the generated updater, what the
// reference does with ~3 launches per weak-form term:
//   _Var_Basic  (reference src/solver/06_FEM_Kernel.jl:1-13)   -> interpolation of words at q-points
//   `vals = @. expr * weights`  (src/solver/05_CodeGenerator.jl:75-76,110-111,136-137) -> Form::point
//   _Res_Basic  (06_FEM_Kernel.jl:65-79)                        -> residual contraction + scatter
//   _Kval_Basic (06_FEM_Kernel.jl:28-45)                        -> tangent contraction + scatter
// and it evaluates the geometry of update_Mesh on the fly (Jacobian, inverse, detJ, physical
// gradients, boundary tangents/normals; src/mesh/unstructured_mesh/4_Update_Integrator.jl:2-75,
// 90-154,173-227) from node coordinates and the reference-element tables.
//
// Data layout in HBM (all internal numbering = the library's locality permutation of nodes):
//   conn   [n_el][NA]      int32  internal node id of local node a
//   xyz    [N][3]          double node coordinates (AoS, one 24 B gather per node)
//   u      [L+1][N][NV]    double x_star, node-major interleaved (one contiguous NV-gather per node)
//   emap   [n_el][NA*NA]   int32  index of node pair (a,b) in the node-graph CSR
//   Kval   [U][NV*NV]      double block values, row-major inside the block (BSR with NV x NV blocks)
//   res    [N][NV]         double residual
//   cpv[i] [N]             double CONTROLPOINT_VAR fields
// Tables (small, L1/L2 resident): ref [n_tab][4][NQ][NA] (slot 0:N, 1..3: d/dX), wq [n_tab][NQ],
//   btan [n_tab][NQ][3][2] for boundary blocks.
#pragma once

#ifndef MFB_MAX_CPV
#define MFB_MAX_CPV 24
#define MFB_MAX_GLOBALS 16
#define MFB_MAX_KPARAMS 4
#endif

struct MfbArgs {
    // mesh
    const int* conn;
    const double* xyz;
    const int* emap;
    long long n_items;          // elements (domain) or facets of the group (boundary)
    long long N;                // nodes
    // boundary: facet -> host element / local face; domain: nullptr
    const int* item_elem;       // [n_items] 0-based element id
    const int* item_face;       // [n_items] 0-based local face id (table index)
    // tables
    const double* ref;          // [n_tab][4][NQ][NA]
    const double* wq;           // [n_tab][NQ]
    const double* btan;         // [n_tab][NQ][3][2]
    // state
    const double* u;            // x_star internal layout
    double* res;
    double* Kval;
    // per-call scalars
    double Kp[MFB_MAX_KPARAMS];
    double glob[MFB_MAX_GLOBALS];
    const double* cpv[MFB_MAX_CPV];
};

#ifdef __CUDACC__

namespace mfb {

__device__ __forceinline__ void red_add(double* p, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// 3x3 inverse exactly as inv_Jac_3D spells it (4_Update_Integrator.jl:90-121): J[i][X] = dx_i/dX
__device__ __forceinline__ double inv3(const double (&J)[3][3], double (&I)[3][3]) {
    double det = J[0][0] * J[1][1] * J[2][2] - J[0][0] * J[1][2] * J[2][1] - J[0][1] * J[1][0] * J[2][2] +
                 J[0][1] * J[1][2] * J[2][0] + J[0][2] * J[1][0] * J[2][1] - J[0][2] * J[1][1] * J[2][0];
    I[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det;
    I[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
    I[0][2] = (J[0][1] * J[1][2] - J[1][1] * J[0][2]) / det;
    I[1][0] = (J[1][2] * J[2][0] - J[2][2] * J[1][0]) / det;
    I[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
    I[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
    I[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det;
    I[2][1] = (J[0][1] * J[2][0] - J[2][1] * J[0][0]) / det;
    I[2][2] = (J[0][0] * J[1][1] - J[1][0] * J[0][1]) / det;
    return det;
}

// Form contract (emitted code):
//   static constexpr int NV, NA, NQ, L1 (= max_time_level + 1), BOUNDARY (0/1), LINEAR (0/1),
//                        NW (inner words), NCW (cp words), NC (cp fields), NT (K terms), HAS_RES, HAS_K, TPB;
//   template <class S> __device__ static void words(const S& s, int q, double* w, double* c);
//        // w[k] = interpolation of inner word k, c[k] = of CONTROLPOINT_VAR word k  (_Var_Basic)
//   __device__ static void point(const double* w, const double* c, const double* nrm, const MfbArgs& A,
//                                double* R /*[NV*4], zeroed*/, double* D /*[NT]*/);   // NOT yet weighted
//   __device__ static void kacc(double* acc /*[NV*NV]*/, const double* Ga /*[4]*/, const double* Gb /*[4]*/,
//                               const double* Dq /*[NT]*/);
//
// One thread block works on one item at a time (grid-stride over items); TPB threads.
// Shared memory per item: G[NQ][NA][4] (physical shape values/gradients), Dq[NQ][NT], Rq[NQ][NV*4],
// node data xe[NA][3], ue[L1][NA][NV], ce[NC][NA], node ids.
template <int NA>
__device__ __forceinline__ double interp(const double* G4 /*stride 4*/, const double* u, int ustride) {
    double s = 0.0;
#pragma unroll 4
    for (int a = 0; a < NA; ++a) s += G4[4 * a] * u[a * ustride];
    return s;
}

template <class F>
struct Smem {
    double G[F::NQ][F::NA][4];
    double D[F::NQ][F::NT > 0 ? F::NT : 1];
    double R[F::NQ][F::NV * 4];
    double xe[F::NA][3];
    double ue[F::L1][F::NA][F::NV];
    double ce[F::NC > 0 ? F::NC : 1][F::NA];
    int node[F::NA];
};

template <class F>
__device__ __forceinline__ void assemble(const MfbArgs& A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<F>& S = *reinterpret_cast<Smem<F>*>(smem_raw);
    const int tid = threadIdx.x;
    constexpr int NA = F::NA, NQ = F::NQ, NV = F::NV;

    for (long long item = blockIdx.x; item < A.n_items; item += gridDim.x) {
        long long e = item;
        int tab = 0;
        if (F::BOUNDARY) {
            e = A.item_elem[item];
            tab = A.item_face[item];
        }
        const double* ref = A.ref + (size_t)tab * 4 * NQ * NA;
        __syncthreads();  // previous item fully consumed
        // ---- phase A: gather node data ----------------------------------------------------
        for (int a = tid; a < NA; a += F::TPB) {
            int g = A.conn[e * NA + a];
            S.node[a] = g;
            S.xe[a][0] = A.xyz[3 * (size_t)g + 0];
            S.xe[a][1] = A.xyz[3 * (size_t)g + 1];
            S.xe[a][2] = A.xyz[3 * (size_t)g + 2];
        }
        if (!F::LINEAR) {
            for (int i = tid; i < F::L1 * NA * NV; i += F::TPB) {
                int v = i % NV, a = (i / NV) % NA, l = i / (NV * NA);
                int g = A.conn[e * NA + a];
                S.ue[l][a][v] = A.u[((size_t)l * A.N + g) * NV + v];
            }
        }
        for (int i = tid; i < F::NC * NA; i += F::TPB) {
            int a = i % NA, c = i / NA;
            S.ce[c][a] = A.cpv[c][A.conn[e * NA + a]];
        }
        __syncthreads();
        // ---- phase B: one thread per quadrature point -------------------------------------
        for (int q = tid; q < NQ; q += F::TPB) {
            double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            for (int a = 0; a < NA; ++a) {
                double d0 = ref[(1 * NQ + q) * NA + a], d1 = ref[(2 * NQ + q) * NA + a], d2 = ref[(3 * NQ + q) * NA + a];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    double x = S.xe[a][i];
                    J[i][0] += d0 * x;
                    J[i][1] += d1 * x;
                    J[i][2] += d2 * x;
                }
            }
            double I[3][3];
            double det = inv3(J, I);
            double wgt;
            double nrm[3] = {0, 0, 0};
            if (F::BOUNDARY) {
                const double* bt = A.btan + ((size_t)tab * NQ + q) * 6;  // [3][2]
                double t[3][2];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int k = 0; k < 2; ++k) t[i][k] = J[i][0] * bt[0 * 2 + k] + J[i][1] * bt[1 * 2 + k] + J[i][2] * bt[2 * 2 + k];
                double r0 = t[1][0] * t[2][1] - t[2][0] * t[1][1];
                double r1 = -t[0][0] * t[2][1] + t[2][0] * t[0][1];
                double r2 = t[0][0] * t[1][1] - t[1][0] * t[0][1];
                double ld = sqrt(r0 * r0 + r1 * r1 + r2 * r2);
                nrm[0] = r0 / ld; nrm[1] = r1 / ld; nrm[2] = r2 / ld;
                wgt = A.wq[tab * NQ + q] * ld;
            } else {
                wgt = A.wq[q] * det;
            }
            for (int a = 0; a < NA; ++a) {
                double d0 = ref[(1 * NQ + q) * NA + a], d1 = ref[(2 * NQ + q) * NA + a], d2 = ref[(3 * NQ + q) * NA + a];
                S.G[q][a][0] = ref[(0 * NQ + q) * NA + a];
#pragma unroll
                for (int s = 0; s < 3; ++s) S.G[q][a][1 + s] = (d0 * I[0][s] + d1 * I[1][s]) + d2 * I[2][s];
            }
            double w[F::NW > 0 ? F::NW : 1], c[F::NCW > 0 ? F::NCW : 1];
            F::words(S, q, w, c);
            double R[NV * 4], D[F::NT > 0 ? F::NT : 1];
#pragma unroll
            for (int k = 0; k < NV * 4; ++k) R[k] = 0.0;
            F::point(w, c, nrm, A, R, D);
#pragma unroll
            for (int k = 0; k < NV * 4; ++k) S.R[q][k] = R[k] * wgt;
#pragma unroll
            for (int k = 0; k < F::NT; ++k) S.D[q][k] = D[k] * wgt;
        }
        __syncthreads();
        // ---- phase C1: residual ------------------------------------------------------------
        if (F::HAS_RES) {
            for (int i = tid; i < NA * NV; i += F::TPB) {
                int v = i % NV, a = i / NV;
                double s = 0.0;
                for (int q = 0; q < NQ; ++q) {
#pragma unroll
                    for (int sl = 0; sl < 4; ++sl) s += S.G[q][a][sl] * S.R[q][v * 4 + sl];
                }
                if (s != 0.0) red_add(A.res + (size_t)S.node[a] * NV + v, s);
            }
        }
        // ---- phase C2: tangent, one (a,b) node pair per thread -----------------------------
        if (F::HAS_K) {
            for (int p = tid; p < NA * NA; p += F::TPB) {
                int a = p / NA, b = p % NA;
                double acc[NV * NV];
#pragma unroll
                for (int k = 0; k < NV * NV; ++k) acc[k] = 0.0;
                for (int q = 0; q < NQ; ++q) {
                    double Ga[4], Gb[4];
#pragma unroll
                    for (int sl = 0; sl < 4; ++sl) { Ga[sl] = S.G[q][a][sl]; Gb[sl] = S.G[q][b][sl]; }
                    F::kacc(acc, Ga, Gb, S.D[q]);
                }
                double* dst = A.Kval + (size_t)A.emap[e * (NA * NA) + p] * (NV * NV);
#pragma unroll
                for (int k = 0; k < NV * NV; ++k)
                    if (acc[k] != 0.0) red_add(dst + k, acc[k]);
            }
        }
    }
}

}  // namespace mfb
#endif  // __CUDACC__
)MFBSKEL";
