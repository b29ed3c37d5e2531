// mfb_skeleton.cuh -- fused element-assembly skeleton for sm_100a (FP64, no tensor cores).
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
//                        NW (inner words), NCW (cp words), NC (cp fields), HAS_RES, HAS_K, TPB,
//                        NSD (dual slots used by K terms), KS (base slots used), ND (= NV*NSD*NV*KS dense tangent entries),
//                        MT, NTC (register tile of the tangent contraction: MT rows x NTC columns per thread);
//   __device__ static constexpr int dslot(int i), bslot(int i);        // slot ids (0:N, 1..3: d/dx) of the used dual/base slots
//   template <class S> __device__ static void words(const S& s, int q, double* w, double* c);
//        // w[k] = interpolation of inner word k, c[k] = of CONTROLPOINT_VAR word k  (_Var_Basic)
//   __device__ static void point(const double* w, const double* c, const double* nrm, const MfbArgs& A,
//                                double* R /*[NV*4], zeroed*/, double* D /*[ND], zeroed*/);   // NOT yet weighted
//        // D[((dp*NSD + dsi)*NV + bp)*KS + ksi] = d(residual integrand of dual (dp, dslot(dsi)))/d(word (bp, bslot(ksi))) * K_params[td]
//
// One thread block works on one item at a time (grid-stride over items); TPB threads.
//   phase A  gather node data of the element into shared memory
//   phase B  one thread per quadrature point: Jacobian, inverse, w*detJ (or facet normal and w*|t1 x t2|), physical
//            gradients G[q][slot][a], interpolation of every word, emitted point function -> R[q][.], D[q][.] (weighted)
//   phase C1 residual: r[a][v] = sum_q sum_slot G[q][slot][a] R[q][v][slot]                       (_Res_Basic)
//   phase C2 tangent, sum-factorised per quadrature point (replaces one _Kval_Basic launch per term):
//            stage 1  T[ks][(a,dp,bp)] = sum_dsi G[q][dslot(dsi)][a] * D[q][dp][dsi][bp][ks]       (NA*NV threads)
//            stage 2  K[(a,dp,bp)][b] += sum_ks T[ks][(a,dp,bp)] * G[q][bslot(ks)][b]             (MT x NTC register tiles,
//                     operands read from shared memory with 128-bit loads: 2 MT + NTC/2 wavefronts per MT*NTC DFMA)
//            then red.global.add.f64 of the element matrix into the block-CSR values.
template <int NA>
__device__ __forceinline__ double interp(const double* Ga /*[NA] contiguous*/, const double* u, int ustride) {
    double s = 0.0;
#pragma unroll 4
    for (int a = 0; a < NA; ++a) s += Ga[a] * u[a * ustride];
    return s;
}

template <class F>
struct Smem {
    static constexpr int MROWS = F::NA * F::NV * F::NV;
    double G[F::NQ][4][F::NA];                     // first two members stay 16 B aligned (even element counts)
    double T[2][F::KS > 0 ? F::KS : 1][MROWS];
    double D[F::NQ][F::ND > 0 ? F::ND : 1];
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
    constexpr int NA = F::NA, NQ = F::NQ, NV = F::NV, BB = NV * NV, MROWS = NA * NV * NV;

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
                S.G[q][0][a] = ref[(0 * NQ + q) * NA + a];
#pragma unroll
                for (int s = 0; s < 3; ++s) S.G[q][1 + s][a] = (d0 * I[0][s] + d1 * I[1][s]) + d2 * I[2][s];
            }
            double w[F::NW > 0 ? F::NW : 1], c[F::NCW > 0 ? F::NCW : 1];
            F::words(S, q, w, c);
            double R[NV * 4], D[F::ND > 0 ? F::ND : 1];
#pragma unroll
            for (int k = 0; k < NV * 4; ++k) R[k] = 0.0;
#pragma unroll
            for (int k = 0; k < F::ND; ++k) D[k] = 0.0;
            F::point(w, c, nrm, A, R, D);
#pragma unroll
            for (int k = 0; k < NV * 4; ++k) S.R[q][k] = R[k] * wgt;
#pragma unroll
            for (int k = 0; k < F::ND; ++k) S.D[q][k] = D[k] * wgt;
        }
        __syncthreads();
        // ---- phase C1: residual ------------------------------------------------------------
        if constexpr (F::HAS_RES) {
            for (int i = tid; i < NA * NV; i += F::TPB) {
                int v = i % NV, a = i / NV;
                double s = 0.0;
                for (int q = 0; q < NQ; ++q) {
#pragma unroll
                    for (int sl = 0; sl < 4; ++sl) s += S.G[q][sl][a] * S.R[q][v * 4 + sl];
                }
                if (s != 0.0) red_add(A.res + (size_t)S.node[a] * NV + v, s);
            }
        }
        // ---- phase C2: tangent ---------------------------------------------------------------
        if constexpr (F::HAS_K) {
            constexpr int MT = F::MT, NTC = F::NTC, KS = F::KS, NSD = F::NSD;
            constexpr int NRG = (MROWS + MT - 1) / MT, NCG = (NA + NTC - 1) / NTC;
            static_assert(NRG * NCG <= F::TPB, "tangent tiling needs more threads than the block has");
            static_assert(MT % 2 == 0 && NTC % 2 == 0 && NA % 2 == 0 && MROWS % 2 == 0, "128-bit shared-memory loads need even tiles");
            const int rg = tid % NRG, cg = tid / NRG;
            const int m0 = rg * MT, b0 = cg * NTC;
            const bool tile_on = tid < NRG * NCG;
            double acc[MT][NTC];
#pragma unroll
            for (int r = 0; r < MT; ++r)
#pragma unroll
                for (int c = 0; c < NTC; ++c) acc[r][c] = 0.0;
            for (int q = 0; q < NQ; ++q) {
                const int buf = q & 1;
                // stage 1: thread (a, dp) fills T[ks][(a,dp,bp)] for all bp, ks
                for (int i = tid; i < NA * NV; i += F::TPB) {
                    const int a = i / NV, dp = i - a * NV;
                    double g[NSD > 0 ? NSD : 1];
#pragma unroll
                    for (int d = 0; d < NSD; ++d) g[d] = S.G[q][F::dslot(d)][a];
                    const double* Dq = &S.D[q][dp * NSD * NV * KS];
#pragma unroll
                    for (int bp = 0; bp < NV; ++bp)
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            double t = 0.0;
#pragma unroll
                            for (int d = 0; d < NSD; ++d) t += g[d] * Dq[(d * NV + bp) * KS + ks];
                            S.T[buf][ks][(a * NV + dp) * NV + bp] = t;
                        }
                }
                __syncthreads();   // T[buf] complete; T[buf^1] (read in the previous iteration) may now be overwritten next time
                if (tile_on) {
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        double ta[MT], gb[NTC];
                        const double2* tp = reinterpret_cast<const double2*>(&S.T[buf][ks][m0]);
                        const double2* gp = reinterpret_cast<const double2*>(&S.G[q][F::bslot(ks)][b0]);
#pragma unroll
                        for (int r = 0; r < MT / 2; ++r) {
                            const bool ok = m0 + 2 * r < MROWS;
                            double2 v = ok ? tp[r] : make_double2(0.0, 0.0);
                            ta[2 * r] = v.x; ta[2 * r + 1] = v.y;
                        }
#pragma unroll
                        for (int c = 0; c < NTC / 2; ++c) {
                            const bool ok = b0 + 2 * c < NA;
                            double2 v = ok ? gp[c] : make_double2(0.0, 0.0);
                            gb[2 * c] = v.x; gb[2 * c + 1] = v.y;
                        }
#pragma unroll
                        for (int r = 0; r < MT; ++r)
#pragma unroll
                            for (int c = 0; c < NTC; ++c) acc[r][c] += ta[r] * gb[c];
                    }
                }
            }
            if (tile_on) {
#pragma unroll
                for (int r = 0; r < MT; ++r) {
                    const int m = m0 + r;
                    if (m >= MROWS) break;
                    const int a = m / BB, rem = m - a * BB;
                    const int* em = A.emap + e * (NA * NA) + a * NA + b0;
#pragma unroll
                    for (int c = 0; c < NTC; ++c) {
                        if (b0 + c >= NA) break;
                        const double v = acc[r][c];
                        if (v != 0.0) red_add(A.Kval + (size_t)em[c] * BB + rem, v);
                    }
                }
            }
        }
    }
}

}  // namespace mfb
#endif  // __CUDACC__
