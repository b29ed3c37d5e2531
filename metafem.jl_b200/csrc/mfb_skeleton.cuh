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
//   emap   [n_el][NA*NA]   int32  scatter target of node pair (a,b): its entry in the node-graph CSR, or one of the entry's side
//                                 slots U + s (deterministic scatter: every accumulator receives at most two atomic adds)
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
#define MFB_MAX_QP 16
#endif

struct MfbArgs {
    // mesh
    const int* conn;
    const double* xyz;
    const int* emap;            // scatter targets of the items' node pairs: [n_el][NA*NA] (domain: entries or side slots of the
                                // deterministic scatter) or [n_items][NA*NA] (boundary group: entries of the facet's host element)
    const int* rmap;            // residual scatter targets [n_el][NA]: nodes or residual side slots (boundary: conn itself)
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
    // INTEGRATION_POINT_VAR arrays, [n_q, n_el] column-major in REFERENCE element order (domain blocks only)
    const int* elem_ref;              // [n_items] reference element (0-based) of each internal element
    const double* qpi[MFB_MAX_QP];    // read by the point function (outputs of the user callback)
    double* qpo[MFB_MAX_QP];          // written by the argument-evaluation kernel (inputs of the user callback)
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

// Radial-return update of the reference's J2 example (iterate_stress!, examples/hypo_elastic_plasticity/J2Plasticity.jl:118-188)
// for ONE quadrature point. e_arg: the callback's arguments e11 e12 e13 e22 e23 e33; ep, b, Y: committed state (Voigt order
// 11 22 33 23 13 12); outputs: the trial state ep_eval, b_eval, Y_eval. Returns true if the point yielded.
// Used by the stand-alone return-map kernel (mfb_qp.cu) and, inlined, by element kernels with a fused callback.
__device__ __forceinline__ bool j2_return_map(const double (&e_arg)[6], const double (&ep_in)[6], const double (&b_in)[6], double Y,
                                              double lambda, double mu, double Eb, double Ep, double f_res,
                                              double (&ep)[6], double (&b)[6], double& Yn) {
    // assemble_strain (:103-112): Voigt slots of the six arguments
    const double et[6] = {e_arg[0], e_arg[3], e_arg[5], e_arg[4], e_arg[2], e_arg[1]};
    double s[6], ee[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { ep[k] = ep_in[k]; b[k] = b_in[k]; ee[k] = et[k] - ep_in[k]; }
    // estimate_stress (:114-126) on e_test - ep
    const double tr = (ee[0] + ee[1]) + ee[2];
#pragma unroll
    for (int k = 0; k < 6; ++k) s[k] = (2 * mu) * ee[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) s[k] += lambda * tr;
#pragma unroll
    for (int k = 0; k < 6; ++k) s[k] -= b[k];
    const double skk = ((s[0] + s[1]) + s[2]) / 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) s[k] -= skk;
    // sum over all (i, j): off-diagonal Voigt slots count twice (:153-155)
    double s2 = 0.0;
#pragma unroll
    for (int ii = 0; ii < 3; ++ii)
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
            const int v = ii == jj ? ii : (ii + jj == 3 ? 3 : (ii + jj == 2 ? 4 : 5));   // (2,3)->4th, (1,3)->5th, (1,2)->6th
            s2 += s[v] * s[v];
        }
    const double mag = sqrt(s2);
    const double f = sqrt(3.0 / 2.0) * mag - Y;
    Yn = Y;
    if (!(f > f_res)) return false;
    const double lp = sqrt(3.0 / 2.0) * f / (3 * mu + Eb + Ep);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double nd = s[k] / mag;
        ep[k] = ep[k] + nd * lp;
        b[k] = b[k] + (2.0 / 3.0 * Eb) * nd * lp;
    }
    Yn = Y + (sqrt(2.0 / 3.0) * Ep) * lp;
    return true;
}

// Form contract (emitted code):
//   static constexpr int NV, NA, NQ, L1 (= max_time_level + 1), BOUNDARY (0/1), LINEAR (0/1),
//                        NW (inner words), NCW (cp words), NC (cp fields), HAS_RES, HAS_K, TPB,
//                        NSD (dual slots used by K terms), KS (base slots used), ND (= NV*NSD*NV*KS dense tangent entries),
//                        W (warps with tangent tiles), LPW (tiles = active lanes per warp), CG (column groups),
//                        NTC (even; columns per tile, CG*NTC >= NA), SMEM (bytes);
//   __device__ static constexpr int dslot(int i), bslot(int i);        // slot ids (0:N, 1..3: d/dx) of the used dual/base slots
//   static constexpr int NGS; gslot(slot) -> storage index of a slot in G (or -1), gslot_id(k) -> slot id of storage index k
//   __device__ static constexpr int wslot(int k), wlev(int k), wpos(int k);   // inner word k: slot, time level, variable
//   __device__ static constexpr int cslot(int k), cfield(int k);              // cp word k: slot, field index
//   static constexpr int EVAL (0/1), NQPI (integration-point words read), NQPO (callback arguments written);
//   __device__ static void point(const double* w, const double* c, const double* nrm, const double* qv, size_t qidx /* q + NQ * reference element */, const MfbArgs& A,
//                                double* R /*[NV*4], zeroed*/, double* D /*[ND], zeroed*/);   // NOT yet weighted
//   __device__ static void qp_eval(const double* w, const double* c, const MfbArgs& A, double* out /*[NQPO]*/);  // EVAL forms:
//        // arguments of the user callback at this quadrature point (phase A of the two-phase update, 08_Tensor.jl:175-183)
//        // D[((dp*NSD + dsi)*NV + bp)*KS + ksi] = d(residual integrand of dual (dp, dslot(dsi)))/d(word (bp, bslot(ksi))) * K_params[td]
//
// One thread block works on one item at a time (grid-stride over items); TPB threads:
//   phase A   gather node data of the element into shared memory; prefetch the element's scatter map
//   phase B1  one pass over the reference-element tables for BOTH the Jacobian and the fields: item (q, X) accumulates
//             J[q][i][X] = sum_a dN_X[q][a] x_i[a] and the reference gradient (X = 0: the value) of every field
//             component gu[q][var][X] = sum_a dN_X[q][a] u_var[a]                              (NQ*4 items; _Var_Basic)
//   phase B2  per q: inverse Jacobian, w*detJ (facets: tangents, normal, w*|t1 x t2|)           (NQ threads)
//   phase B3  warp 0: words at the q-points (gradient words = reference gradients x inverse Jacobian) and the emitted
//             point function -> R[q][.], D[q][.] (weighted);  the other warps meanwhile: physical gradients
//             G[q][slot][a] = dN[q][a] . I[q]                                                   (NQ*NA outputs)
//   phase C   tangent, sum-factorised per quadrature point (replaces one _Kval_Basic launch per term). One lane owns the
//             NV x NTC tile of rows (a, dp, bp = 0..NV-1) and columns [cg*NTC, cg*NTC + NTC); no barrier and no shared-memory
//             intermediate inside the q loop:
//             stage 1  T[bp][ks] = sum_dsi G[q][dslot(dsi)][a] * D[q][dp][dsi][bp][ks]         (registers; D[q][dp][.] is one
//                      contiguous, 16 B aligned run read with 128-bit loads, at most two distinct dp per warp)
//             stage 2  acc[bp][c] += T[bp][ks] * G[q][bslot(ks)][cg*NTC + c]                   (G row: warp-uniform 128-bit loads)
//             the residual r[a][dp] = sum_q sum_slot G[q][slot][a] R[q][dp][slot] (_Res_Basic) rides in the same q loop
//   phase D   element matrix -> shared memory in (node pair, block entry) order -> red.global.add.f64 with the NV*NV
//             entries of a node pair on adjacent lanes (72 contiguous bytes for NV = 3: ~3 L2 sectors per pair, not 9).
template <class F>
struct Smem {
    static constexpr int MROWS = F::NA * F::NV * F::NV;
    static constexpr int NAP = F::NA + (F::NA & 1);        // G rows padded to an even length (16 B aligned rows)
    static constexpr int DPB = F::NSD * F::NV * F::KS;     // tangent entries per dual variable dp ...
    static constexpr int DPS = DPB + (DPB & 1);            // ... padded to an even stride
    static constexpr int NDS = F::ND > 0 ? F::NV * DPS : 1;
    static constexpr int NVLI = F::LINEAR ? 0 : F::L1 * F::NV;   // field components interpolated: unknowns at every level ...
    static constexpr int NVL = NVLI + F::NC;                     // ... then the CONTROLPOINT_VAR fields
    static constexpr int NGS = F::NGS;                           // gradient slots actually stored (used by a dual, a base or a residual term)
    static constexpr int RLD = (NVL > 0 ? NVL : 1) * 4;          // leading dimension of gu, and of R which reuses its storage
    struct alignas(16) GD {
        double G[F::NQ][NGS > 0 ? NGS : 1][NAP];   // G[q][gslot(slot)][a]
        double D[F::NQ][NDS];                      // D[q][dp][(dsi*NV + bp)*KS + ks], dp stride DPS
    };
    struct alignas(16) Geo {
        double I[F::NQ][9];                        // Jacobian, then its inverse
        double wgt[F::NQ];
        double nrm[F::NQ][3];
    };
    union {
        GD gd;
        double Ke[F::HAS_K ? MROWS * F::NA : 1];   // phase D staging, (node pair, block entry) order (G and D are dead by then)
    };
    union {
        Geo geo;                                   // phases B1..B3
        int em[F::HAS_K ? F::NA * F::NA : 1];      // phases C, D: node pair -> block-CSR entry of this element
    };
    // gu[q][var][X]: value (X = 0) and reference gradient of every field component (phases B1..B3). The point-function
    // thread of q-point q overwrites ITS row with the weighted residual integrands R[q][v*4 + slot] once it has read it.
    alignas(16) double gu[F::NQ][RLD];
    double xe[F::NA][3];
    double ue[F::L1][F::NA][F::NV];
    double ce[F::NC > 0 ? F::NC : 1][F::NA];
    int rnode[F::NA];                              // residual scatter target of every local node
};

template <class F>
__device__ __forceinline__ void assemble(const MfbArgs& A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<F>& S = *reinterpret_cast<Smem<F>*>(smem_raw);
    static_assert(sizeof(Smem<F>) <= F::SMEM, "emitter under-estimated the shared-memory footprint");
    const int tid = threadIdx.x;
    constexpr int NA = F::NA, NQ = F::NQ, NV = F::NV, BB = NV * NV, MROWS = NA * NV * NV, TPB = F::TPB;
    constexpr int NVLI = Smem<F>::NVLI, NVL = Smem<F>::NVL;
    static_assert(TPB >= 64 && TPB % 32 == 0, "the block needs one point-function warp and at least one geometry warp");
    static_assert(!F::HAS_RES || NVL >= NV, "R reuses the storage of gu");
    constexpr int NGS = F::NGS;

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
        for (int a = tid; a < NA; a += TPB) {
            int g = A.conn[e * NA + a];
            S.rnode[a] = A.rmap[e * NA + a];
            S.xe[a][0] = A.xyz[3 * (size_t)g + 0];
            S.xe[a][1] = A.xyz[3 * (size_t)g + 1];
            S.xe[a][2] = A.xyz[3 * (size_t)g + 2];
        }
        if (!F::LINEAR) {
            for (int i = tid; i < F::L1 * NA * NV; i += TPB) {
                int v = i % NV, a = (i / NV) % NA, l = i / (NV * NA);
                int g = A.conn[e * NA + a];
                S.ue[l][a][v] = A.u[((size_t)l * A.N + g) * NV + v];
            }
        }
        for (int i = tid; i < F::NC * NA; i += TPB) {
            int a = i % NA, c = i / NA;
            S.ce[c][a] = A.cpv[c][A.conn[e * NA + a]];
        }
        constexpr int NEM = F::HAS_K ? (NA * NA + TPB - 1) / TPB : 1;
        int emr[NEM];                                   // scatter map of this element: in flight during phase B
        if constexpr (F::HAS_K) {
            const int* em = A.emap + (F::BOUNDARY ? item : e) * (NA * NA);
#pragma unroll
            for (int k = 0; k < NEM; ++k) emr[k] = tid + k * TPB < NA * NA ? em[tid + k * TPB] : 0;
        }
        __syncthreads();
        // ---- phase B1: Jacobian and reference gradients of every field, one pass over the tables ----
        for (int o = tid; o < NQ * 4; o += TPB) {
            const int q = o >> 2, X = o & 3;            // X = 0: shape functions, 1..3: d/dX
            const double* dN = ref + (X * NQ + q) * NA;
            double j0 = 0.0, j1 = 0.0, j2 = 0.0;
            double acc[NVL > 0 ? NVL : 1];
#pragma unroll
            for (int v = 0; v < NVL; ++v) acc[v] = 0.0;
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                const double d = dN[a];
                j0 += d * S.xe[a][0];
                j1 += d * S.xe[a][1];
                j2 += d * S.xe[a][2];
#pragma unroll
                for (int v = 0; v < NVLI; ++v) acc[v] += d * S.ue[v / NV][a][v % NV];
#pragma unroll
                for (int c = 0; c < F::NC; ++c) acc[NVLI + c] += d * S.ce[c][a];
            }
            if (X > 0) {
                S.geo.I[q][0 * 3 + X - 1] = j0;
                S.geo.I[q][1 * 3 + X - 1] = j1;
                S.geo.I[q][2 * 3 + X - 1] = j2;
            }
#pragma unroll
            for (int v = 0; v < NVL; ++v) S.gu[q][v * 4 + X] = acc[v];
        }
        __syncthreads();
        // ---- phase B2: inverse, weight, normal -----------------------------------------------
        for (int q = tid; q < NQ; q += TPB) {
            double J[3][3], I[3][3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int X = 0; X < 3; ++X) J[i][X] = S.geo.I[q][i * 3 + X];
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
#pragma unroll
            for (int X = 0; X < 3; ++X)
#pragma unroll
                for (int s = 0; s < 3; ++s) S.geo.I[q][X * 3 + s] = I[X][s];
            S.geo.wgt[q] = wgt;
            S.geo.nrm[q][0] = nrm[0]; S.geo.nrm[q][1] = nrm[1]; S.geo.nrm[q][2] = nrm[2];
        }
        __syncthreads();
        // ---- phase B3: warp 0 evaluates the point function, the other warps the physical gradients ----
        if (tid < 32) {
            for (int q = tid; q < NQ; q += 32) {
                const double* I = S.geo.I[q];
                double w[F::NW > 0 ? F::NW : 1], c[F::NCW > 0 ? F::NCW : 1];
#pragma unroll
                for (int k = 0; k < F::NW + F::NCW; ++k) {
                    const int slot = k < F::NW ? F::wslot(k) : F::cslot(k - F::NW);
                    const int var = k < F::NW ? F::wlev(k) * NV + F::wpos(k) : NVLI + F::cfield(k - F::NW);
                    const double* gu = &S.gu[q][var * 4];
                    const double v = slot == 0 ? gu[0]
                                               : (gu[1] * I[0 * 3 + slot - 1] + gu[2] * I[1 * 3 + slot - 1]) + gu[3] * I[2 * 3 + slot - 1];
                    if (k < F::NW) w[k] = v; else c[k - F::NW] = v;
                }
                if constexpr (F::EVAL) {
                    // argument arrays of the quadrature-point callback (phase A of the two-phase update)
                    double out[F::NQPO > 0 ? F::NQPO : 1];
                    F::qp_eval(w, c, A, out);
                    const size_t qbase = (size_t)A.elem_ref[e] * NQ;
#pragma unroll
                    for (int k = 0; k < F::NQPO; ++k) A.qpo[k][qbase + q] = out[k];
                } else {
                    const double wgt = S.geo.wgt[q];
                    const double nrm[3] = {S.geo.nrm[q][0], S.geo.nrm[q][1], S.geo.nrm[q][2]};
                    double qv[F::NQPI > 0 ? F::NQPI : 1];
                    size_t qidx = 0;
                    if constexpr (F::NQPI > 0) {
                        qidx = (size_t)A.elem_ref[e] * NQ + q;
#pragma unroll
                        for (int k = 0; k < F::NQPI; ++k) qv[k] = A.qpi[k][qidx];
                    }
                    double R[NV * 4], D[F::ND > 0 ? F::ND : 1];
#pragma unroll
                    for (int k = 0; k < NV * 4; ++k) R[k] = 0.0;
#pragma unroll
                    for (int k = 0; k < F::ND; ++k) D[k] = 0.0;
                    F::point(w, c, nrm, qv, qidx, A, R, D);
#pragma unroll
                    if constexpr (F::HAS_RES)
                        for (int k = 0; k < NV * 4; ++k) S.gu[q][k] = R[k] * wgt;       // all words of this q are in registers by now
                    if constexpr (F::ND > 0) {
                        constexpr int DPB = Smem<F>::DPB, DPS = Smem<F>::DPS;
#pragma unroll
                        for (int k = 0; k < F::ND; ++k) S.gd.D[q][(k / DPB) * DPS + (k % DPB)] = D[k] * wgt;
                    }
                }
            }
        } else if (!F::EVAL) {
            for (int o = tid - 32; o < NQ * NA; o += TPB - 32) {
                const int q = o / NA, a = o - q * NA;
                const double d0 = ref[(1 * NQ + q) * NA + a], d1 = ref[(2 * NQ + q) * NA + a], d2 = ref[(3 * NQ + q) * NA + a];
                const double* I = S.geo.I[q];
#pragma unroll
                for (int k = 0; k < NGS; ++k) {
                    const int s = F::gslot_id(k);                    // slot id of storage index k: 0 = N, 1..3 = d/dx
                    S.gd.G[q][k][a] = s == 0 ? ref[(0 * NQ + q) * NA + a]
                                             : (d0 * I[0 * 3 + s - 1] + d1 * I[1 * 3 + s - 1]) + d2 * I[2 * 3 + s - 1];
                }
            }
        }
        if constexpr (F::EVAL) continue;
        __syncthreads();   // geo is dead from here on: its storage becomes em
        // ---- phase C: tangent (+ residual) ----------------------------------------------------
        if constexpr (F::HAS_K) {
#pragma unroll
            for (int k = 0; k < NEM; ++k)
                if (tid + k * TPB < NA * NA) S.em[tid + k * TPB] = emr[k];
            constexpr int NTC = F::NTC, KS = F::KS, NSD = F::NSD, CG = F::CG, LPW = F::LPW, NPAIR = NA * NV;
            constexpr int DPS = Smem<F>::DPS;
            static_assert(NTC % 2 == 0 && CG * NTC >= NA && LPW <= 32 && F::W * LPW >= NPAIR * CG && F::W * 32 <= TPB, "bad tangent tiling");
            const int lane = tid & 31, warp = tid >> 5;
            const int t = warp * LPW + lane;                         // tile index: column group major, then dp, then node
            const bool tile_on = warp < F::W && lane < LPW && t < NPAIR * CG;
            const int cgi = tile_on ? t / NPAIR : 0, p = tile_on ? t - cgi * NPAIR : 0;
            const int dp = p / NA, a = p - dp * NA;
            const int b0 = cgi * NTC;
            double acc[NV][NTC];
#pragma unroll
            for (int r = 0; r < NV; ++r)
#pragma unroll
                for (int c = 0; c < NTC; ++c) acc[r][c] = 0.0;
            double racc = 0.0;
            if (tile_on) {
#pragma unroll 1
                for (int q = 0; q < NQ; ++q) {
                    // stage 1 (registers): T[bp][ks] = sum_d G[q][dslot(d)][a] * D[q][dp][d][bp][ks]
                    double gs[NGS > 0 ? NGS : 1];
#pragma unroll
                    for (int k = 0; k < NGS; ++k) gs[k] = S.gd.G[q][k][a];
                    if (F::HAS_RES && cgi == 0) {
                        const double2 r01 = *reinterpret_cast<const double2*>(&S.gu[q][dp * 4]);
                        const double2 r23 = *reinterpret_cast<const double2*>(&S.gu[q][dp * 4 + 2]);
                        const double r4[4] = {r01.x, r01.y, r23.x, r23.y};
#pragma unroll
                        for (int k = 0; k < NGS; ++k) racc += gs[k] * r4[F::gslot_id(k)];
                    }
                    double dv[DPS];
                    const double2* Dq = reinterpret_cast<const double2*>(&S.gd.D[q][dp * DPS]);
#pragma unroll
                    for (int i = 0; i < DPS / 2; ++i) {
                        const double2 v = Dq[i];
                        dv[2 * i] = v.x; dv[2 * i + 1] = v.y;
                    }
                    double T[NV][KS > 0 ? KS : 1];
#pragma unroll
                    for (int bp = 0; bp < NV; ++bp)
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            double tv = 0.0;
#pragma unroll
                            for (int d = 0; d < NSD; ++d) tv += gs[F::gslot(F::dslot(d))] * dv[(d * NV + bp) * KS + ks];
                            T[bp][ks] = tv;
                        }
                    // stage 2: rank-KS update of the lane's NV x NTC tile
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        const double2* gp = reinterpret_cast<const double2*>(&S.gd.G[q][F::gslot(F::bslot(ks))][b0]);
#pragma unroll
                        for (int c = 0; c < NTC / 2; ++c) {
                            const double2 v = gp[c];
#pragma unroll
                            for (int bp = 0; bp < NV; ++bp) {
                                acc[bp][2 * c] += T[bp][ks] * v.x;
                                acc[bp][2 * c + 1] += T[bp][ks] * v.y;
                            }
                        }
                    }
                }
                if (F::HAS_RES && cgi == 0 && racc != 0.0) red_add(A.res + (size_t)S.rnode[a] * NV + dp, racc);
            }
            __syncthreads();        // every warp is done with G, D and R: their storage becomes Ke
            // ---- phase D: stage the element matrix, then scatter with node-pair blocks on adjacent lanes ----
            if (tile_on) {
#pragma unroll
                for (int c = 0; c < NTC; ++c)
                    if (b0 + c < NA) {
#pragma unroll
                        for (int bp = 0; bp < NV; ++bp) S.Ke[(a * NA + b0 + c) * BB + dp * NV + bp] = acc[bp][c];
                    }
            }
            __syncthreads();
            {
                constexpr int DK = TPB % BB, DPAIR = TPB / BB;      // entry i = pair*BB + k advances by TPB per step
                int pr = tid / BB, k = tid - pr * BB;
#pragma unroll 4
                for (int i = tid; i < MROWS * NA; i += TPB) {
                    const double v = S.Ke[i];
                    if (v != 0.0) red_add(A.Kval + (size_t)S.em[pr] * BB + k, v);
                    k += DK; pr += DPAIR;
                    if (k >= BB) { k -= BB; ++pr; }
                }
            }
        } else if constexpr (F::HAS_RES) {
            // residual-only kernels (_Res_Basic): ONE add per (element, node, variable) -- the deterministic scatter relies on
            // every accumulator receiving at most two adds, so the q range is not split over threads
            for (int i = tid; i < NA * NV; i += TPB) {
                const int v = i % NV, a = i / NV;
                double s = 0.0;
                for (int q = 0; q < NQ; ++q) {
#pragma unroll
                    for (int k = 0; k < NGS; ++k) s += S.gd.G[q][k][a] * S.gu[q][v * 4 + F::gslot_id(k)];
                }
                if (s != 0.0) red_add(A.res + (size_t)S.rnode[a] * NV + v, s);
            }
        }
    }
}

}  // namespace mfb
#endif  // __CUDACC__
