/* metafem_b200.h -- C ABI of libmetafem_b200.so
 *
 * B200-native (sm_100a) replacement for MetaFEM.jl's hot path: element evaluation ->
 * scatter-assembly into CSR -> Jacobi-preconditioned Krylov solve. These entry points are what
 * the Julia side binds with `ccall`; every test and benchmark in this
 * repository drives exactly the same symbols through Python ctypes.
 *
 * Conventions (reference: FEM_Int = Int32, FEM_Float = Float64, src/misc/02_Global_Macros.jl:123-124):
 *  - all index arrays are 1-based int32, all tables column-major (first index fastest), exactly
 *    as the reference stores them;
 *  - every array argument may be a device pointer, a CUDA unified pointer (the reference's
 *    default array type, src/misc/04_GPU_Utils.jl:8) or a plain/pinned host pointer; host
 *    buffers are staged through the context's stream inside the call;
 *  - small parameter vectors (K_params, alpha/beta/gamma_params, sparse_mapping, el_cp_outer_id, name lists) and all
 *    output scalars are HOST pointers;
 *  - the library borrows caller memory only for the duration of a call;
 *  - every function returns MFB_OK (0) or a negative error code; mfb_last_error() gives text.
 *    MFB_NOT_CONVERGED (1) is a warning: results are valid, the reference only prints in that
 *    case (src/solver/linear_solver/02_Preconditioner.jl:66-68).
 */
#ifndef METAFEM_B200_H
#define METAFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mfb_ctx mfb_ctx;

enum {
    MFB_OK = 0,
    MFB_NOT_CONVERGED = 1,
    MFB_ERR_CUDA = -1,
    MFB_ERR_NVRTC = -2,
    MFB_ERR_ARG = -3,
    MFB_ERR_STATE = -4,
    MFB_ERR_NCCL = -5
};

/* Krylov methods: Sv_func! of iterative_Solve! (02_Preconditioner.jl:32) */
enum {
    MFB_IDRS = 0,         /* idrs!,         04_IDRs.jl:26-95 */
    MFB_BICGSTABL_GS = 1, /* bicgstabl_GS!, 03_BiCGstabl.jl:18-96 */
    MFB_BICGSTABL = 2,    /* bicgstabl!,    03_BiCGstabl.jl:98-162 (MR part by LU of the Gram matrix) */
    MFB_GMRES = 3,        /* gmres!,        05_GMRES.jl:46-101 (restart length s) */
    MFB_CGS = 4,          /* cgs!,          07_CGS.jl:10-50 */
    MFB_CGS2 = 5,         /* cgs2!,         07_CGS.jl:52-105 */
    MFB_TFQMR = 6,        /* tfqmr!,        08_QMR.jl:3-76 */
    MFB_LSQR = 7,         /* lsqr!,         06_LSQR.jl:10-73 (needs A' x: transposed block SpMV) */
    MFB_IDRS_ORIGINAL = 8 /* idrs_original!, 04_IDRs.jl:97-169 ("not used" in the reference; restated as it is, including :147) */
};
/* Pr_func! / Pl_func of iterative_Solve! (02_Preconditioner.jl:78-177) */
enum { MFB_PR_JACOBI = 0 /* Pr_Jacobi! by diagonal (default) */, MFB_PR_JACOBI_COLUMN = 1 /* normalized_by_column = true */,
       MFB_PR_IDENTITY = 2 };
enum { MFB_PL_IDENTITY = 0 /* default */, MFB_PL_JACOBI = 1 /* Pl_Jacobi by diagonal */, MFB_PL_JACOBI_ROW = 2 /* normalized_by_row */,
       MFB_PL_ILU = 3 /* Pl_ILU (:179-194): zero-fill block ILU of the right-scaled matrix, level-scheduled sweeps; elimination order
                         chosen for parallelism (colour classes of a greedy colouring) where cuSPARSE ilu02! follows the row order,
                         packed factors of the sweeps stored in FP32 (MFB_ILU_FP64=1: doubles) -- same incomplete factorisation
                         property, solutions agree at solver tolerance */ };

/* vectors of GlobalField (src/solver/01_Types.jl:110-132) */
enum { MFB_VEC_X = 0, MFB_VEC_DX = 1, MFB_VEC_X_STAR = 2, MFB_VEC_RESIDUE = 3 };
/* matrices of GlobalField */
enum { MFB_MAT_K_LINEAR = 0, MFB_MAT_K_TOTAL = 1 };

/* ---- context ------------------------------------------------------------------------- */
int mfb_create(mfb_ctx **ctx, int device);
int mfb_destroy(mfb_ctx *ctx);
const char *mfb_last_error(mfb_ctx *ctx);
/* use a caller stream (cudaStream_t); NULL = the context's own stream */
int mfb_set_stream(mfb_ctx *ctx, void *cuda_stream);
/* number of kernels this library has launched on ctx since creation (bench.py "gpu_launches") */
int64_t mfb_launch_count(mfb_ctx *ctx);
int mfb_synchronize(mfb_ctx *ctx);
/* CUDA-event timers on the context's stream. ids: 0 SpMV, 1 K_nonlinear_func, 2 K_linear_func,
 * 3 Krylov solve, 4 domain element kernel, 5 interface exchange-add (multi-GPU), 6 Krylov reductions (+ allreduce),
 * 7 boundary element kernels during assembly, Pl_ILU applications (both sweeps) during a solve. mfb_profile_get sums and
 * clears them: ms[8], count[8]. on = 1: all but 5 and 6; on = 2: also 5 and 6 (one event pair per reduction group and per
 * exchange perturbs the stream: for diagnosis, not for timed runs). */
/* Measured FP64 FMA throughput of the device in TFLOP/s (register-resident FMA chains on every SM): the peak the
 * FP64-bound element kernels are reported against. */
int mfb_measure_fp64_peak(mfb_ctx *ctx, double *tflops);
int mfb_profile_enable(mfb_ctx *ctx, int on);
int mfb_profile_get(mfb_ctx *ctx, double *ms, int64_t *count);

/* ---- mesh tables ---------------------------------------------------------------------
 * Replaces the element tables produced by mesh_Classical + update_Mesh
 * (src/mesh/unstructured_mesh/2_Interface.jl:7-39,98-108; 4_Update_Integrator.jl:2-33):
 * the library takes the connectivity, node coordinates and REFERENCE-element tables and
 * evaluates Jacobians / physical gradients / w*detJ on the fly inside the element kernels, so
 * the 34.6 kB/element integral_vals table is never materialised.
 *   controlpoint_IDs [n_a, n_el]  elements.controlpoint_IDs (== global_cpIDs for one workpiece)
 *   x1,x2,x3         [N]          controlpoints.x1/x2/x3
 *   ref_itp_vals     [n_q, n_a, 2, 2, 2]  Classical_Discretization.ref_itp_vals (max_sd_order = 1)
 *   itg_weight       [n_q]
 */
int mfb_mesh_set(mfb_ctx *ctx, int n_a, int64_t n_el, int64_t N, int n_q,
                 const int32_t *controlpoint_IDs, const double *x1, const double *x2, const double *x3,
                 const double *ref_itp_vals, const double *itg_weight);

/* Second-order mesh tables from a first-order mesh, built on the device (replaces, for the hot path's inputs, the segment /
 * face tables of construct_TotalMesh_3D -- src/mesh/ref_geometry/002_Initialization.jl:113-217 through the GPU hash of
 * src/misc/06_GPU_Dict.jl --, get_BoundaryMesh, and the control-point allocation of mesh_Classical,
 * src/mesh/unstructured_mesh/3_InitializeMesh.jl:70-178):
 *   x1,x2,x3 [n_vert], connections [vpb, n_el] (1-based vertex IDs): the output of make_Brick / read_Mesh;
 *   element-type tables (101_Structures.jl:129-196,224-247), 1-based: segment_vertices [2, n_seg], vertex_cp_ids [vpb],
 *   segment_cp_ids [n_seg] (control-point slot of each vertex / segment), face_vertices [vpf, n_faces].
 * Control points: vertices first in input order (as the reference), then one per segment in sorted order of
 * (max vertex, min vertex) -- deterministic and locality-preserving, where the reference's order is that of its racy hash
 * table. Boundary facets (faces owned by one element) come in (local face, element) order with their centroids, for the
 * script's geometric selection of boundary groups. Results stay on the device (mfb_mesh_build_device_ptrs, to be passed
 * to mfb_mesh_set) and can be copied out (mfb_mesh_build_get; any pointer may be NULL). */
int mfb_mesh_build_second_order(mfb_ctx *ctx, int64_t n_vert, const double *x1, const double *x2, const double *x3,
                                int vpb, int64_t n_el, const int32_t *connections, int n_seg,
                                const int32_t *segment_vertices, const int32_t *vertex_cp_ids,
                                const int32_t *segment_cp_ids, int n_faces, int vpf, const int32_t *face_vertices,
                                int64_t *n_controlpoints, int64_t *n_boundary_facets);
int mfb_mesh_build_get(mfb_ctx *ctx, int32_t *controlpoint_IDs, double *x1, double *x2, double *x3,
                       int32_t *bfacet_element_ID, int32_t *bfacet_element_eindex, double *bfacet_centroids);
int mfb_mesh_build_device_ptrs(mfb_ctx *ctx, const int32_t **controlpoint_IDs, const double **x1, const double **x2,
                               const double **x3);

/* First-order geometry tables of construct_TotalMesh_3D (src/mesh/ref_geometry/002_Initialization.jl:113-217) on the device,
 * for tets (vpb = 4) and hexes (vpb = 8); connections [vpb, n_blocks] are 1-based vertex IDs. Outputs (mfb_total_mesh_get, any
 * pointer may be NULL; 1-based, column-major like the reference's tables):
 *   segment_vertex_IDs [2, n_segments] (max vertex first, :144-164), block_segment_IDs [6|12, n_blocks],
 *   face_vertex_IDs / face_segment_IDs [3|4, n_faces] (:188-213), block_face_IDs [4|6, n_blocks],
 *   boundary_face_IDs [n_boundary_faces] ascending (get_BoundaryMesh :285-290) with the block that owns each of them and its
 *   local face number (what mesh_Classical's facet allocation / specify_eindex computes, 3_InitializeMesh.jl:119-130,165-178).
 * numbering: MFB_NUMBERING_SORTED -- IDs = rank of the (max, next) key, one radix sort (deterministic, locality-preserving);
 *            MFB_NUMBERING_REFERENCE -- the reference's pass-by-pass hash-slot order, FEM_Dict (src/misc/06_GPU_Dict.jl:2-162)
 *            reproduced with sequential insertion (one legal outcome of the reference's racing insertion; serial, for parity). */
enum { MFB_NUMBERING_SORTED = 0, MFB_NUMBERING_REFERENCE = 1 };
int mfb_total_mesh_build(mfb_ctx *ctx, int64_t n_vert, int vpb, int64_t n_blocks, const int32_t *connections, int numbering,
                         int64_t *n_segments, int64_t *n_faces, int64_t *n_boundary_faces);
int mfb_total_mesh_get(mfb_ctx *ctx, int32_t *segment_vertex_IDs, int32_t *block_segment_IDs, int32_t *face_vertex_IDs,
                       int32_t *face_segment_IDs, int32_t *block_face_IDs, int32_t *boundary_face_IDs,
                       int32_t *boundary_face_block, int32_t *boundary_face_eindex);

/* Boundary tables (4_Update_Integrator.jl:35-75; 3_InitializeMesh.jl:119-130,165-178):
 *   bdy_ref_itp_vals       [n_qb, n_a, 2, 2, 2, n_faces]   one table per local face (eindex)
 *   bdy_itg_weights        [n_qb, n_faces]
 *   bdy_tangent_directions [n_qb, 3, 2, n_faces]
 *   facet_element_ID / facet_element_eindex [n_facets]     facets.element_ID / element_eindex
 */
int mfb_facets_set(mfb_ctx *ctx, int n_faces, int n_qb, const double *bdy_ref_itp_vals,
                   const double *bdy_itg_weights, const double *bdy_tangent_directions,
                   int64_t n_facets, const int32_t *facet_element_ID, const int32_t *facet_element_eindex);
/* bg_fIDs[bg_ID] (1_Types.jl:61): facet IDs (1-based) of one boundary group */
int mfb_boundary_group_set(mfb_ctx *ctx, int bg_ID, int64_t n, const int32_t *facet_IDs);

/* ---- DOF numbering + sparsity pattern ---------------------------------------------------
 * Replaces assemble_Global_Variables! / assemble_SparseID! (src/solver/03_GlobalAssembly.jl:6-37,77-168)
 * and sort_CUSPARSE_COO!/generate_J_ptr (src/misc/04_GPU_Utils.jl:87-118).
 *   n_var          number of basic variables (length(basic_vars))
 *   max_time_level highest time-derivative order (x holds (max_time_level+1)*n_var*N values)
 *   sparse_mapping [2, n_blocks] (dual_pos, base_pos) of block m, 0-based, in block order
 * Outputs: nnz = n_blocks*sparse_unitsize, sparse_unitsize = number of distinct (node,node) pairs.
 * The pattern is kept on the device in CSR order; sparse IDs are DEFINED as CSR positions
 * (the reference's hash-slot order is an internal permutation), so
 * K_val_ids is the identity.
 */
int mfb_pattern_build(mfb_ctx *ctx, int n_var, int max_time_level, int n_blocks,
                      const int32_t *sparse_mapping, int64_t *nnz, int64_t *sparse_unitsize);
/* Reference-layout copies for parity: K_I, K_J [nnz] (sorted by row then column, 1-based),
 * K_J_ptr [n+1] (1-based), K_val_ids [nnz] (identity, see above). Any pointer may be NULL. */
int mfb_pattern_get(mfb_ctx *ctx, int32_t *K_I, int32_t *K_J, int32_t *K_J_ptr, int32_t *K_val_ids);
/* elements.sparse_IDs_by_el [n_a, n_a, n_el] (03_GlobalAssembly.jl:111-118; row = dual node a, column = base node b, elements
 * in the caller's order) for variable block `block` (index into sparse_mapping): the 1-based position, in the arrays of
 * mfb_pattern_get / mfb_matrix_get, of the entry that node pair (a, b) of element e accumulates into. The reference's
 * entry ID `sparse_IDs_by_el[a,b,e] + block * sparse_unitsize` (06_FEM_Kernel.jl:36) addresses its hash-ordered value
 * array; through K_val_ids it names the same CSR position, which is what this getter returns directly. */
int mfb_sparse_ids_get(mfb_ctx *ctx, int block, int32_t *sparse_IDs_by_el);

/* ---- fields ------------------------------------------------------------------------------
 * CONTROLPOINT_VAR external fields (controlpoints.<sym>, 05_CodeGenerator.jl:28-35), by local symbol. */
int mfb_field_set(mfb_ctx *ctx, const char *local_sym, const double *values /* [N] */);
/* GLOBAL_VAR scalars (physics.global_vars[:sym], read at call time, 05_CodeGenerator.jl:21-27) */
int mfb_global_set(mfb_ctx *ctx, const char *sym, double value);
/* GlobalField vectors; n = (max_time_level+1)*n_var*N for x/dx/x_star, n_var*N for residue.
 * Reference layout: index = g + v*N + l*n  (03_GlobalAssembly.jl:52). */
int mfb_vector_set(mfb_ctx *ctx, int which, const double *values, int64_t n);
int mfb_vector_get(mfb_ctx *ctx, int which, double *values, int64_t n);
/* CSR values (K_linear / K_total) in reference CSR order: K[K_val_ids] of the reference */
int mfb_matrix_get(mfb_ctx *ctx, int which, double *values, int64_t nnz);

/* ---- generated element kernels ----------------------------------------------------------
 * Replaces compile_Updater_GPU (src/solver/05_CodeGenerator.jl:265-291). `cuda_src` is the CUDA C
 * emitted by the front end's term emitter (sibling of parse_Term2Expr!, src/symbolics/08_Tensor.jl:214-233)
 * on top of the library's assembly skeleton ("mfb_skeleton.cuh", resolvable as an #include);
 * it is compiled with NVRTC for sm_100a. One descriptor per generated block
 * (domain first, then boundary groups, as gen_CodeBody orders them, :156-196). */
typedef struct {
    int32_t kind;                 /* 0 = domain elements, 1 = boundary group */
    int32_t bg_ID;                /* boundary group ID when kind == 1 */
    const char *linear_kernel;    /* entry point writing K_linear, or NULL */
    const char *nonlinear_kernel; /* entry point writing residue (+ K_total), or NULL */
    int32_t n_cp_vars;            /* CONTROLPOINT_VAR fields the kernels read, in argument order */
    const char *const *cp_var_names;
    int32_t n_globals;            /* GLOBAL_VAR scalars, in argument order */
    const char *const *global_names;
    int32_t threads_per_block;    /* launch shape chosen by the emitter */
    int32_t smem_bytes;           /* dynamic shared memory per block */
    int32_t has_nonlinear_K;      /* 1 if the nonlinear kernel adds K terms */
    /* quadrature-point callbacks (INTEGRATION_POINT_VAR words, symbolics/08_Tensor.jl:175-183): */
    const char *eval_kernel;      /* entry point filling the callback's argument arrays, or NULL */
    int32_t n_qp_in;              /* integration-point arrays the nonlinear kernel reads (callback outputs) */
    const char *const *qp_in_names;
    int32_t n_qp_out;             /* integration-point arrays the eval kernel writes (callback arguments) */
    const char *const *qp_out_names;
} mfb_block_desc;
int mfb_kernel_compile(mfb_ctx *ctx, const char *cuda_src, int n_blocks, const mfb_block_desc *blocks);
/* Compile-only check of an emitted translation unit (needs no device): writes the NVRTC log (or the
 * error text) to log_out and, if cubin_path is not NULL, the sm_100a cubin to that file. */
int mfb_kernel_check(const char *cuda_src, const char *cubin_path, char *log_out, int log_len);
/* NVRTC log of the last compilation (valid until the next compile) */
const char *mfb_compile_log(mfb_ctx *ctx);

/* K_linear_func: K_linear .= 0, then every linear gradient term (05_CodeGenerator.jl:265-276) */
int mfb_assemble_linear(mfb_ctx *ctx, const double *K_params, int n_params);
/* K_nonlinear_func: residue .= 0; K_total .= K_linear; residual + nonlinear gradient terms,
 * evaluated at the context's x_star (05_CodeGenerator.jl:278-288). t and dt are the GLOBAL_VARs :t/:dt. */
int mfb_assemble_nonlinear(mfb_ctx *ctx, const double *K_params, int n_params, double t, double dt);

/* ---- integration-point variables and the quadrature-point callback ------------------------
 * The J2 example defines `ep{i,j} = strain_updater(e{1,1}, ..., e{3,3})`
 * (examples/hypo_elastic_plasticity/J2Plasticity.jl:55): the generated updater evaluates the six arguments on whole
 * [n_q, n_el] device arrays, calls Main.strain_updater on them and reads the six outputs as external words
 * (symbolics/08_Tensor.jl:175-183,210). Here the nonlinear update is two-phase:
 *   mfb_eval_qp_args      fills the argument arrays (named "<func>_arg<k>");
 *   the caller runs its callback on the arrays (device pointers from mfb_qp_array, e.g. wrapped as CuArrays,
 *   or the built-in J2 return map below) and leaves the outputs in the arrays named like the output words;
 *   mfb_assemble_nonlinear reads them.
 * Arrays are library-owned, [n_q, n_el] column-major (q fastest) in the reference's element order, zero on creation. */
int mfb_qp_array(mfb_ctx *ctx, const char *name, double **device_ptr, int64_t *n);
int mfb_qp_set(mfb_ctx *ctx, const char *name, const double *values, int64_t n);
int mfb_qp_get(mfb_ctx *ctx, const char *name, double *values, int64_t n);
int mfb_eval_qp_args(mfb_ctx *ctx, double t, double dt);

/* Built-in return map = the example's MaterialState callable, iterate_stress! and update_States!
 * (J2Plasticity.jl:76-198) as ONE element-wise kernel over all quadrature points. State arrays
 * "<prefix>.ep1..6", "<prefix>.b1..6", "<prefix>.Y" (committed) and "<prefix>.b_eval1..6", "<prefix>.Y_eval"
 * are integration-point arrays of the context; ep_eval is written to the arrays named ep_names (Voigt order).
 *   e_names  [6]  argument arrays in the callback's order e11, e12, e13, e22, e23, e33
 *   ep_names [6]  output arrays in Voigt order 11, 22, 33, 23, 13, 12 (symbolics/03_Word.jl:37) */
typedef struct {
    double lambda, mu, Eb, Ep, f_res;
} mfb_j2_params;
int mfb_j2_init(mfb_ctx *ctx, const char *prefix, double Y_initial, const char *const *e_names,
                const char *const *ep_names);
int mfb_j2_iterate_stress(mfb_ctx *ctx, const char *prefix, const mfb_j2_params *params, int64_t *n_yielded);
int mfb_j2_update_states(mfb_ctx *ctx, const char *prefix);
/* Fused variant: when the emitter inlines the return map into the residual kernel (descriptor: no eval_kernel, qp_in_names =
 * committed state "<prefix>.ep1..6, .b1..6, .Y", qp_out_names = ep outputs, "<prefix>.b_eval1..6", ".Y_eval", ".count";
 * parameters as GLOBAL_VARs "<prefix>_lam, _mu, _Eb, _Ep, _fres"), mfb_eval_qp_args / mfb_j2_iterate_stress are not needed;
 * this returns the number of points that yielded in the last mfb_assemble_nonlinear. */
int mfb_j2_yield_count(mfb_ctx *ctx, const char *prefix, int64_t *n_yielded);

/* ---- linear algebra ---------------------------------------------------------------------
 * y = K * x in reference numbering (mul!, src/misc/04_GPU_Utils.jl:131). */
int mfb_spmv(mfb_ctx *ctx, int which_matrix, const double *x, double *y, int64_t n);

/* Development aid: ms per launch of one tuning variant of the 3-variable block SpMV on K_total (0 = production kernel). */
int mfb_spmv_variant_bench(mfb_ctx *ctx, int variant, int reps, double *ms_per_launch, double *max_abs_diff);

typedef struct {
    int32_t passes;        /* restart passes used */
    int32_t iterations;    /* Krylov iterations summed over passes */
    int32_t spmv_count;    /* SpMV launches */
    int32_t converged;     /* 1 if res < tol */
    double residual;       /* final true residual ||b - A x||_2 / sqrt(n) */
    double initial_residual;
} mfb_solve_info;
/* iterative_Solve!(globalfield; Sv_func!, Pr_func! = Pr_Jacobi!, max_pass, maxiter, s)
 * (02_Preconditioner.jl:32-76): solves K_total * delta = residue with right-Jacobi scaling,
 * absolute tolerance on ||r||/sqrt(n); writes delta (reference numbering) to delta_out [n]
 * (may be NULL: result stays in the context for mfb_update_dx). */
int mfb_krylov_solve(mfb_ctx *ctx, int method, int s, int maxiter, int max_pass, double tol,
                     uint64_t seed, double *delta_out, mfb_solve_info *info);

/* Same with the preconditioner choices of iterative_Solve!: Pr_func! in {Pr_Jacobi!, Pr_Jacobi!(normalized_by_column),
 * Identity}, Pl_func in {Identity, Pl_Jacobi, Pl_Jacobi(normalized_by_row)}; with a left preconditioner the per-pass
 * tolerance is rescaled by min(||Pl r|| / ||r||, 1) (:50-53). checkiter: residual check period of tfqmr!.
 * Pl_ILU: see MFB_PL_ILU. */
int mfb_krylov_solve_ex(mfb_ctx *ctx, int method, int s, int maxiter, int max_pass, double tol, uint64_t seed,
                        int pr_mode, int pl_mode, int checkiter, double *delta_out, mfb_solve_info *info);

/* Test / diagnosis of the ILU: factorises K_total as it stands and returns max |(L U - A)_ij| over the sparsity pattern relative to
 * max |A_ij| (zero for an exact incomplete factorisation up to rounding), the number of dependency levels of the elimination order,
 * and applies U^-1 L^-1 to v_inout (reference layout, n = n_var*N; may be NULL). */
int mfb_ilu_selftest(mfb_ctx *ctx, double *rel_defect, int32_t *n_levels, double *v_inout, int64_t n);

/* ---- time stepping (src/solver/04_Time_Domain.jl) ----------------------------------------
 * Device-resident versions of initialize_dx! (:20-30), update_x_star! (:41-49),
 * update_dx! with the last solve's delta (:32-39, called with -delta as update_OneStep! does, :77),
 * x .+= dx (:79) and normalized_norm(residue) (:51,69). */
int mfb_initialize_dx(mfb_ctx *ctx, double dt, const double *gamma_params, int n_gamma);
int mfb_update_x_star(mfb_ctx *ctx, const double *alpha_params, int n_alpha);
int mfb_update_dx(mfb_ctx *ctx, const double *beta_params, int n_beta, double sign);
int mfb_commit_step(mfb_ctx *ctx);
int mfb_residue_norm(mfb_ctx *ctx, double *normalized_norm);

/* ---- result hand-off -------------------------------------------------------------------------
 * dessemble_X! + write_VTK (src/solver/03_GlobalAssembly.jl:63-75, src/mesh/unstructured_mesh/5_VTK.jl:7-157) in one call:
 * legacy ASCII unstructured grid with the reference's section order. cell_type / el_cp_outer_id are the element type's VTK
 * id and node permutation (5_VTK.jl:26-118; 1-based local node ids); symbol k is variable sym_var[k] (0-based position in
 * basic_vars) at time level sym_level[k]; coordinates are (x + shift) * scale, shift = the 3 variables starting at
 * shift_var (level 0) or none for shift_var < 0. Numbers are printed shortest-round-trip. */
int mfb_write_vtk(mfb_ctx *ctx, const char *path, int cell_type, int cell_size, const int32_t *el_cp_outer_id,
                  int n_syms, const char *const *sym_names, const int32_t *sym_var, const int32_t *sym_level,
                  double scale, int shift_var);

/* ---- multi-GPU: element-block partitioning (new relative to the single-GPU reference, README.md:4) -----------
 * One process per GPU. Each rank calls mfb_mesh_set with ITS block of elements and the nodes they touch (local
 * numbering); matrices/residuals stay unassembled across ranks and the library completes every SpMV result,
 * residual and Jacobi diagonal with an interface exchange-add over NCCL, and every reduction with one allreduce.
 *   mfb_comm_unique_id: rank 0 obtains the 128-byte NCCL id and hands it to the other ranks out of band.
 *   mfb_interface_set : neighbour ranks (ascending); offsets [n_neighbors+1] into shared_nodes; shared_nodes =
 *                       local node IDs (1-based) shared with each neighbour, ordered identically on both sides
 *                       (by global ID); owned [N] = 1 where this rank owns the node (lowest sharing rank);
 *                       global_ids [N] = global node ID (1-based) of each local node. */
int mfb_comm_unique_id(void *id128);
int mfb_comm_init(mfb_ctx *ctx, int rank, int n_ranks, const void *id128);
int mfb_interface_set(mfb_ctx *ctx, int n_neighbors, const int32_t *neighbor_ranks, const int64_t *offsets,
                      const int32_t *shared_nodes, const uint8_t *owned, const int64_t *global_ids);

#ifdef __cplusplus
}
#endif
#endif /* METAFEM_B200_H */
