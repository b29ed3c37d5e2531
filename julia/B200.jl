# B200.jl -- the Julia side of the drop-in boundary: MetaFEM.jl's hot path re-pointed at libmetafem_b200.so.
#
# Install: copy this file and parse_Term2CUDA.jl to src/b200/ and add `include("b200/B200.jl")` after the solver includes of
# src/MetaFEM.jl (:60-66). Every exported name and signature of the reference is kept; the methods below REPLACE the
# reference's bodies (same function objects, so example scripts run unchanged):
#   assemble_Global_Variables!   src/solver/03_GlobalAssembly.jl:6-37      -> mfb_mesh_set / mfb_facets_set / mfb_pattern_build
#   update_Mesh                  src/mesh/unstructured_mesh/2_Interface.jl:98-108  -> no-op (geometry lives in the element kernels)
#   compile_Updater_GPU          src/solver/05_CodeGenerator.jl:265-291    -> emit_CUDA + mfb_kernel_compile (NVRTC, sm_100a)
#   iterative_Solve!             src/solver/linear_solver/02_Preconditioner.jl:32-76 -> mfb_krylov_solve_ex
#   update_OneStep! helpers      src/solver/04_Time_Domain.jl:20-51,79     -> mfb_initialize_dx ... mfb_residue_norm
#   assemble_X! / dessemble_X!   src/solver/03_GlobalAssembly.jl:44-75     -> mfb_vector_set / mfb_vector_get
#   write_VTK                    src/mesh/unstructured_mesh/5_VTK.jl:7-157 -> mfb_write_vtk
# C prototypes: include/metafem_b200.h. The Python mirror metafem.jl_b200/api.py issues the same call sequence and is what the
# test-suite executes (Julia is not installed in the library's build environment; this file is delivered for review).

const LIBMFB = get(ENV, "MFB_LIB", "libmetafem_b200.so")
include("parse_Term2CUDA.jl")

const MFB_VEC_X, MFB_VEC_DX, MFB_VEC_X_STAR, MFB_VEC_RESIDUE = Cint(0), Cint(1), Cint(2), Cint(3)
const MFB_MAT_K_LINEAR, MFB_MAT_K_TOTAL = Cint(0), Cint(1)

mutable struct MfbCtx
    h::Ptr{Cvoid}
    qp_calls::Vector{QpCall}
    n_q::Int
end

function mfb_check(ctx::MfbCtx, rc::Integer, what::String)
    rc < 0 && error("$what: ", unsafe_string(ccall((:mfb_last_error, LIBMFB), Cstring, (Ptr{Cvoid},), ctx.h)))
    rc == 1 && println("$what: not converged")          # the reference only prints (02_Preconditioner.jl:66-68)
    return rc
end

function MfbCtx(device::Integer = 0)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    ccall((:mfb_create, LIBMFB), Cint, (Ref{Ptr{Cvoid}}, Cint), r, device) == 0 || error("mfb_create failed: no usable CUDA device (there is no CPU fallback)")
    ctx = MfbCtx(r[], QpCall[], 0)
    finalizer(c -> ccall((:mfb_destroy, LIBMFB), Cint, (Ptr{Cvoid},), c.h), ctx)
    return ctx
end

# one context per FEM_Domain (FEM_Domain has no spare field in the reference: keep a side table keyed by object identity)
const MFB_CONTEXTS = IdDict{Any, MfbCtx}()
b200(fem_domain::FEM_Domain) = get!(() -> MfbCtx(CUDA.deviceid(CUDA.device())), MFB_CONTEXTS, fem_domain)

# ---- C structs ------------------------------------------------------------------------------------------------------------
struct MfbBlockDesc                # mfb_block_desc
    kind::Int32
    bg_ID::Int32
    linear_kernel::Cstring
    nonlinear_kernel::Cstring
    n_cp_vars::Int32
    cp_var_names::Ptr{Cstring}
    n_globals::Int32
    global_names::Ptr{Cstring}
    threads_per_block::Int32
    smem_bytes::Int32
    has_nonlinear_K::Int32
    eval_kernel::Cstring
    n_qp_in::Int32
    qp_in_names::Ptr{Cstring}
    n_qp_out::Int32
    qp_out_names::Ptr{Cstring}
end

struct MfbSolveInfo                # mfb_solve_info
    passes::Int32
    iterations::Int32
    spmv_count::Int32
    converged::Int32
    residual::Float64
    initial_residual::Float64
end
MfbSolveInfo() = MfbSolveInfo(0, 0, 0, 0, 0.0, 0.0)

struct MfbJ2Params                 # mfb_j2_params
    lambda::Float64
    mu::Float64
    Eb::Float64
    Ep::Float64
    f_res::Float64
end

# ---- assemble_Global_Variables! (03_GlobalAssembly.jl:6-37) ------------------------------------------------------------------
function assemble_Global_Variables!(; fem_domain::FEM_Domain{ArrayType}) where {ArrayType}
    length(fem_domain.workpieces) == 1 || error("libmetafem_b200: one workpiece per FEM_Domain (the reference's multi-workpiece path has no cpID shift, 03_GlobalAssembly.jl:150)")
    wp = fem_domain.workpieces[1]
    mesh, sp, la = wp.mesh, wp.element_space, wp.local_assembly
    ctx = b200(fem_domain)
    elIDs = findall(mesh.elements.is_occupied)
    cp = mesh.elements.controlpoint_IDs[:, elIDs]
    cpIDs = findall(mesh.controlpoints.is_occupied)
    N = mesh.variable_size = length(cpIDs)
    N == size(mesh.controlpoints.is_occupied, 1) || error("libmetafem_b200: control-point table with holes (global_cpID != table slot)")
    mesh.controlpoints.global_cpID[cpIDs] .= 1:N
    mesh.elements.global_cpIDs[:, elIDs] .= cp
    ctx.n_q = sp.itg_func_num
    mfb_check(ctx, ccall((:mfb_mesh_set, LIBMFB), Cint,
        (Ptr{Cvoid}, Cint, Int64, Int64, Cint, CuPtr{Int32}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}),
        ctx.h, size(cp, 1), size(cp, 2), N, sp.itg_func_num, cp, mesh.controlpoints.x1, mesh.controlpoints.x2, mesh.controlpoints.x3,
        sp.ref_itp_vals, sp.itg_weight), "mfb_mesh_set")
    # facets: bdy_ref_itp_vals / bdy_itg_weights / bdy_tangent_directions are one array per local face -> one trailing face dimension
    fIDs = findall(mesh.facets.is_occupied)
    mfb_check(ctx, ccall((:mfb_facets_set, LIBMFB), Cint,
        (Ptr{Cvoid}, Cint, Cint, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Int64, CuPtr{Int32}, CuPtr{Int32}),
        ctx.h, length(sp.bdy_ref_itp_vals), sp.bdy_itg_func_num, cat(sp.bdy_ref_itp_vals...; dims = 6), hcat(sp.bdy_itg_weights...),
        cat(sp.bdy_tangent_directions...; dims = 4), length(fIDs), mesh.facets.element_ID[fIDs], mesh.facets.element_eindex[fIDs]), "mfb_facets_set")
    for (bg_ID, bg_fIDs) in pairs(mesh.bg_fIDs)
        mfb_check(ctx, ccall((:mfb_boundary_group_set, LIBMFB), Cint, (Ptr{Cvoid}, Cint, Int64, CuPtr{Int32}), ctx.h, bg_ID, length(bg_fIDs), bg_fIDs), "mfb_boundary_group_set")
    end
    # sparse_mapping: (dual_pos, base_pos) -> block number, blocks in number order (02_LocalAssembly.jl:70-74,104-105)
    pairs_sorted = sort(collect(la.sparse_mapping); by = last)
    mapping = Int32[k[i] for i = 1:2, (k, _) in pairs_sorted]                    # [2, n_blocks]
    nnz, unit = Ref{Int64}(0), Ref{Int64}(0)
    L = get_MaxTimeSteps(wp)
    mfb_check(ctx, ccall((:mfb_pattern_build, LIBMFB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int32}, Ref{Int64}, Ref{Int64}),
        ctx.h, length(la.basic_vars), L, size(mapping, 2), mapping, nnz, unit), "mfb_pattern_build")
    la.sparse_unitsize = unit[]
    la.sparse_entry_ID = nnz[]
    gf = fem_domain.globalfield
    gf.basicfield_size = length(la.basic_vars) * N
    gf.max_time_level = L
    assemble_X!(fem_domain.workpieces, gf; fem_domain = fem_domain)
    return nothing
end

# geometry at the quadrature points is evaluated inside the element kernels: nothing to tabulate
update_Mesh(dim::Integer, wp::WorkPiece, space::Classical_Discretization) = nothing

# K_I / K_J / K_J_ptr / K_val_ids in the reference's layout, for scripts that inspect them
function b200_pattern(fem_domain::FEM_Domain)
    ctx, gf = b200(fem_domain), fem_domain.globalfield
    nnz = Int(fem_domain.workpieces[1].local_assembly.sparse_entry_ID)
    K_I, K_J, K_val_ids = CUDA.zeros(Int32, nnz), CUDA.zeros(Int32, nnz), CUDA.zeros(Int32, nnz)
    K_J_ptr = CUDA.zeros(Int32, gf.basicfield_size + 1)
    mfb_check(ctx, ccall((:mfb_pattern_get, LIBMFB), Cint, (Ptr{Cvoid}, CuPtr{Int32}, CuPtr{Int32}, CuPtr{Int32}, CuPtr{Int32}),
                         ctx.h, K_I, K_J, K_J_ptr, K_val_ids), "mfb_pattern_get")
    return K_I, K_J, K_J_ptr, K_val_ids
end

# ---- assemble_X! / dessemble_X! (03_GlobalAssembly.jl:44-75) -------------------------------------------------------------------
function assemble_X!(workpieces::Vector{WorkPiece}, globalfield::GlobalField; fem_domain::FEM_Domain)
    ctx = b200(fem_domain)
    wp = workpieces[1]
    N, n = wp.mesh.variable_size, globalfield.basicfield_size
    x = CUDA.zeros(FEM_Float, (globalfield.max_time_level + 1) * n)
    cpts = wp.mesh.controlpoints
    cpIDs = findall(cpts.is_occupied)
    for (local_sym, basic_pos, td_order) in wp.local_assembly.local_innervar_infos
        s = basic_pos * N + td_order * n
        x[(s + 1):(s + N)] .= getproperty(cpts, local_sym)[cpIDs]
    end
    mfb_check(ctx, ccall((:mfb_vector_set, LIBMFB), Cint, (Ptr{Cvoid}, Cint, CuPtr{Float64}, Int64), ctx.h, MFB_VEC_X, x, length(x)), "assemble_X!")
end

function dessemble_X!(workpieces::Vector{WorkPiece}, globalfield::GlobalField; fem_domain::FEM_Domain)
    ctx = b200(fem_domain)
    wp = workpieces[1]
    N, n = wp.mesh.variable_size, globalfield.basicfield_size
    x = CUDA.zeros(FEM_Float, (globalfield.max_time_level + 1) * n)
    mfb_check(ctx, ccall((:mfb_vector_get, LIBMFB), Cint, (Ptr{Cvoid}, Cint, CuPtr{Float64}, Int64), ctx.h, MFB_VEC_X, x, length(x)), "dessemble_X!")
    cpts = wp.mesh.controlpoints
    cpIDs = findall(cpts.is_occupied)
    for (local_sym, basic_pos, td_order) in wp.local_assembly.local_innervar_infos
        s = basic_pos * N + td_order * n
        getproperty(cpts, local_sym)[cpIDs] .= x[(s + 1):(s + N)]
    end
end

# ---- compile_Updater_GPU (05_CodeGenerator.jl:265-291) --------------------------------------------------------------------------
# controlpoints.<sym> and physics.global_vars are read at CALL time in the reference (05_CodeGenerator.jl:21-35)
function sync_fields!(ctx::MfbCtx, fem_domain::FEM_Domain, blocks::Vector{BlockText})
    wp = fem_domain.workpieces[1]
    cpts = wp.mesh.controlpoints
    cpIDs = findall(cpts.is_occupied)
    for sym in unique(vcat([b.form.fields for b in blocks]...))
        v = getproperty(cpts, sym)[cpIDs]
        mfb_check(ctx, ccall((:mfb_field_set, LIBMFB), Cint, (Ptr{Cvoid}, Cstring, CuPtr{Float64}), ctx.h, string(sym), v), "mfb_field_set($sym)")
    end
    for sym in unique(vcat([b.form.globals for b in blocks]...))
        (sym == :t || sym == :dt) && continue
        ccall((:mfb_global_set, LIBMFB), Cint, (Ptr{Cvoid}, Cstring, Float64), ctx.h, string(sym), wp.physics.global_vars[sym])
    end
end

# library-owned [n_q, n_el] array (reference element order) as a CuArray view, no copy
function qp_array(ctx::MfbCtx, name::String)
    p, n = Ref{CuPtr{Float64}}(), Ref{Int64}(0)
    mfb_check(ctx, ccall((:mfb_qp_array, LIBMFB), Cint, (Ptr{Cvoid}, Cstring, Ref{CuPtr{Float64}}, Ref{Int64}), ctx.h, name, p, n), "mfb_qp_array($name)")
    return unsafe_wrap(CuArray, p[], (ctx.n_q, n[] ÷ ctx.n_q))
end

function compile_Updater_GPU(; domain_ID::Integer, fem_domain::FEM_Domain)
    ctx = b200(fem_domain)
    cuda_src, blocks = emit_CUDA(fem_domain)
    # descriptors: every string must stay rooted while the C call runs
    GC.@preserve cuda_src blocks begin
        keep = Any[]
        cstr(s::Nothing) = Cstring(C_NULL)
        cstr(s::String) = (push!(keep, s); Base.unsafe_convert(Cstring, s))
        function cstr_list(v::Vector{String})
            isempty(v) && return Ptr{Cstring}(C_NULL)
            ptrs = Cstring[cstr(s) for s in v]
            push!(keep, ptrs)
            return pointer(ptrs)
        end
        descs = MfbBlockDesc[MfbBlockDesc(b.kind, b.bg_ID, cstr(b.linear_kernel), cstr(b.nonlinear_kernel),
                                          length(b.form.fields), cstr_list(string.(b.form.fields)),
                                          length(b.form.globals), cstr_list(string.(b.form.globals)),
                                          b.form.tpb, b.form.smem, b.form.has_K ? 1 : 0, cstr(b.eval_kernel),
                                          length(b.form.qp_in), cstr_list(b.form.qp_in), length(b.form.qp_out), cstr_list(b.form.qp_out))
                              for b in blocks]
        GC.@preserve keep descs mfb_check(ctx, ccall((:mfb_kernel_compile, LIBMFB), Cint, (Ptr{Cvoid}, Cstring, Cint, Ptr{MfbBlockDesc}),
                                                      ctx.h, cuda_src, length(descs), descs), "mfb_kernel_compile")
    end
    ctx.qp_calls = vcat([b.qp_calls for b in blocks]...)

    fem_domain.K_linear_func = function (time_discretization::GeneralAlpha; fem_domain::FEM_Domain)
        sync_fields!(ctx, fem_domain, blocks)
        kp = Vector{Float64}(time_discretization.K_params)
        mfb_check(ctx, ccall((:mfb_assemble_linear, LIBMFB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), ctx.h, kp, length(kp)), "update_K_Linear_$domain_ID")
    end
    fem_domain.K_nonlinear_func = function (time_discretization::GeneralAlpha; fem_domain::FEM_Domain)
        sync_fields!(ctx, fem_domain, blocks)
        gf = fem_domain.globalfield
        if ~isempty(ctx.qp_calls)
            # two-phase update around the user's quadrature-point callbacks (08_Tensor.jl:175-183,210): argument arrays ->
            # Main.<func> on whole [n_q, n_el] CuArrays, unchanged -> outputs into the arrays the residual kernel reads
            mfb_check(ctx, ccall((:mfb_eval_qp_args, LIBMFB), Cint, (Ptr{Cvoid}, Float64, Float64), ctx.h, gf.t, gf.dt), "mfb_eval_qp_args")
            for call in ctx.qp_calls
                args = [qp_array(ctx, nm) for nm in call.arg_names]
                outs = getfield(Main, call.func)(args...)
                outs isa Tuple || (outs = (outs,))
                for (sym, o) in zip(call.outs, outs)
                    copyto!(qp_array(ctx, string(sym)), o)
                end
            end
        end
        kp = Vector{Float64}(time_discretization.K_params)
        mfb_check(ctx, ccall((:mfb_assemble_nonlinear, LIBMFB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Float64, Float64),
                             ctx.h, kp, length(kp), gf.t, gf.dt), "update_K_NonLinear_$domain_ID")
    end
    return cuda_src          # the reference returns the generated Exprs; here: the generated CUDA C
end

# ---- iterative_Solve! (02_Preconditioner.jl:32-76) -------------------------------------------------------------------------------
const MFB_METHOD = IdDict{Function, Cint}(idrs! => 0, bicgstabl_GS! => 1, bicgstabl! => 2, gmres! => 3, cgs! => 4, cgs2! => 5,
                                          tfqmr! => 6, lsqr! => 7, idrs_original! => 8)
struct DeviceDelta; ctx::MfbCtx; end          # the solve result stays on the device; update_dx! dispatches on it
Base.:-(d::DeviceDelta) = d                   # update_OneStep! passes `.- delta_x` (04_Time_Domain.jl:77): the sign is applied in update_dx!
Base.broadcasted(::typeof(-), d::DeviceDelta) = d

function iterative_Solve!(globalfield::GlobalField; Sv_func!::Function = idrs!, Pr_func!::Function = Pr_Jacobi!, Pl_func::Function = Identity,
                          max_pass::Integer = 4, maxiter::Integer = 2000, s::Integer = 4, checkiter::Integer = 200,
                          normalized_by_column::Bool = false, normalized_by_row::Bool = false, fem_domain::FEM_Domain = CURRENT_DOMAIN[])
    ctx = b200(fem_domain)
    haskey(MFB_METHOD, Sv_func!) || error("$(Sv_func!) is not provided by libmetafem_b200")
    pr = Pr_func! === Pr_Jacobi! ? (normalized_by_column ? 1 : 0) : 2                     # Identity -> 2
    pl = Pl_func === Identity ? 0 : Pl_func === Pl_Jacobi ? (normalized_by_row ? 2 : 1) : Pl_func === Pl_ILU ? 3 : error("unknown Pl_func")
    info = Ref(MfbSolveInfo())
    mfb_check(ctx, ccall((:mfb_krylov_solve_ex, LIBMFB), Cint,
        (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Float64, UInt64, Cint, Cint, Cint, CuPtr{Float64}, Ref{MfbSolveInfo}),
        ctx.h, MFB_METHOD[Sv_func!], s, maxiter, max_pass, globalfield.converge_tol, rand(UInt64), pr, pl, checkiter, CU_NULL, info), "iterative_Solve!")
    println("pass $(info[].passes) with res = $(info[].residual) iter = $(info[].iterations).")   # 02_Preconditioner.jl:55,67
    return DeviceDelta(ctx)
end
# scripts install `x -> iterative_Solve!(x; ...)` with x = globalfield (static_Neo_Hookean.jl:80): the closure has no handle on
# the domain, so update_OneStep! records the one it is stepping
const CURRENT_DOMAIN = Ref{Any}(nothing)

# ---- update_OneStep! helpers (04_Time_Domain.jl:20-51,79) -----------------------------------------------------------------------
function initialize_dx!(globalfield::GlobalField, gamma_params; fem_domain::FEM_Domain = CURRENT_DOMAIN[])
    g = Vector{Float64}(collect(gamma_params))
    mfb_check(b200(fem_domain), ccall((:mfb_initialize_dx, LIBMFB), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}, Cint), b200(fem_domain).h, globalfield.dt, g, length(g)), "initialize_dx!")
end
function update_x_star!(globalfield::GlobalField, alpha_params; fem_domain::FEM_Domain = CURRENT_DOMAIN[])
    a = Vector{Float64}(collect(alpha_params))
    mfb_check(b200(fem_domain), ccall((:mfb_update_x_star, LIBMFB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), b200(fem_domain).h, a, length(a)), "update_x_star!")
end
function update_dx!(globalfield::GlobalField, delta::DeviceDelta, beta_params)
    b = Vector{Float64}(collect(beta_params))
    mfb_check(delta.ctx, ccall((:mfb_update_dx, LIBMFB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Float64), delta.ctx.h, b, length(b), -1.0), "update_dx!")
end
function b200_residue_norm(fem_domain::FEM_Domain)
    r = Ref{Float64}(0.0)
    mfb_check(b200(fem_domain), ccall((:mfb_residue_norm, LIBMFB), Cint, (Ptr{Cvoid}, Ref{Float64}), b200(fem_domain).h, r), "normalized_norm(residue)")
    return r[]
end

function update_OneStep!(time_discretization::GeneralAlpha; max_iter::Integer = 4, fem_domain::FEM_Domain)
    CURRENT_DOMAIN[] = fem_domain
    @Takeout (globalfield, K_linear_func, K_nonlinear_func, linear_solver) FROM fem_domain
    @Takeout (alpha_params, gamma_params) FROM time_discretization
    update_Time!(globalfield, time_discretization)                                   # host scalars: unchanged (:10-18)
    initialize_dx!(globalfield, gamma_params; fem_domain = fem_domain)
    K_linear_func(time_discretization; fem_domain = fem_domain)
    counter = -1
    while true
        update_x_star!(globalfield, alpha_params; fem_domain = fem_domain)
        K_nonlinear_func(time_discretization; fem_domain = fem_domain)
        res = b200_residue_norm(fem_domain)
        counter += 1
        println("step $counter residue = $res")
        ((res < globalfield.converge_tol) || (counter > max_iter)) && break
        delta_x = linear_solver(globalfield)
        update_dx!(globalfield, .-delta_x, time_discretization.beta_params)
    end
    mfb_check(b200(fem_domain), ccall((:mfb_commit_step, LIBMFB), Cint, (Ptr{Cvoid},), b200(fem_domain).h), "x .+= dx")
end

# ---- write_VTK (5_VTK.jl:7-157): dessemble_X! included, the solution is read from the device-resident x ----------------------------
function write_VTK(fname::String, wp::WorkPiece; scale = 1., shift_sym = :none, fem_domain::FEM_Domain = CURRENT_DOMAIN[])
    ctx = b200(fem_domain)
    cell_type, el_cp_outer_id = vtk_cell_table(wp.element_space.element_attributes)   # the (cell type, node permutation) table of 5_VTK.jl:26-118
    la = wp.local_assembly
    names = String[string(local_sym) for (local_sym, _, _) in la.local_innervar_infos]
    var = Int32[basic_pos for (_, basic_pos, _) in la.local_innervar_infos]
    lev = Int32[td_order for (_, _, td_order) in la.local_innervar_infos]
    shift = shift_sym == :none ? -1 : findfirst(==(Symbol(shift_sym, 1)), la.basic_vars) - 1
    mfb_check(ctx, ccall((:mfb_write_vtk, LIBMFB), Cint,
        (Ptr{Cvoid}, Cstring, Cint, Cint, Ptr{Int32}, Cint, Ptr{Cstring}, Ptr{Int32}, Ptr{Int32}, Float64, Cint),
        ctx.h, fname, cell_type, length(el_cp_outer_id), Int32.(el_cp_outer_id), length(names), names, var, lev, Float64(scale), shift), "write_VTK")
end

# ---- built-in J2 return map (optional fast path; examples/hypo_elastic_plasticity/J2Plasticity.jl:76-198) --------------------------
# The example's MaterialState callable (~40 broadcast kernels and a findall per call) keeps working through the generic
# callback path above. To use the library's one-kernel return map instead, define in the script:
#   state = B200J2State(fem_domain, "j2"; Y_initial, λ, μ, Eb, Ep, f_res)
#   strain_updater(e...) = state(e...)            # same name the weak form refers to
#   update_States!(state)                          # after each converged pseudo-time step (:274)
struct B200J2State
    ctx::MfbCtx
    prefix::String
    params::MfbJ2Params
    n_yielded::Base.RefValue{Int64}
end
function B200J2State(fem_domain::FEM_Domain, prefix::String; Y_initial, λ, μ, Eb, Ep, f_res, func::Symbol = :strain_updater)
    ctx = b200(fem_domain)
    call = ctx.qp_calls[findfirst(c -> c.func == func, ctx.qp_calls)]
    e_names, ep_names = call.arg_names, string.(call.outs)
    mfb_check(ctx, ccall((:mfb_j2_init, LIBMFB), Cint, (Ptr{Cvoid}, Cstring, Float64, Ptr{Cstring}, Ptr{Cstring}), ctx.h, prefix, Y_initial, e_names, ep_names), "mfb_j2_init")
    return B200J2State(ctx, prefix, MfbJ2Params(λ, μ, Eb, Ep, f_res), Ref{Int64}(0))
end
function (st::B200J2State)(e...)
    mfb_check(st.ctx, ccall((:mfb_j2_iterate_stress, LIBMFB), Cint, (Ptr{Cvoid}, Cstring, Ref{MfbJ2Params}, Ref{Int64}), st.ctx.h, st.prefix, Ref(st.params), st.n_yielded), "iterate_stress!")
    call = st.ctx.qp_calls[1]
    return Tuple(qp_array(st.ctx, string(s)) for s in call.outs)       # already in place: copyto! onto itself is a no-op
end
update_States!(st::B200J2State) = mfb_check(st.ctx, ccall((:mfb_j2_update_states, LIBMFB), Cint, (Ptr{Cvoid}, Cstring), st.ctx.h, st.prefix), "update_States!")
