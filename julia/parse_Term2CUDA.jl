# parse_Term2CUDA.jl -- CUDA C emitter of MetaFEM's symbolic layer (north star: "src/symbolics gains a CUDA C emitter").
#
# Sibling of parse_Term2Expr! (src/symbolics/08_Tensor.jl:169-233): the same traversal of GroundTerm trees
# (FEM_Float | SymbolicWord | SymbolicTerm, src/symbolics/01_Types.jl:36-58), the same hoisting of defined tensors into
# temporaries and the same word names (word_To_TotalSym, src/symbolics/03_Word.jl:86), but it prints C expression strings
# instead of Julia Exprs. gen_Form_CUDA / emit_CUDA (below) are the siblings of gen_K_Linear_GPU / gen_Res_K_NonLinear_GPU /
# gen_CodeBody (src/solver/05_CodeGenerator.jl:52-258): per generated block they produce ONE `Form` struct on top of
# csrc/mfb_skeleton.cuh -- all terms of the block fused into one kernel -- and the mfb_block_desc the C ABI wants.
# The executable specification of the emitted text is metafem.jl_b200/emitter.py (same struct members, same tables);
# tests/test_emitter_golden.py pins that text. This file is included from src/b200/B200.jl AFTER src/solver.
#
# Julia is not available in the build environment of libmetafem_b200, so this file is delivered for review, not executed there.

# ---- identifiers and literals --------------------------------------------------------------------------------------------
# MetaFEM symbols are arbitrary Unicode (σˡ6, τᵇ, μ, λ ...); C identifiers are not. Non [A-Za-z0-9_] code points -> _uXXXX.
function c_ident(sym::Symbol)
    io = IOBuffer()
    for ch in string(sym)
        if ('a' <= ch <= 'z') || ('A' <= ch <= 'Z') || ('0' <= ch <= '9') || ch == '_'
            print(io, ch)
        else
            print(io, "_u", uppercase(string(UInt32(ch); base = 16, pad = 4)))
        end
    end
    s = String(take!(io))
    return ('0' <= s[1] <= '9') ? "_" * s : s
end

# shortest round-trip decimal of a Float64, always with a '.' or an exponent so that C parses a double
function c_literal(x::FEM_Float)
    s = repr(Float64(x))
    s == "Inf" && return "(1.0/0.0)"
    s == "-Inf" && return "(-1.0/0.0)"
    return (occursin('.', s) || occursin('e', s)) ? s : s * ".0"
end

const C_FUNCTIONS = Dict{Symbol, String}(:log => "log", :exp => "exp", :sqrt => "sqrt", :sin => "sin", :cos => "cos", :tan => "tan",
                                         :tanh => "tanh", :abs => "fabs", :atan => "atan")

# One user callback at the quadrature points: `ep{i,j} = strain_updater(e{1,1}, ...)` (J2Plasticity.jl:55). The reference emits
# `(ep1, ..., ep6) = Main.strain_updater(e1_1, ...)` on whole [n_q, n_el] arrays (08_Tensor.jl:175-183, 210); here the call
# becomes a two-phase update (mfb_eval_qp_args -> callback -> mfb_assemble_nonlinear) or a built-in fused into the kernel.
struct QpCall
    func::Symbol
    args::Vector{String}          # C expressions of the arguments
    arg_names::Vector{String}     # "<func>_arg<k>": library-owned [n_q, n_el] arrays filled by the argument kernel
    outs::Vector{Symbol}          # INTEGRATION_POINT_VAR output symbols, read by the residual as external words
end

mutable struct CudaEmitState
    code::Vector{String}          # hoisted temporaries, "const double sym = expr;"
    declared::Set{Symbol}
    qp_calls::Vector{QpCall}
    words::Set{Symbol}            # every INTERNAL/EXTERNAL word symbol met (they must be declared by the block header)
end
CudaEmitState() = CudaEmitState(String[], Set{Symbol}(), QpCall[], Set{Symbol}())

_parse_Term2CUDA!(st::CudaEmitState, tb::TensorTable, this_num::FEM_Float) = c_literal(this_num)

function _parse_Term2CUDA!(st::CudaEmitState, tb::TensorTable, this_word::SymbolicWord)
    totalsym = word_To_TotalSym(tb.dim, this_word)
    if ~(totalsym in st.declared)
        attributes = get_VarAttribute(this_word)
        if (:INTERNAL_VAR in attributes) || (:EXTERNAL_VAR in attributes)
            if (:INTEGRATION_POINT_VAR in attributes) && (this_word.base_variable != :n)
                # defined by a user function of other words: all related output symbols at once (08_Tensor.jl:175-183)
                all_syms = generates_All_Related_ITG_Symbols(tb, this_word, attributes)
                _, raw_def = deepcopy(DEFINITION_TABLE[this_word.base_variable])
                def_term = propagate_Symbol(tb, raw_def)
                (def_term isa SymbolicTerm && ~haskey(C_FUNCTIONS, def_term.operation) && ~Base.isoperator(def_term.operation)) ||
                    error("INTEGRATION_POINT_VAR $(this_word.base_variable) must be defined by a user function call")
                args = String[_parse_Term2CUDA!(st, tb, subterm) for subterm in def_term.subterms]
                func = def_term.operation
                push!(st.qp_calls, QpCall(func, args, ["$(func)_arg$k" for k = 1:length(args)], all_syms))
                union!(st.declared, all_syms)
                union!(st.words, all_syms)
            else
                push!(st.declared, totalsym)
                push!(st.words, totalsym)
            end
        else
            # a defined tensor component: hoisted into a temporary exactly like parse_Term2Expr! does (08_Tensor.jl:191-199)
            defs = parse_Term2CUDA!(st, tb, evaluate_Tensor(tb, this_word))
            push!(st.code, "const double $(c_ident(totalsym)) = $(join(defs, " + "));")
            push!(st.declared, totalsym)
        end
    end
    return c_ident(totalsym)
end

function _parse_Term2CUDA!(st::CudaEmitState, tb::TensorTable, this_term::SymbolicTerm)
    op = this_term.operation
    args = String[_parse_Term2CUDA!(st, tb, subterm) for subterm in this_term.subterms]
    if op == :+ || op == :*
        # n-ary, evaluated left to right like the broadcast of the reference (sub-terms are ordered by Julia's hash, 04_Term.jl:57,77)
        return "(" * join(args, op == :+ ? " + " : " * ") * ")"
    elseif op == :^
        # all literals are Float64, so the reference goes through pow (04_Term.jl:9,104); small integer powers are expanded
        base, ex = args
        e = this_term.subterms[2]
        if e isa Number
            e == 2.0 && return "((" * base * ")*(" * base * "))"
            e == -1.0 && return "(1.0/(" * base * "))"
            e == -2.0 && return "(1.0/((" * base * ")*(" * base * ")))"
        end
        return "pow(" * base * ", " * ex * ")"
    elseif haskey(C_FUNCTIONS, op)
        return C_FUNCTIONS[op] * "(" * join(args, ", ") * ")"
    else
        error("function $op inside an expression: user functions are supported as definitions of INTEGRATION_POINT_VAR tensors only")
    end
end

"""
    parse_Term2CUDA!(st, tb, term) -> Vector{String}

C expressions whose sum is `term` (the reference splits long sums into chunks of <= 64 words that are accumulated with `+=`,
08_Tensor.jl:214-233; one C expression per chunk keeps the same association).
"""
function parse_Term2CUDA!(st::CudaEmitState, tb::TensorTable, this_term)
    inlined_term = propagate_Symbol(tb, this_term)
    if inlined_term isa SymbolicTerm && inlined_term.operation == :+
        word_nums = count_Words.(inlined_term.subterms)
        limit, counter = 64, 0
        buffer, defs = GroundTerm[], String[]
        for (id, word_num) in enumerate(word_nums)
            if (counter += word_num) > limit
                counter = 0
                push!(defs, _parse_Term2CUDA!(st, tb, ⨁(buffer)))
                empty!(buffer)
            end
            push!(buffer, inlined_term.subterms[id])
        end
        return push!(defs, _parse_Term2CUDA!(st, tb, ⨁(buffer)))
    else
        return String[_parse_Term2CUDA!(st, tb, inlined_term)]
    end
end

# ---- one Form struct per generated block (mirrors metafem.jl_b200/emitter.py::_form) ------------------------------------------
sd_slot(sd_order) = isempty(sd_order) ? 0 : (length(sd_order) == 1 ? Int(sd_order[1]) : error("spatial derivative order > 1 is outside max_sd_order = 1"))
c_table(fn, vals) = "  __device__ static constexpr int $fn(int i) { constexpr int t[] = {$(join(isempty(vals) ? [0] : vals, ", "))}; return t[i]; }"
align16(n) = cld(n, 16) * 16

# tangent tiling: one lane owns the rows (a, dp, bp = 0..NV-1) and NTC columns of the element matrix (emitter.py::_tile)
function tangent_tile(n_a, nv; max_acc = 60)
    cg = 1
    ntc = 0
    while true
        ntc = cld(n_a, cg)
        ntc += ntc & 1
        (nv * ntc <= max_acc || ntc == 2) && break
        cg += 1
    end
    tiles = n_a * nv * cg
    w = cld(tiles, 32)
    return (NTC = ntc, CG = cg, W = w, LPW = cld(tiles, w))
end

struct FormText
    body::String
    fields::Vector{Symbol}        # CONTROLPOINT_VAR local symbols (mfb_field_set names)
    globals::Vector{Symbol}       # GLOBAL_VAR symbols (mfb_global_set names; :t and :dt are filled by the library)
    smem::Int
    has_K::Bool
    tpb::Int
    qp_in::Vector{String}
    qp_out::Vector{String}
end

"""
    gen_Form_CUDA(name, tb, asm_wf, basic_n, max_time_level, n_a, n_q; is_boundary, linear, evalk)

The C++ `Form` of one block: `linear` -> K_linear kernel (gen_K_Linear_GPU, 05_CodeGenerator.jl:52-91), `evalk` -> the
argument kernel of the block's quadrature-point callbacks, else residue + K_total (gen_Res_K_NonLinear_GPU, :93-154).
"""
function gen_Form_CUDA(name::String, tb::TensorTable, asm_wf::AssembleWeakform, nv::Integer, max_time_level::Integer, n_a::Integer, n_q::Integer;
                       is_boundary::Bool, linear::Bool, evalk::Bool = false)
    dim = tb.dim
    st = CudaEmitState()
    terms = linear ? asm_wf.linear_gradients : asm_wf.nonlinear_gradients
    residues = linear ? AssembleBilinear[] : asm_wf.residues
    inner = linear ? InnervarInfo[] : sort(asm_wf.innervar_infos)
    ext = sort(linear ? asm_wf.linear_extervar_infos : asm_wf.extervar_infos)
    # expressions first: they tell which temporaries and callbacks the block needs
    res_lines, k_lines = String[], String[]
    dslots = sort(unique([sd_slot(b.dual_info[3]) for b in terms]))
    bslots = sort(unique([sd_slot(b.derivative_info[3]) for b in terms]))
    nsd, ks = length(dslots), length(bslots)
    for b in residues
        _, _, dual_sd, dual_pos = b.dual_info
        for ex in parse_Term2CUDA!(st, tb, b.base_term)
            push!(res_lines, "    R[$(dual_pos * 4 + sd_slot(dual_sd))] += $ex;")
        end
    end
    for b in terms
        _, _, dual_sd, dual_pos = b.dual_info
        _, d_td, d_sd, d_pos = b.derivative_info
        idx = ((dual_pos * nsd + (findfirst(==(sd_slot(dual_sd)), dslots) - 1)) * nv + d_pos) * ks + (findfirst(==(sd_slot(d_sd)), bslots) - 1)
        for ex in parse_Term2CUDA!(st, tb, b.base_term)
            push!(k_lines, "    D[$idx] += ($ex) * A.Kp[$d_td];")       # K_params[derivative_td_order + 1] (:75,136)
        end
    end
    evalk && (empty!(res_lines); empty!(k_lines))
    cpw = [e for e in ext if :CONTROLPOINT_VAR in get_VarAttribute(e[3])]
    fields = sort(unique([e[2] for e in cpw]))
    globs = [e[1] for e in ext if :GLOBAL_VAR in get_VarAttribute(e[3])]
    normals = [e for e in ext if e[3] == :n]
    qp_words = (linear || evalk) ? Symbol[] : vcat([c.outs for c in st.qp_calls]...)
    qpo = evalk ? vcat([collect(zip(c.args, c.arg_names)) for c in st.qp_calls]...) : Tuple{String, String}[]
    (is_boundary && (~isempty(qp_words) || ~isempty(qpo))) && error("INTEGRATION_POINT_VAR words are supported in domain blocks only")
    gslots = sort(unique(vcat(dslots, bslots, [sd_slot(b.dual_info[3]) for b in residues])))
    tl = isempty(terms) ? (NTC = 2, CG = 1, W = 1, LPW = 1) : tangent_tile(n_a, nv)
    tpb = 32 * max(isempty(terms) ? 2 : tl.W, 2)
    L1 = max_time_level + 1
    nd = nv * nsd * nv * ks
    io = String[]
    push!(io, "struct $name {")
    push!(io, "  static constexpr int NV = $nv, NA = $n_a, NQ = $n_q, L1 = $L1, BOUNDARY = $(Int(is_boundary)), " *
              "LINEAR = $(Int(linear)), NW = $(length(inner)), NCW = $(length(cpw)), NC = $(length(fields)), " *
              "HAS_RES = $(Int(~isempty(res_lines))), HAS_K = $(Int(~isempty(k_lines))), TPB = $tpb, " *
              "NSD = $nsd, KS = $ks, ND = $nd, NTC = $(tl.NTC), CG = $(tl.CG), W = $(tl.W), " *
              "LPW = $(tl.LPW), SMEM = @SMEM@, " *
              "EVAL = $(Int(evalk)), NQPI = $(length(qp_words)), NQPO = $(length(qpo)), NGS = $(length(gslots));")
    push!(io, c_table("gslot", [something(findfirst(==(sl), gslots), 0) - 1 for sl in 0:3]))
    push!(io, c_table("gslot_id", gslots))
    push!(io, c_table("dslot", dslots))
    push!(io, c_table("bslot", bslots))
    push!(io, c_table("wslot", [sd_slot(w[3]) for w in inner]))
    push!(io, c_table("wlev", [w[2] for w in inner]))
    push!(io, c_table("wpos", [w[4] for w in inner]))
    push!(io, c_table("cslot", [sd_slot(e[4]) for e in cpw]))
    push!(io, c_table("cfield", [findfirst(==(e[2]), fields) - 1 for e in cpw]))
    declare_words = function (lines)
        for (k, w) in enumerate(inner); push!(lines, "    const double $(c_ident(w[1])) = w[$(k - 1)];"); end
        for (k, e) in enumerate(cpw); push!(lines, "    const double $(c_ident(e[1])) = c[$(k - 1)];"); end
        for (k, g) in enumerate(globs); push!(lines, "    const double $(c_ident(g)) = A.glob[$(k - 1)];"); end
    end
    if evalk
        push!(io, "  __device__ static __forceinline__ void qp_eval(const double* w, const double* c, const MfbArgs& A, double* out) {")
        declare_words(io)
        for (k, (a, _)) in enumerate(qpo); push!(io, "    out[$(k - 1)] = $a;"); end
        push!(io, "  }")
    end
    push!(io, "  __device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, " *
              "const double* qv, size_t qidx, const MfbArgs& A, double* R, double* D) {")
    if ~evalk
        declare_words(io)
        for (k, s) in enumerate(qp_words); push!(io, "    const double $(c_ident(s)) = qv[$(k - 1)];"); end
        for e in normals; push!(io, "    const double $(c_ident(e[1])) = nrm[$(e[5][1] - 1)];"); end      # facets.normal_directions[:, c, facet]
        append!(io, ["    " * l for l in st.code])
        append!(io, res_lines)
        append!(io, k_lines)
    end
    push!(io, "  }")
    push!(io, "};")
    # shared-memory footprint of mfb::Smem<Form> (the skeleton static_asserts that this bound holds)
    has_K = ~isempty(k_lines)
    mrows = n_a * nv * nv
    nap = n_a + (n_a & 1)
    dpb = nsd * nv * ks
    nds = nd > 0 ? nv * (dpb + (dpb & 1)) : 1
    gd = align16(8 * (n_q * max(length(gslots), 1) * nap + n_q * nds))
    ke = 8 * (has_K ? mrows * n_a : 1)
    nvl = (linear ? 0 : L1 * nv) + length(fields)
    geo = align16(max(8 * n_q * (9 + 1 + 3), has_K ? 4 * n_a * n_a : 4))
    smem = align16(align16(max(gd, ke)) + geo + 8 * (n_q * 4 * max(nvl, 1) + n_a * 3 + L1 * n_a * nv + max(length(fields), 1) * n_a) + 4 * n_a)
    body = replace(join(io, "\n"), "@SMEM@" => string(smem))
    return FormText(body, fields, globs, smem, has_K, tpb, string.(qp_words), [n for (_, n) in qpo]), st.qp_calls
end

# resident blocks per SM to ask of __launch_bounds__ (emitter.py::_min_blocks)
min_blocks(tpb, smem; regs = 170) = max(1, min(232448 ÷ (smem + 1024), 65536 ÷ (tpb * regs), 32))

struct BlockText
    kind::Int32                   # 0 domain, 1 boundary group
    bg_ID::Int32
    linear_kernel::Union{String, Nothing}
    nonlinear_kernel::Union{String, Nothing}
    eval_kernel::Union{String, Nothing}
    form::FormText
    qp_calls::Vector{QpCall}
end

"""
    emit_CUDA(fem_domain) -> (cuda_source::String, blocks::Vector{BlockText})

Sibling of gen_CodeBody (05_CodeGenerator.jl:156-258) for the single-workpiece Classical_Discretization path: the domain
block, then one block per boundary group, each with up to three entry points on the skeleton.
"""
function emit_CUDA(fem_domain::FEM_Domain)
    tb = fem_domain.tensor_table
    wp = fem_domain.workpieces[1]
    sp, la = wp.element_space, wp.local_assembly
    nv, L = length(la.basic_vars), get_MaxTimeSteps(wp)
    n_a, n_q, n_qb = sp.itp_func_num, sp.itg_func_num, sp.bdy_itg_func_num
    src = String["#include \"mfb_skeleton.cuh\"", ""]
    blocks = BlockText[]
    todo = Any[(0, 0, la.assembled_weakform)]
    for bg_ID in sort(collect(keys(la.assembled_boundary_weakform_pairs)))
        push!(todo, (1, bg_ID, la.assembled_boundary_weakform_pairs[bg_ID]))
    end
    for (i, (kind, bg_ID, asm_wf)) in enumerate(todo)
        i0 = i - 1
        nq = kind == 1 ? n_qb : n_q
        variants = Tuple{String, Bool, Bool}[]
        isempty(asm_wf.linear_gradients) || push!(variants, ("lin", true, false))
        (isempty(asm_wf.residues) && isempty(asm_wf.nonlinear_gradients)) || push!(variants, ("nl", false, false))
        forms = Dict{String, Tuple{FormText, Vector{QpCall}}}()
        for (tag, lin, ev) in variants
            forms[tag] = gen_Form_CUDA("F_b$(i0)_$tag", tb, asm_wf, nv, L, n_a, nq; is_boundary = kind == 1, linear = lin, evalk = ev)
        end
        if haskey(forms, "nl") && ~isempty(forms["nl"][2])          # callbacks -> argument kernel (phase A of the two-phase update)
            forms["ev"] = gen_Form_CUDA("F_b$(i0)_ev", tb, asm_wf, nv, L, n_a, nq; is_boundary = kind == 1, linear = false, evalk = true)
            push!(variants, ("ev", false, true))
        end
        block_tpb = maximum([forms[t][1].tpb for (t, _, _) in variants]; init = 32)
        names = Dict{String, Union{String, Nothing}}("lin" => nothing, "nl" => nothing, "ev" => nothing)
        smem = 0
        for (tag, _, _) in variants
            ft = forms[tag][1]
            body = replace(ft.body, r"TPB = \d+" => "TPB = $block_tpb"; count = 1)
            push!(src, body)
            push!(src, "extern \"C\" __global__ void __launch_bounds__($block_tpb, $(min_blocks(block_tpb, ft.smem))) mfb_b$(i0)_$tag(const MfbArgs A) { mfb::assemble<F_b$(i0)_$tag>(A); }")
            push!(src, "")
            names[tag] = "mfb_b$(i0)_$tag"
            smem = max(smem, ft.smem)
        end
        main = forms[haskey(forms, "nl") ? "nl" : first(variants)[1]]
        ft = main[1]
        push!(blocks, BlockText(kind, bg_ID, names["lin"], names["nl"], names["ev"],
                                FormText(ft.body, ft.fields, ft.globals, smem, ft.has_K, block_tpb, ft.qp_in,
                                         haskey(forms, "ev") ? forms["ev"][1].qp_out : ft.qp_out), main[2]))
    end
    return join(src, "\n"), blocks
end
