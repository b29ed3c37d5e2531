"""CUDA C term emitter: the sibling of MetaFEM's ``parse_Term2Expr!``
(reference src/symbolics/08_Tensor.jl:169-233) that the north star adds to src/symbolics.

Input: the kernel spec (frontend/weakform.py) -- per generated block the inner/external words,
hoisted temporaries and the residual / gradient terms with C expressions.
Output: one CUDA C translation unit on top of ``mfb_skeleton.cuh`` with two entry points per block
(``mfb_b<i>_lin`` writing K_linear, ``mfb_b<i>_nl`` writing residue + K_total) and the block
descriptors ``mfb_kernel_compile`` needs. Mirrors gen_K_Linear_GPU / gen_Res_K_NonLinear_GPU
(reference src/solver/05_CodeGenerator.jl:52-154), but fuses all terms of a block into one kernel.
"""


def _slot(sd):
    if len(sd) > 1:
        raise ValueError("spatial derivative order > 1 is outside the supported max_sd_order")
    return 0 if not sd else int(sd[0])


def _form(name, spec, blk, n_a, n_q, linear, tpb):
    nv = len(spec["basic_vars"])
    L1 = spec["max_time_level"] + 1
    boundary = 1 if blk["kind"] == "boundary" else 0
    terms = blk["linear_gradients"] if linear else blk["nonlinear_gradients"]
    residues = [] if linear else blk["residues"]
    inner = [] if linear else blk["innervars"]
    ext = blk["extervars"]
    cpw = [w for w in ext if w["kind"] == "cp"]
    fields = sorted({w["local"] for w in cpw})
    globs = [w["sym"] for w in ext if w["kind"] == "global"]
    # symbols referenced by what this kernel evaluates (temps are included wholesale; unused ones are dead code)
    lines = []
    lines.append(f"struct {name} {{")
    lines.append(f"  static constexpr int NV = {nv}, NA = {n_a}, NQ = {n_q}, L1 = {L1}, BOUNDARY = {boundary}, "
                 f"LINEAR = {int(linear)}, NW = {len(inner)}, NCW = {len(cpw)}, NC = {len(fields)}, NT = {len(terms)}, "
                 f"HAS_RES = {int(bool(residues))}, HAS_K = {int(bool(terms))}, TPB = {tpb};")
    lines.append("  template <class S> __device__ static __forceinline__ void words(const S& s, int q, double* w, double* c) {")
    for k, w in enumerate(inner):
        lines.append(f"    w[{k}] = mfb::interp<NA>(&s.G[q][0][{_slot(w['sd'])}], &s.ue[{w['td']}][0][{w['pos']}], NV);")
    for k, w in enumerate(cpw):
        lines.append(f"    c[{k}] = mfb::interp<NA>(&s.G[q][0][{_slot(w['sd'])}], &s.ce[{fields.index(w['local'])}][0], 1);")
    lines.append("  }")
    lines.append("  __device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, "
                 "const MfbArgs& A, double* R, double* D) {")
    for k, w in enumerate(inner):
        lines.append(f"    const double {w['sym']} = w[{k}];")
    for k, w in enumerate(cpw):
        lines.append(f"    const double {w['sym']} = c[{k}];")
    for w in ext:
        if w["kind"] == "normal":
            lines.append(f"    const double {w['sym']} = nrm[{w['c'] - 1}];")
        elif w["kind"] == "global":
            lines.append(f"    const double {w['sym']} = A.glob[{globs.index(w['sym'])}];")
    known = {w["sym"] for w in inner} | {w["sym"] for w in ext}
    import re
    ident = re.compile(r"[A-Za-z_][A-Za-z_0-9]*")
    for t in blk["temps"]:
        names = set(ident.findall(t["expr"])) - {"pow", "log", "exp", "sqrt", "fabs"}
        if all((n in known) or n[0].isdigit() or _is_number(n) for n in names):
            lines.append(f"    const double {t['sym']} = {t['expr']};")
            known.add(t["sym"])
    for t in residues:
        lines.append(f"    R[{t['dual_pos'] * 4 + _slot(t['dual_sd'])}] += {t['expr']};")
    for k, t in enumerate(terms):
        lines.append(f"    D[{k}] = ({t['expr']}) * A.Kp[{t['deriv_td']}];")
    lines.append("  }")
    lines.append("  __device__ static __forceinline__ void kacc(double* acc, const double* Ga, const double* Gb, const double* Dq) {")
    for k, t in enumerate(terms):
        lines.append(f"    acc[{t['dual_pos'] * nv + t['deriv_pos']}] += Ga[{_slot(t['dual_sd'])}] * Gb[{_slot(t['deriv_sd'])}] * Dq[{k}];")
    lines.append("  }")
    lines.append("};")
    smem = 8 * (n_q * n_a * 4 + n_q * max(len(terms), 1) + n_q * nv * 4 + n_a * 3 + L1 * n_a * nv
                + max(len(fields), 1) * n_a) + 4 * n_a + 16
    return "\n".join(lines), fields, globs, smem, bool(terms), bool(residues)


def _is_number(s):
    try:
        float(s)
        return True
    except ValueError:
        return False


def emit(spec, n_a, n_q, n_qb, tpb=128):
    """Returns (cuda_src, [block descriptor dicts])."""
    src = ['#include "mfb_skeleton.cuh"', ""]
    descs = []
    for i, blk in enumerate(spec["blocks"]):
        nq = n_qb if blk["kind"] == "boundary" else n_q
        d = dict(kind=1 if blk["kind"] == "boundary" else 0, bg_ID=blk["bg_ID"], linear_kernel=None,
                 nonlinear_kernel=None, cp_var_names=[], global_names=[], threads_per_block=tpb, smem_bytes=0,
                 has_nonlinear_K=0)
        fields = globs = None
        if blk["linear_gradients"]:
            body, fields, globs, smem, _, _ = _form(f"F_b{i}_lin", spec, blk, n_a, nq, True, tpb)
            src += [body, f'extern "C" __global__ void __launch_bounds__({tpb}) mfb_b{i}_lin(const MfbArgs A) '
                          f'{{ mfb::assemble<F_b{i}_lin>(A); }}', ""]
            d["linear_kernel"] = f"mfb_b{i}_lin"
            d["smem_bytes"] = max(d["smem_bytes"], smem)
        if blk["residues"] or blk["nonlinear_gradients"]:
            body, fields, globs, smem, hask, _ = _form(f"F_b{i}_nl", spec, blk, n_a, nq, False, tpb)
            src += [body, f'extern "C" __global__ void __launch_bounds__({tpb}) mfb_b{i}_nl(const MfbArgs A) '
                          f'{{ mfb::assemble<F_b{i}_nl>(A); }}', ""]
            d["nonlinear_kernel"] = f"mfb_b{i}_nl"
            d["smem_bytes"] = max(d["smem_bytes"], smem)
            d["has_nonlinear_K"] = int(hask)
        d["cp_var_names"] = fields or []
        d["global_names"] = globs or []
        descs.append(d)
    return "\n".join(src), descs
