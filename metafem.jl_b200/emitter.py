"""CUDA C term emitter: the sibling of MetaFEM's ``parse_Term2Expr!``
(reference src/symbolics/08_Tensor.jl:169-233) that the north star adds to src/symbolics.

Input: the kernel spec (frontend/weakform.py) -- per generated block the inner/external words,
hoisted temporaries and the residual / gradient terms with C expressions.
Output: one CUDA C translation unit on top of ``mfb_skeleton.cuh`` with two entry points per block
(``mfb_b<i>_lin`` writing K_linear, ``mfb_b<i>_nl`` writing residue + K_total) and the block
descriptors ``mfb_kernel_compile`` needs. Mirrors gen_K_Linear_GPU / gen_Res_K_NonLinear_GPU
(reference src/solver/05_CodeGenerator.jl:52-154), but fuses all terms of a block into one kernel.
"""


import re


def _slot(sd):
    if len(sd) > 1:
        raise ValueError("spatial derivative order > 1 is outside the supported max_sd_order")
    return 0 if not sd else int(sd[0])


def _tile(n_a, nv):
    """Register tile (MT rows x NTC columns per thread) of the tangent contraction and the threads it needs."""
    ntc = 10 if n_a % 10 == 0 else max(d for d in range(2, 11, 2) if n_a % d == 0)
    mt = {1: 2, 2: 4, 3: 6, 4: 8}.get(nv, 2)
    mrows = n_a * nv * nv
    threads = -(-mrows // mt) * -(-n_a // ntc)
    return mt, ntc, threads


def _form(name, spec, blk, n_a, n_q, linear):
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
    dslots = sorted({_slot(t["dual_sd"]) for t in terms})
    bslots = sorted({_slot(t["deriv_sd"]) for t in terms})
    nsd, ks = len(dslots), len(bslots)
    nd = nv * nsd * nv * ks
    mt, ntc, tile_threads = _tile(n_a, nv)
    tpb = 32 * max(-(-tile_threads // 32) if terms else 1, -(-n_q // 32), 1)
    lines = []
    lines.append(f"struct {name} {{")
    lines.append(f"  static constexpr int NV = {nv}, NA = {n_a}, NQ = {n_q}, L1 = {L1}, BOUNDARY = {boundary}, "
                 f"LINEAR = {int(linear)}, NW = {len(inner)}, NCW = {len(cpw)}, NC = {len(fields)}, "
                 f"HAS_RES = {int(bool(residues))}, HAS_K = {int(bool(terms))}, TPB = {tpb}, "
                 f"NSD = {nsd}, KS = {ks}, ND = {nd}, MT = {mt}, NTC = {ntc};")
    for fn, sl in (("dslot", dslots), ("bslot", bslots)):
        body = "{" + ", ".join(str(v) for v in (sl or [0])) + "}"
        lines.append(f"  __device__ static constexpr int {fn}(int i) {{ constexpr int t[] = {body}; return t[i]; }}")
    lines.append("  template <class S> __device__ static __forceinline__ void words(const S& s, int q, double* w, double* c) {")
    for k, w in enumerate(inner):
        lines.append(f"    w[{k}] = mfb::interp<NA>(&s.G[q][{_slot(w['sd'])}][0], &s.ue[{w['td']}][0][{w['pos']}], NV);")
    for k, w in enumerate(cpw):
        lines.append(f"    c[{k}] = mfb::interp<NA>(&s.G[q][{_slot(w['sd'])}][0], &s.ce[{fields.index(w['local'])}][0], 1);")
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
    ident = re.compile(r"[A-Za-z_][A-Za-z_0-9]*")
    for t in blk["temps"]:
        names = set(ident.findall(t["expr"])) - {"pow", "log", "exp", "sqrt", "fabs"}
        if all((n in known) or _is_number(n) for n in names):
            lines.append(f"    const double {t['sym']} = {t['expr']};")
            known.add(t["sym"])
    for t in residues:
        lines.append(f"    R[{t['dual_pos'] * 4 + _slot(t['dual_sd'])}] += {t['expr']};")
    for t in terms:
        idx = ((t["dual_pos"] * nsd + dslots.index(_slot(t["dual_sd"]))) * nv + t["deriv_pos"]) * ks \
            + bslots.index(_slot(t["deriv_sd"]))
        lines.append(f"    D[{idx}] += ({t['expr']}) * A.Kp[{t['deriv_td']}];")
    lines.append("  }")
    lines.append("};")
    mrows = n_a * nv * nv
    smem = 8 * (n_q * 4 * n_a + 2 * max(ks, 1) * mrows + n_q * max(nd, 1) + n_q * nv * 4 + n_a * 3 + L1 * n_a * nv
                + max(len(fields), 1) * n_a) + 4 * n_a + 32
    return "\n".join(lines), fields, globs, smem, bool(terms), tpb


def _is_number(s):
    try:
        float(s)
        return True
    except ValueError:
        return False


def emit(spec, n_a, n_q, n_qb, tpb=None):
    """Returns (cuda_src, [block descriptor dicts]). ``tpb`` is ignored (kept for API compatibility): the launch
    shape follows from the register tiling of each block."""
    src = ['#include "mfb_skeleton.cuh"', ""]
    descs = []
    for i, blk in enumerate(spec["blocks"]):
        nq = n_qb if blk["kind"] == "boundary" else n_q
        d = dict(kind=1 if blk["kind"] == "boundary" else 0, bg_ID=blk["bg_ID"], linear_kernel=None,
                 nonlinear_kernel=None, cp_var_names=[], global_names=[], threads_per_block=32, smem_bytes=0,
                 has_nonlinear_K=0)
        fields = globs = None
        variants = []
        if blk["linear_gradients"]:
            variants.append(("lin", True))
        if blk["residues"] or blk["nonlinear_gradients"]:
            variants.append(("nl", False))
        forms = {tag: _form(f"F_b{i}_{tag}", spec, blk, n_a, nq, lin) for tag, lin in variants}
        # both kernels of a block are launched with the same shape
        block_tpb = max((f[5] for f in forms.values()), default=32)
        for tag, lin in variants:
            body, fields, globs, smem, hask, _ = forms[tag]
            body = re.sub(r"TPB = \d+", f"TPB = {block_tpb}", body, count=1)
            src += [body, f'extern "C" __global__ void __launch_bounds__({block_tpb}) mfb_b{i}_{tag}(const MfbArgs A) '
                          f'{{ mfb::assemble<F_b{i}_{tag}>(A); }}', ""]
            d["linear_kernel" if lin else "nonlinear_kernel"] = f"mfb_b{i}_{tag}"
            d["smem_bytes"] = max(d["smem_bytes"], smem)
            if not lin:
                d["has_nonlinear_K"] = int(hask)
        d["threads_per_block"] = block_tpb
        d["cp_var_names"] = fields or []
        d["global_names"] = globs or []
        descs.append(d)
    return "\n".join(src), descs
