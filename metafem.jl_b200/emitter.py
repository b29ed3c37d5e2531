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


def _tile(n_a, nv, max_acc=None):
    """Tangent tiling: one lane owns the rows (a, dp, bp = 0..NV-1) and NTC columns of the element matrix, i.e. NV*NTC
    FP64 accumulators; CG column groups cover the NA columns (NTC even for 128-bit operand loads). The NA*NV*CG tiles
    are dealt to W warps, LPW active lanes each."""
    import os
    if max_acc is None:
        max_acc = int(os.environ.get("MFB_MAX_ACC", "60"))       # experiments: 30 = half-width tiles on twice the warps
    cg = 1
    while True:
        ntc = -(-n_a // cg)
        ntc += ntc & 1
        if nv * ntc <= max_acc or ntc == 2:
            break
        cg += 1
    tiles = n_a * nv * cg
    w = -(-tiles // 32)
    return dict(NTC=ntc, CG=cg, W=w, LPW=-(-tiles // w))


def _table(fn, vals):
    body = "{" + ", ".join(str(v) for v in (vals or [0])) + "}"
    return f"  __device__ static constexpr int {fn}(int i) {{ constexpr int t[] = {body}; return t[i]; }}"


def _align16(n):
    return (n + 15) // 16 * 16


def _form(name, spec, blk, n_a, n_q, linear, evalk=False):
    """One Form struct. ``linear``: K_linear kernel; ``evalk``: argument-evaluation kernel of the quadrature-point
    callbacks of the block (phase A of the two-phase nonlinear update); otherwise the residual + K_total kernel."""
    nv = len(spec["basic_vars"])
    L1 = spec["max_time_level"] + 1
    boundary = 1 if blk["kind"] == "boundary" else 0
    terms = blk["linear_gradients"] if linear else blk["nonlinear_gradients"]
    residues = [] if linear else blk["residues"]
    inner = [] if linear else blk["innervars"]
    ext = blk["extervars"]
    calls = blk.get("qp_calls", [])
    if evalk:
        terms, residues = [], []
    qpw = [] if (linear or evalk) else [w for w in ext if w["kind"] == "qp"]
    qpo = [(a, n) for c in calls for a, n in zip(c["args"], c["arg_names"])] if evalk else []
    fused = [c for c in calls if c.get("builtin")] if not (linear or evalk) else []
    if len(fused) > 1 or (fused and len(fused) != len(calls)):
        raise ValueError("one built-in quadrature-point callback per block, not mixed with user callbacks")
    qp_in_names, qp_out_names = [w["sym"] for w in qpw], [n for _, n in qpo]
    if fused:
        if fused[0]["builtin"] != "j2_return_map":
            raise ValueError(f"unknown built-in callback {fused[0]['builtin']!r}")
        pre = fused[0]["prefix"]
        qp_in_names = [f"{pre}.ep{k}" for k in range(1, 7)] + [f"{pre}.b{k}" for k in range(1, 7)] + [f"{pre}.Y"]
        qp_out_names = list(fused[0]["outs"]) + [f"{pre}.b_eval{k}" for k in range(1, 7)] + [f"{pre}.Y_eval", f"{pre}.count"]
    if (qpw or qpo) and boundary:
        raise ValueError("INTEGRATION_POINT_VAR words are supported in domain blocks only")
    cpw = [w for w in ext if w["kind"] == "cp"]
    fields = sorted({w["local"] for w in cpw})
    globs = [w["sym"] for w in ext if w["kind"] == "global"]
    dslots = sorted({_slot(t["dual_sd"]) for t in terms})
    bslots = sorted({_slot(t["deriv_sd"]) for t in terms})
    nsd, ks = len(dslots), len(bslots)
    nd = nv * nsd * nv * ks
    gslots = sorted(set(dslots) | set(bslots) | {_slot(t["dual_sd"]) for t in residues})      # gradient slots G must hold
    tl = _tile(n_a, nv) if terms else dict(NTC=2, CG=1, W=1, LPW=1)
    import os
    tpb = 32 * max((tl["W"] + int(os.environ.get("MFB_EXTRA_WARPS", "0"))) if terms else 2, 2)
    lines = []
    lines.append(f"struct {name} {{")
    lines.append(f"  static constexpr int NV = {nv}, NA = {n_a}, NQ = {n_q}, L1 = {L1}, BOUNDARY = {boundary}, "
                 f"LINEAR = {int(linear)}, NW = {len(inner)}, NCW = {len(cpw)}, NC = {len(fields)}, "
                 f"HAS_RES = {int(bool(residues))}, HAS_K = {int(bool(terms))}, TPB = {tpb}, "
                 f"NSD = {nsd}, KS = {ks}, ND = {nd}, NTC = {tl['NTC']}, CG = {tl['CG']}, W = {tl['W']}, "
                 f"LPW = {tl['LPW']}, SMEM = @SMEM@, "
                 f"EVAL = {int(evalk)}, NQPI = {len(qp_in_names)}, NQPO = {len(qpo)}, NGS = {len(gslots)};")
    lines.append(_table("gslot", [gslots.index(sl) if sl in gslots else -1 for sl in range(4)]))
    lines.append(_table("gslot_id", gslots))
    lines.append(_table("dslot", dslots))
    lines.append(_table("bslot", bslots))
    lines.append(_table("wslot", [_slot(w["sd"]) for w in inner]))
    lines.append(_table("wlev", [w["td"] for w in inner]))
    lines.append(_table("wpos", [w["pos"] for w in inner]))
    lines.append(_table("cslot", [_slot(w["sd"]) for w in cpw]))
    lines.append(_table("cfield", [fields.index(w["local"]) for w in cpw]))
    if evalk:
        lines.append("  __device__ static __forceinline__ void qp_eval(const double* w, const double* c, const MfbArgs& A, "
                     "double* out) {")
        for k, w in enumerate(inner):
            lines.append(f"    const double {w['sym']} = w[{k}];")
        for k, w in enumerate(cpw):
            lines.append(f"    const double {w['sym']} = c[{k}];")
        for w in ext:
            if w["kind"] == "global":
                lines.append(f"    const double {w['sym']} = A.glob[{globs.index(w['sym'])}];")
        for k, (a, _) in enumerate(qpo):
            lines.append(f"    out[{k}] = {a};")
        lines.append("  }")
    lines.append("  __device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, "
                 "const double* qv, size_t qidx, const MfbArgs& A, double* R, double* D) {")
    for k, w in enumerate([] if evalk else inner):
        lines.append(f"    const double {w['sym']} = w[{k}];")
    for k, w in enumerate([] if evalk else cpw):
        lines.append(f"    const double {w['sym']} = c[{k}];")
    for k, w in enumerate([] if fused else qpw):
        lines.append(f"    const double {w['sym']} = qv[{k}];")
    for w in ([] if evalk else ext):
        if w["kind"] == "normal":
            lines.append(f"    const double {w['sym']} = nrm[{w['c'] - 1}];")
        elif w["kind"] == "global":
            lines.append(f"    const double {w['sym']} = A.glob[{globs.index(w['sym'])}];")
    if fused:
        # the library's radial return inlined (reference: the MaterialState callable of J2Plasticity.jl:97-188)
        c0, pre = fused[0], fused[0]["prefix"]
        lines.append("    const double j2_e[6] = {" + ", ".join(c0["args"]) + "};")
        lines.append("    const double j2_ep0[6] = {qv[0], qv[1], qv[2], qv[3], qv[4], qv[5]};")
        lines.append("    const double j2_b0[6] = {qv[6], qv[7], qv[8], qv[9], qv[10], qv[11]};")
        lines.append("    double j2_ep[6], j2_b[6], j2_Y;")
        lines.append(f"    if (mfb::j2_return_map(j2_e, j2_ep0, j2_b0, qv[12], {pre}_lam, {pre}_mu, {pre}_Eb, {pre}_Ep, {pre}_fres, "
                     "j2_ep, j2_b, j2_Y))")
        lines.append("      atomicAdd(reinterpret_cast<unsigned long long*>(A.qpo[13]), 1ull);")
        lines.append("    for (int k = 0; k < 6; ++k) { A.qpo[k][qidx] = j2_ep[k]; A.qpo[6 + k][qidx] = j2_b[k]; }")
        lines.append("    A.qpo[12][qidx] = j2_Y;")
        for k, o in enumerate(c0["outs"]):
            lines.append(f"    const double {o} = j2_ep[{k}];")
    known = {w["sym"] for w in inner} | {w["sym"] for w in ext}
    ident = re.compile(r"[A-Za-z_][A-Za-z_0-9]*")
    for t in ([] if evalk else blk["temps"]):
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
    # shared-memory footprint of mfb::Smem<Form> (the skeleton static_asserts that this bound holds)
    mrows = n_a * nv * nv
    nap = n_a + (n_a & 1)
    dpb = nsd * nv * ks
    nds = nv * (dpb + (dpb & 1)) if nd else 1
    gd = _align16(8 * (n_q * max(len(gslots), 1) * nap + n_q * nds))
    ke = 8 * (mrows * n_a if terms else 1)
    nvl = (0 if linear else L1 * nv) + len(fields)
    geo = _align16(max(8 * n_q * (9 + 1 + 3), 4 * n_a * n_a if terms else 4))
    smem = _align16(_align16(max(gd, ke)) + geo + 8 * (n_q * 4 * max(nvl, 1) + n_a * 3 + L1 * n_a * nv
                                                      + max(len(fields), 1) * n_a) + 4 * n_a)
    body = "\n".join(lines).replace("@SMEM@", str(smem))
    return body, fields, globs, smem, bool(terms), tpb, qp_in_names, qp_out_names


def _is_number(s):
    try:
        float(s)
        return True
    except ValueError:
        return False


def _min_blocks(tpb, smem, regs=170):
    """Resident blocks per SM to ask of __launch_bounds__: what 227 KB of shared memory allow (1 KB reserved per block),
    capped so that every thread keeps at least ``regs`` registers of the 64 K file."""
    import os
    if os.environ.get("MFB_MINB"):
        return int(os.environ["MFB_MINB"])
    regs = int(os.environ.get("MFB_REGS", regs))
    return max(1, min(232448 // (smem + 1024), 65536 // (tpb * regs), 32))


def emit(spec, n_a, n_q, n_qb, tpb=None):
    """Returns (cuda_src, [block descriptor dicts]). ``tpb`` is ignored (kept for API compatibility): the launch
    shape follows from the register tiling of each block."""
    src = ['#include "mfb_skeleton.cuh"', ""]
    descs = []
    for i, blk in enumerate(spec["blocks"]):
        nq = n_qb if blk["kind"] == "boundary" else n_q
        d = dict(kind=1 if blk["kind"] == "boundary" else 0, bg_ID=blk["bg_ID"], linear_kernel=None,
                 nonlinear_kernel=None, cp_var_names=[], global_names=[], threads_per_block=32, smem_bytes=0,
                 has_nonlinear_K=0, eval_kernel=None, qp_in_names=[], qp_out_names=[])
        fields = globs = None
        variants = []
        if blk["linear_gradients"]:
            variants.append(("lin", True))
        if blk["residues"] or blk["nonlinear_gradients"]:
            variants.append(("nl", False))
        if any(not c.get("builtin") for c in blk.get("qp_calls", [])):
            variants.append(("ev", False))
        forms = {tag: _form(f"F_b{i}_{tag}", spec, blk, n_a, nq, lin, evalk=(tag == "ev")) for tag, lin in variants}
        # both kernels of a block are launched with the same shape
        block_tpb = max((f[5] for f in forms.values()), default=32)
        for tag, lin in variants:
            body, fields, globs, smem, hask, _, qpin, qpout = forms[tag]
            body = re.sub(r"TPB = \d+", f"TPB = {block_tpb}", body, count=1)
            src += [body, f'extern "C" __global__ void __launch_bounds__({block_tpb}, {_min_blocks(block_tpb, smem)}) mfb_b{i}_{tag}(const MfbArgs A) '
                          f'{{ mfb::assemble<F_b{i}_{tag}>(A); }}', ""]
            d[{"lin": "linear_kernel", "nl": "nonlinear_kernel", "ev": "eval_kernel"}[tag]] = f"mfb_b{i}_{tag}"
            d["smem_bytes"] = max(d["smem_bytes"], smem)
            if tag == "nl":
                d["has_nonlinear_K"] = int(hask)
                d["qp_in_names"] = qpin
                if qpout:                       # fused built-in callback: the residual kernel writes the trial state
                    d["qp_out_names"] = qpout
            if tag == "ev":
                d["qp_out_names"] = qpout
        d["threads_per_block"] = block_tpb
        d["cp_var_names"] = fields or []
        d["global_names"] = globs or []
        descs.append(d)
    return "\n".join(src), descs
