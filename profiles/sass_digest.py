"""SASS digest of the hot kernels (runs without a GPU): mnemonic counts from `cuobjdump -sass` for
  * the NVRTC-compiled fused element kernel of the Neo-Hookean config (mfb_b0_nl) and of the thermo-elastic K_linear kernel,
  * k_spmv_mr<3> (production SpMV), k_spmv_bsr<3> (round-1 SpMV), k_fused<0..24>, k_dots, k_light, the Pl_ILU kernels
    in libmetafem_b200.so.
usage: python profiles/sass_digest.py > profiles/sass_digest_r2.txt"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

KEYS = ["DFMA", "DMUL", "DADD", "MUFU", "LDS.128", "LDS.64", "LDS", "STS", "LDG", "LD.E.128", "LD", "STG", "ST", "REDG", "RED", "ATOMG", "ATOM", "LDL", "STL", "SHFL",
        "BAR", "BRA", "CALL", "UBLKCP", "SYNCS"]


def digest(path, pattern):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    out, name = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1) if re.search(pattern, m.group(1)) else None
            if name:
                out[name] = collections.Counter()
            continue
        if name:
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                op = m.group(1)
                out[name]["TOTAL"] += 1
                for k in KEYS:
                    if op == k or op.startswith(k + "."):
                        out[name][k] += 1
                if op.startswith("LDS.128") or ".128" in op and op.startswith("LDS"):
                    pass
    return out


def show(title, d):
    print(f"## {title}")
    for fn, c in d.items():
        short = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()[:110]
        print(f"{short}\n   " + "  ".join(f"{k}={c[k]}" for k in ["TOTAL"] + KEYS if c[k]))
    print()


def main():
    import metafem_b200 as m
    from helpers import spec_for
    for name, pat in (("neo_hookean", r"mfb_b0_nl"), ("thermo_elasticity", r"mfb_b0_lin"), ("linear_elasticity", r"mfb_b0_(lin|nl)")):
        src, _ = m.emitter.emit(spec_for(name), 20, 27, 9)
        with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as f:
            path = f.name
        m.lib.kernel_check(src, path)
        show(f"NVRTC element kernels, {name} (hex20 x 27)", digest(path, pat))
        os.unlink(path)
    so = os.path.join(ROOT, "metafem.jl_b200", "libmetafem_b200.so")
    show("libmetafem_b200.so: SpMV", digest(so, r"k_spmv_mrILi3ELi5ELi16ELb0ELi4ELi0|k_spmv_bsrILi3ELi5ELi0ELi64"))
    show("libmetafem_b200.so: fused vector programs", digest(so, r"k_fusedILi|k_dotsILi|k_lightILi"))
    show("libmetafem_b200.so: Pl_ILU (NV = 3): factorisation, production sweeps (stream kernel on FP32 factors), fallbacks",
         digest(so, r"k_ilu_factorILi3|k_sweep_mrILi3ELi5ELi8ELi[12]ELi4EfLi3|k_ilu_sweep(_packed)?ILi3"))


if __name__ == "__main__":
    main()
