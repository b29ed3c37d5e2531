"""Build libmetafem_b200.so in-tree with nvcc for sm_100a (no JIT cache, no CPU fallback)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmetafem_b200.so")
SOURCES = ["mfb_api.cu", "mfb_pattern.cu", "mfb_krylov.cu", "mfb_dist.cu", "mfb_qp.cu", "mfb_vtk.cpp", "mfb_meshbuild.cu", "mfb_totalmesh.cu", "mfb_ilu.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "--extended-lambda"]


def _embed_skeleton():
    src = open(os.path.join(CSRC, "mfb_skeleton.cuh")).read()
    out = os.path.join(CSRC, "mfb_skeleton_embed.h")
    body = "// generated from mfb_skeleton.cuh by build.py -- do not edit\n"
    body += 'static const char mfb_skeleton_src[] = R"MFBSKEL(' + src + ')MFBSKEL";\n'
    if not os.path.exists(out) or open(out).read() != body:
        open(out, "w").write(body)
    return out


def build(force=False, verbose=False):
    emb = _embed_skeleton()
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    obj_of = lambda s: os.path.splitext(s)[0] + ".o"
    deps = srcs + [emb, os.path.join(CSRC, "mfb_internal.h"), os.path.join(CSRC, "mfb_skeleton.cuh"),
                   os.path.join(HERE, "..", "include", "metafem_b200.h")]
    objs = []
    for s in srcs:
        o = obj_of(s)
        objs.append(o)
        hdrs = [d for d in deps if not d.endswith((".cu", ".cpp"))]
        if (not force and os.path.exists(o)
                and all(os.path.getmtime(o) >= os.path.getmtime(d) for d in [s] + hdrs)):
            continue
        cmd = ["nvcc"] + NVCC_FLAGS + (["-x", "cu"] if s.endswith(".cpp") else []) + (["-Xptxas", "-v"] if verbose else []) \
            + ["-c", s, "-o", o]
        subprocess.check_call(cmd)
    if (force or not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(o) for o in objs)):
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-lnvrtc", "-ldl", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
