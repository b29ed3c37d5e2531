"""Build the oracle's C helpers into oracle/_build/liboracle.so (test infrastructure)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")
SRCS = [os.path.join(HERE, "c", f) for f in ("femdict.c", "refpath.c")]


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = [s for s in SRCS if os.path.exists(s)]
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs)):
        return LIB
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-o", LIB] + srcs + ["-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
