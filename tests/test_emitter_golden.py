"""The emitted CUDA C of every BASELINE config is pinned byte for byte (tests/golden/emitted/*.cu, made by
tests/golden/make_emitted.py): a change of the emitter or of the weak-form derivation shows up as a diff that has to be
regenerated on purpose. julia/parse_Term2CUDA.jl documents the same struct layout for the Julia emitter; the structural
lines it must produce (constants, tables, entry points) are checked against the golden text."""
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_emitted  # noqa: E402


@pytest.fixture(scope="module")
def emitted():
    return make_emitted.emit_all()


@pytest.mark.parametrize("name", sorted(make_emitted.CASES))
def test_emitted_translation_unit_is_byte_identical(emitted, name):
    golden = open(os.path.join(ROOT, "tests", "golden", "emitted", name + ".cu")).read()
    assert emitted[name] == golden


def test_julia_emitter_prints_the_same_struct_frame():
    """Every fixed piece of text of the Form struct (member names in order, table names, entry-point line) appears verbatim in
    julia/parse_Term2CUDA.jl, so the two emitters cannot drift apart silently."""
    jl = open(os.path.join(ROOT, "julia", "parse_Term2CUDA.jl")).read()
    golden = open(os.path.join(ROOT, "tests", "golden", "emitted", "neo_hookean.cu")).read()
    consts = re.search(r"static constexpr int (.*?);", golden).group(1)
    names = [c.split("=")[0].strip() for c in consts.split(",")]
    pos = -1
    for nm in names:                                                    # same constants, same order
        nxt = jl.find(f"{nm} = ", pos + 1)
        assert nxt > pos, nm
        pos = nxt
    for table in ("gslot", "gslot_id", "dslot", "bslot", "wslot", "wlev", "wpos", "cslot", "cfield"):
        assert f'c_table("{table}"' in jl and f"constexpr int {table}(int i)" in golden
    assert "__device__ static __forceinline__ void point(const double* w, const double* c, const double* nrm, " in jl
    assert "const double* qv, size_t qidx, const MfbArgs& A, double* R, double* D) {" in jl
    assert 'mfb_b$(i0)_$tag(const MfbArgs A) { mfb::assemble<F_b$(i0)_$tag>(A); }' in jl
    assert re.search(r"mfb_b0_nl\(const MfbArgs A\) \{ mfb::assemble<F_b0_nl>\(A\); \}", golden)
