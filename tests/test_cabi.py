"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/metafem_b200.h declares,
fails loudly without a GPU, and the emitted kernels compile for sm_100a (NVRTC needs no device)."""
import ctypes
import os
import re

import pytest

import metafem_b200 as m

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "metafem_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mfb_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(built_lib):
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in include/metafem_b200.h but not exported"
    # and the ctypes binding declares exactly the header's functions
    assert sorted(m.lib.exported_symbols()) == names


def test_no_cpu_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    assert built_lib.mfb_create(ctypes.byref(h), 0) < 0
    with pytest.raises(m.lib.MfbError):
        m.lib.Context(0)


def test_null_context_is_rejected(built_lib):
    assert built_lib.mfb_mesh_set(None, 1, 1, 1, 1, None, None, None, None, None, None) == -3   # MFB_ERR_ARG
    assert built_lib.mfb_last_error(None) == b"null context"
    assert built_lib.mfb_launch_count(None) == 0


@pytest.mark.parametrize("name,na,nq,nqb", [("thermal", 10, 14, 7), ("linear_elasticity", 20, 27, 9), ("neo_hookean", 20, 27, 9),
                                            ("thermo_elasticity", 20, 27, 9), ("j2", 20, 27, 9), ("j2_fused", 20, 27, 9)])
def test_emitted_kernels_compile_for_sm100a(built_lib, name, na, nq, nqb, tmp_path):
    from helpers import spec_for
    src, descs = m.emitter.emit(spec_for(name), na, nq, nqb)
    cubin = str(tmp_path / "k.cubin")
    m.lib.kernel_check(src, cubin)
    blob = open(cubin, "rb").read()
    for d in descs:
        for k in (d["linear_kernel"], d["nonlinear_kernel"], d["eval_kernel"]):
            if k:
                assert k.encode() in blob
    assert all(d["smem_bytes"] < 227 * 1024 for d in descs)
    # the two-phase J2 form has an argument kernel, the fused one does not and writes the trial state from the residual kernel
    if name == "j2":
        assert descs[0]["eval_kernel"] and len(descs[0]["qp_out_names"]) == 6 and len(descs[0]["qp_in_names"]) == 6
    if name == "j2_fused":
        assert descs[0]["eval_kernel"] is None and len(descs[0]["qp_in_names"]) == 13 and descs[0]["qp_out_names"][-1] == "j2.count"


def test_tangent_tiling_invariants():
    """The skeleton's static_asserts, checked for a range of element types without compiling."""
    for n_a in (4, 8, 10, 20, 27):
        for nv in (1, 2, 3, 4, 6):
            t = m.emitter._tile(n_a, nv)
            assert t["NTC"] % 2 == 0 and t["CG"] * t["NTC"] >= n_a and t["LPW"] <= 32
            assert t["W"] * t["LPW"] >= n_a * nv * t["CG"]
            assert nv * t["NTC"] <= 60 or t["NTC"] == 2


def test_nvrtc_errors_are_reported(built_lib):
    with pytest.raises(m.lib.MfbError) as e:
        m.lib.kernel_check('#include "mfb_skeleton.cuh"\nthis is not CUDA;')
    assert "error" in str(e.value).lower()
