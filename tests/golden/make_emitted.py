"""Regenerates tests/golden/emitted/*.cu: the CUDA C translation unit the emitter produces for every BASELINE config.
Run after a DELIBERATE change of metafem.jl_b200/emitter.py or frontend/weakform.py:  python tests/golden/make_emitted.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = {"thermal": (10, 14, 7), "linear_elasticity": (20, 27, 9), "neo_hookean": (20, 27, 9), "thermo_elasticity": (20, 27, 9),
         "j2": (20, 27, 9), "j2_fused": (20, 27, 9)}


def emit_all():
    import metafem_b200 as m
    sys.path.insert(0, os.path.dirname(HERE))
    from helpers import spec_for
    return {name: m.emitter.emit(spec_for(name), *shape)[0] for name, shape in CASES.items()}


if __name__ == "__main__":
    os.makedirs(os.path.join(HERE, "emitted"), exist_ok=True)
    for name, src in emit_all().items():
        open(os.path.join(HERE, "emitted", name + ".cu"), "w").write(src)
        print(name, len(src))
