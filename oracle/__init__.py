"""ORACLE -- test infrastructure, not product code.

CPU restatement (numpy + a little C) of MetaFEM.jl's hot path and of the front-end tables that
feed it, written from the reference's algorithm with file:line citations in every function.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product (``metafem.jl_b200``) never does.

Pinning status: the reference is pure Julia on CUDA.jl, Julia is not installed here and the
reference ships no tests, so the oracle cannot be checked against the reference executable.
It IS pinned at solver tolerance against the committed result files of the reference's own
examples (tests/test_oracle_golden.py): examples/thermal_conduction/3D_MetaFEM_Result.vtk and
examples/linear_elasticity/stress_concentration/3D_MetaFEM.vtk (digests of both are committed
as fixtures under tests/golden/ by tests/golden/make_golden.py), and against the reference's
known answers: the analytical load-displacement table of examples/hypo_elastic_plasticity/
J2Plasticity.jl:223-228 (tests/test_oracle_j2.py: callback semantics, history variables, second time
derivatives) and the closed form of static_Neo_Hookean.jl:124. Element-level values are
"parity unpinned" beyond that.
"""
