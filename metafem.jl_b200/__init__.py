"""metafem.jl_b200 -- B200-native hot path of MetaFEM.jl (element evaluation, CSR assembly, Krylov solve).

The directory name carries a dot, so it is imported through ``metafem_b200.py`` at the repository root.
"""
from . import lib, emitter, api  # noqa: F401
from .api import (FEM_Domain, GeneralAlpha, GlobalField, MeshTables, assemble_Global_Variables, assemble_X,  # noqa: F401
                  dessemble_X, compile_Updater_GPU, update_OneStep, iterative_Solve, comm_unique_id, init_distributed,
                  write_VTK)
