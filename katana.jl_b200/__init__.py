"""katana.jl_b200 -- B200-native ECP separation round behind Katana.jl's separator API.

Import as `katana_jl_b200` (the repo-root shim `katana_jl_b200.py` registers this
directory, whose name is not a valid Python identifier, under that module name).
"""
from . import binding, expr  # noqa: F401
from .binding import CutBatch, KtnError, KtnLibrary, WireRows, load_cuda_library  # noqa: F401
from .lp import HighsLP  # noqa: F401
from .model import KatanaNonlinearModel  # noqa: F401
from .nlpeval import EpigraphNLPEvaluator, ExprNLPEvaluator  # noqa: F401
from .separators import AbstractKatanaSeparator, AffExpr, KatanaGPUSeparator, linear_oa_cut  # noqa: F401
from .solver import KatanaSolver, LinearQuadraticModel, Model, NonlinearModel, getKatanaModel  # noqa: F401
