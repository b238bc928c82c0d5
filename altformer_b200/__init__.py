"""Import shim: the product package lives in the directory `st-gcn-altformer_b200/` (not a valid
Python identifier), so `import altformer_b200` resolves to it through this package's __path__."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "st-gcn-altformer_b200"))

from . import _lib, ops, functional  # noqa: E402,F401
from .functional import set_precision, get_precision  # noqa: E402,F401
from .model.net import Unit2D, conv_init, import_class  # noqa: E402,F401
from .model.unit_agcn import unit_agcn  # noqa: E402,F401
from .model.tcn_gcn import TCN_GCN_unit  # noqa: E402,F401
from .model.AltFormer.model_ST import ST, Mlp, Attention, Block  # noqa: E402,F401
from .model.AltFormer.model_TS import TS  # noqa: E402,F401
from .model.AltFormer.ST_GCN_AltFormer import ST_GCN_AltFormer  # noqa: E402,F401
from .trainer import DataParallelTrainer  # noqa: E402,F401
from .inference import GraphedInference  # noqa: E402,F401
from . import streams  # noqa: E402,F401
from .STR_TTR import STR, TTR, STR_TTR  # noqa: E402,F401

unit_gcn = unit_agcn  # north_star alias; the reference only defines unit_agcn (model/unit_agcn.py:31)
