"""The single-stage ablation models of the reference's STR_TTR/ directory (spatial-only STR, temporal-only TTR and the
STR_TTR assembly), on the same kernels as ST / TS."""
from .STR import STR  # noqa: F401
from .TTR import TTR  # noqa: F401
from .STR_TTR import STR_TTR  # noqa: F401
