"""Skeleton graphs (host-side constants, init-time only).  `graph.SHRE` / `graph.LMDHG` keep the names the
reference resolves by string (`import_class('graph.SHRE')`, model/net.py:68-74, graph/__init__.py:1-2)."""
from .skeletons import SHRE, LMDHG, Skeleton  # noqa: F401
