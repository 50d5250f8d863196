r"""Execution engine behind the Azula-compatible surface: coefficient tables, the fused
graph-captured sampling loop and the native sm_100a ADM backbone."""

from . import ops  # noqa: F401,E402  (registers the engine entry points with the ctypes loader)
