r"""Guidance wrappers on the generation path (interface of ``azula/guidance``): classifier-free guidance."""

from . import cfg  # noqa: F401
