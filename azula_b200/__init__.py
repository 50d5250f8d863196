r"""azula_b200 -- a Blackwell (sm_100a) engine for the generation path of Azula.

Keeps Azula's Sampler / Denoiser / Schedule surface (probabilists/azula v0.11.1,
``azula/sample.py``, ``azula/denoise.py``, ``azula/noise.py``, ``azula/plugins/adm``) and
replaces what runs underneath on a CUDA device with hand-written sm_100a kernels reached
through the C ABI of ``libazb.so`` (``include/azb.h``).
"""

__version__ = "0.1.0"

from . import denoise, noise, sample  # noqa: F401
