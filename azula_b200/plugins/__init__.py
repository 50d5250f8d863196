r"""Plugins: pre-trained model families behind the Denoiser interface (``azula/plugins``)."""
