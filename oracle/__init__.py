"""CPU oracle for the generation hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code. Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and there
only as the checker / CPU baseline -- never as the thing shipped or measured as the engine.

What it is: a plain-torch fp32 restatement (functional, state_dict driven, no nn.Module
classes) of the reference algorithm on the path ``Sampler.__call__`` -> ``step`` ->
``Denoiser.forward`` -> backbone, each function citing the reference ``file:line`` it
follows (paths relative to the reference checkout, probabilists/azula @ bec12b8).

Pinning: the reference holds NO golden vectors for this path, so the
oracle is pinned against outputs of the reference itself, imported read-only in the build
container by ``oracle/gen_golden.py``; the resulting small fixtures live in
``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks the oracle against them.
"""
