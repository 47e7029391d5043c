"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU
baseline.  ``muzero_b200`` never imports this package.

Parity status: PINNED.  The reference (michaelnny/muzero) has no tests of
``mcts.py`` / ``network.py``; the oracle is pinned instead against outputs of
the reference itself, imported from ``/root/reference`` in the build
container (``tests/golden/make_golden.py`` writes the fixtures that
``tests/test_oracle_golden.py`` replays on any box).
"""
