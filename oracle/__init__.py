"""CPU oracle for the Geo4D inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or the timed CPU baseline.  The product (``geo4d_b200``) never imports
this package and fails loudly when its CUDA extension is missing.

Every function here is a plain fp32 PyTorch/numpy restatement of the
reference's algorithm (jzr99/Geo4D @ 2e57abc) and cites the reference
file:line it follows.  The restatements are pinned against the reference's
own modules (imported from /root/reference in the build container) by
``oracle/gen_golden.py``, which writes the fixtures under ``tests/golden/``.
The reference ships no tests or golden vectors of its own,
so those fixtures -- outputs of the reference itself on seeded inputs -- are
the pin.  Third-party closed-form pieces that are not vendored in the
reference (roma, evo) are restated from their published algorithm and are
"parity unpinned" beyond analytic known-answer tests; see oracle/align.py.
"""
