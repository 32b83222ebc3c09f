"""Oracle: CPU restatement of the sympleints hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of
``bench.py`` may import it.  The product package (``sympleints_b200``) never
imports from here and fails loudly when its CUDA library is missing.
"""
