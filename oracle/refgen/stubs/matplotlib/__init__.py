"""Empty stand-in for ``matplotlib`` (import-only use in sympleints/graphs/triangle_rec.py:7)."""
