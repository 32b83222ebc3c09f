"""Minimal stand-in for ``pathos`` (absent in this image); see pools.py."""
