"""Minimal stand-in for ``colorama`` (absent in this image).

The reference imports it for two ANSI constants only (sympleints/helpers.py:10,35).
"""


class _Codes:
    def __getattr__(self, name):
        return ""


Fore = _Codes()
Back = _Codes()
Style = _Codes()
