"""Stub of `colorama` so the unmodified reference imports in this container.
Deliberately has no __version__: numba then disables its optional colour support."""


class _Blank:
    def __getattr__(self, name):
        return ""


Fore = _Blank()
