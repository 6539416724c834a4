"""Minimal stand-in for colorama (the reference's log.py only reads colour constants)."""


class _Blank:
    def __getattr__(self, name):
        return ""


Fore = _Blank()
Back = _Blank()
Style = _Blank()


def init(*args, **kwargs):
    pass
