"""Cosmetic stand-in for `colorama` (absent from this image) so the reference's
driver modules can be imported by oracle/make_golden.py.  Test infrastructure."""


class _Codes:
    def __getattr__(self, name):
        return ""


Fore = _Codes()
Style = _Codes()
Back = _Codes()


def init(*args, **kwargs):
    return None
