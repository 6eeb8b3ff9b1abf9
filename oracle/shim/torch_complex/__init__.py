"""Stub of the un-vendored torch_complex dependency: imported by CRN_ELU.py:6-7 but unused on the path."""


class ComplexTensor:  # noqa: D101 - placeholder only
    pass


from . import functional  # noqa: E402,F401
