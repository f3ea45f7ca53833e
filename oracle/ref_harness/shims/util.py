"""Stub of the reference's unreleased in-house `util` module (imported at inexact_Lanczos.py:9,
util_funcs.py:7, printUtils.py:1).  Only au2unit is ever called, and only for convertUnit != 'au'."""


def au2unit(arr, unit):
    if unit.lower() in ("au", "a.u.", "hartree"):
        return arr
    raise NotImplementedError("unit conversion is not available in the harness")
