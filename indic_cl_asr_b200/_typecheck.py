"""Minimal stand-in for NeMo's ``@typecheck()`` call contract.

NeMo-typed modules only accept keyword arguments (reference NeMo/nemo/core/classes/common.py:1058:
"All arguments must be passed by kwargs only for typed methods"); the reference's training_step calls
joint / ctc_decoder / loss modules that way, and the parity tests here do too.
"""
from __future__ import annotations

import functools


def kwargs_only(fn):
    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        if args:
            raise TypeError("All arguments must be passed by kwargs only for typed methods")
        return fn(self, **kwargs)

    return wrapper
