"""minibase-columnar-database_b200: B200-native columnar scan hot path of MiniBase-Columnar-Database.

Layout
    csrc/        hand-written sm_100a CUDA kernels + the C ABI (include/mbcol.h) -> libmbcol.so
    _native.py   ctypes binding of the C ABI (stand-in for the Java FFM/JNI stub)
    engine.py    Context / Table / Result objects over the ABI
    global_.py, heap.py, iterator.py, columnar.py, bitmap.py, index.py, input.py
                 host-side mirror of the reference's Java operator surface for this path
                 (same class and method names as minijava/src/<package>/)

Import as ``import mbcol`` (alias module at the repo root).
"""
from . import _native
from ._native import MbcError
from .engine import Context, Table, Result, Shard, Term, bitmap_join

__all__ = ["Context", "Table", "Result", "Shard", "Term", "bitmap_join", "MbcError", "_native"]
