"""Stage the reference's own numba transducer loss (CPU + numba-CUDA paths) under oracle/_ref/ so that it can run on the
GPU box, where /root/reference does not exist.

    python -m oracle.stage_ref          # authoring container; __graft_entry__.build() calls it when the reference is present

TEST / BENCHMARK INFRASTRUCTURE.  oracle/_ref/ is git-ignored (no reference source enters this repository's history) but
travels with the gpurun snapshot.  What is copied is exactly the self-contained subtree
NeMo/nemo/collections/asr/parts/numba/rnnt_loss/ (13 files; it imports only torch, numba, numpy and itself), byte for
byte.  bench.py times it as (a) the GPU incumbent — numba-CUDA RNNTLossNumba + torch joint + ATen CTC on the same B200 —
and (b) the reference's own CPU transducer at B = 1; the parity tests never depend on it (they use tests/golden/).
"""
from __future__ import annotations

import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/NeMo/nemo/collections/asr/parts/numba/rnnt_loss"
DST_ROOT = os.path.join(HERE, "_ref")
DST = os.path.join(DST_ROOT, "nemo", "collections", "asr", "parts", "numba", "rnnt_loss")


def stage() -> bool:
    if not os.path.isdir(SRC):
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DST_ROOT, "STAGED_FROM.txt"), "w") as fh:
        fh.write(SRC + "\n")
    return True


def available() -> bool:
    return os.path.isfile(os.path.join(DST, "rnnt_pytorch.py"))


def load_rnnt_loss_numba():
    """-> the reference's RNNTLossNumba class, imported from the staged copy through bare namespace modules (the parent
    packages' heavy __init__ files — hydra, lightning — never run; same mechanism as oracle/ref_import.py)."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged (run `python -m oracle.stage_ref` where /root/reference exists)")
    for name in ("nemo", "nemo.collections", "nemo.collections.asr", "nemo.collections.asr.parts",
                 "nemo.collections.asr.parts.numba"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(DST_ROOT, *name.split("."))]
            sys.modules[name] = m
    from nemo.collections.asr.parts.numba.rnnt_loss.rnnt_pytorch import RNNTLossNumba

    return RNNTLossNumba


if __name__ == "__main__":
    print("staged" if stage() else "reference not present", DST)
