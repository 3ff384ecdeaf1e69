"""CPU-side checks: the C-ABI library loads and exports every symbol include/clasr_b200.h declares;
host-side layout logic; error behaviour that needs no GPU."""
import ctypes
import os

import numpy as np
import pytest
import torch

from indic_cl_asr_b200 import _lib
from indic_cl_asr_b200.cl.flat import SWEEP_CHUNK, Layout


def test_library_exports_every_declared_symbol():
    syms = _lib.declared_symbols()
    assert len(syms) >= 22, syms
    l = _lib.lib()
    for s in syms:
        assert hasattr(l, s), f"{s} declared in include/clasr_b200.h but not exported"
        assert s in _lib._PROTOS, f"{s} has no ctypes prototype"
    assert l.clasr_version() >= 100


def test_workspace_size_queries_need_no_gpu():
    l = _lib.lib()
    B, T, U1 = 32, 250, 101
    cells = B * (T + U1 - 1) * U1
    assert l.clasr_rnnt_workspace_bytes(B, T, U1) >= cells * 20 + 8 * B
    assert l.clasr_rnnt_workspace_bytes(0, T, U1) == 0
    assert l.clasr_ctc_workspace_bytes(B, T, 100) >= 2 * B * T * 201 * 4


def test_invalid_arguments_return_status_not_crash():
    l = _lib.lib()
    # null pointers / bad sizes are rejected before any CUDA call (status 1 = invalid value)
    assert l.clasr_cl_fisher_accum(0, 0, 16, 0, 0) == 1
    assert "fisher" in _lib.last_error()
    assert l.clasr_rnnt_loss_fwd(0, 0, 0, 0, 1, 1, 1, 1, 0, 0.0, 0, 0, 0, 0) == 1
    with pytest.raises(RuntimeError):
        _lib.check(1, "x")


def test_layout_offsets_and_sweep_items():
    lay = Layout([("a.weight", torch.Size([5, 7])), ("a.bias", torch.Size([5])), ("b", torch.Size([3 * SWEEP_CHUNK + 2]))])
    assert lay.offsets == [0, 36, 44]  # each tensor padded to 4 floats
    assert lay.total == 44 + 3 * SWEEP_CHUNK + 4
    it = lay.sweep_items()
    assert it.dtype.itemsize == 16
    # items never straddle tensors, cover every padded float exactly once
    cover = np.zeros(lay.total, dtype=np.int32)
    for s, n, seg in it:
        assert s % 4 == 0 and 0 < n <= SWEEP_CHUNK
        assert lay.offsets[seg] <= s and s + n <= lay.offsets[seg] + lay.numels[seg]
        cover[s:s + n] += 1
    live = np.zeros(lay.total, dtype=bool)
    for off, n in zip(lay.offsets, lay.numels):
        live[off:off + n] = True
    assert (cover[live] == 1).all() and (cover[~live] == 0).all()


def test_cpu_tensors_are_rejected_loudly():
    from indic_cl_asr_b200 import CTCLoss, RNNTLossNumba

    with pytest.raises(RuntimeError, match="no CPU path"):
        RNNTLossNumba(blank=0)(torch.randn(1, 2, 3, 5), torch.zeros(1, 2, dtype=torch.long),
                               torch.tensor([2]), torch.tensor([2]))
    with pytest.raises(RuntimeError, match="no CPU path"):
        CTCLoss(num_classes=4)(log_probs=torch.randn(1, 5, 5), targets=torch.zeros(1, 2, dtype=torch.long),
                               input_lengths=torch.tensor([5]), target_lengths=torch.tensor([2]))


def test_kwargs_only_contract():
    from indic_cl_asr_b200 import CTCLoss, RNNTLoss

    with pytest.raises(TypeError, match="kwargs only"):
        RNNTLoss(num_classes=4)(torch.randn(1, 2, 3, 5), None, None, None)
    with pytest.raises(TypeError, match="kwargs only"):
        CTCLoss(num_classes=4)(torch.randn(1, 5, 5), None, None, None)


def test_constructor_errors_match_reference():
    from indic_cl_asr_b200 import CTCLoss, RNNTJoint, RNNTLoss

    with pytest.raises(ValueError):
        RNNTLoss(num_classes=4, reduction="bogus")
    with pytest.raises(ValueError):
        CTCLoss(num_classes=4, reduction="bogus")
    with pytest.raises(ValueError):
        RNNTLoss(num_classes=4, loss_name="not_a_loss")
    jn = dict(encoder_hidden=8, pred_hidden=8, joint_hidden=64, activation="tanh")
    with pytest.raises(ValueError, match="fused_batch_size"):
        RNNTJoint(jointnet=jn, num_classes=4, fuse_loss_wer=True)
    with pytest.raises(ValueError, match="activation"):
        RNNTJoint(jointnet=dict(jn, activation="gelu"), num_classes=4)
    j = RNNTJoint(jointnet=jn, num_classes=4)
    with pytest.raises(ValueError):
        j.set_loss(object())  # fuse_loss_wer not set
    # state_dict names are the reference's
    assert sorted(j.state_dict()) == ["enc.bias", "enc.weight", "joint_net.1.bias", "joint_net.1.weight",
                                      "pred.bias", "pred.weight"]
    jm = RNNTJoint(jointnet=dict(jn, dropout=0.2), num_classes=8, multilingual=True, language_keys=["hi", "bn"])
    assert "joint_net.2.hi.weight" in jm.state_dict() and jm.joint_net[2]["bn"].out_features == 5


def test_joint_stash_size_and_limit(monkeypatch):
    """The stash a differentiated forward call keeps: 4*pad32(Vp) bytes of logits + 2 (bf16) or 4 (bf16 hi/lo) bytes
    per hidden feature, per padded lattice row; the CLASR_JOINT_STASH switch / limit is host logic."""
    from indic_cl_asr_b200 import fused

    l = _lib.lib()
    B, T, U1, H, Vp = 32, 250, 101, 640, 1025
    rows = B * ((T * U1 + 127) // 128) * 128
    x3 = l.clasr_joint_stash_bytes(B, T, U1, H, Vp, _lib.PREC["bf16x3"])
    x1 = l.clasr_joint_stash_bytes(B, T, U1, H, Vp, _lib.PREC["bf16"])
    assert x3 >= rows * (1056 * 4 + 2 * H * 2) and x3 < rows * (1056 * 4 + 2 * H * 2) + 4096
    assert x3 - x1 >= rows * H * 2 and x3 - x1 < rows * H * 2 + 4096
    assert l.clasr_joint_stash_bytes(0, T, U1, H, Vp, _lib.PREC["bf16x3"]) == 0
    monkeypatch.delenv("CLASR_JOINT_STASH", raising=False)
    assert fused._stash_limit_bytes() == 0            # default: the logits never reach HBM (backward recomputes)
    assert fused._stash_limit_bytes(48.0) == 48 << 30  # RNNTJoint(backward_mode="stash")
    monkeypatch.setenv("CLASR_JOINT_STASH", "0")
    assert fused._stash_limit_bytes() == 0
    monkeypatch.setenv("CLASR_JOINT_STASH", "1.5")
    assert fused._stash_limit_bytes() == 3 << 29
    # no gradient wanted -> nothing is kept, whatever the limit
    assert fused._stash(torch.empty(1), B, T, U1, H, Vp, _lib.PREC["bf16x3"], False) == (None, 0)


def test_nemo_layout_view_detection():
    """linear.py::_is_transposed_view (host logic): only `base.transpose(1, 2)` of a contiguous fp32 [B, K, T] tensor takes
    the tiled-transpose path; everything else keeps torch's own layout handling."""
    import torch
    from indic_cl_asr_b200.linear import _is_transposed_view

    base = torch.randn(3, 8, 5)
    assert _is_transposed_view(base.transpose(1, 2))
    assert _is_transposed_view(base[1:3].transpose(1, 2))            # a batch slice is still one contiguous block
    assert not _is_transposed_view(base)                               # already [B, T, K]
    assert not _is_transposed_view(base.transpose(1, 2).contiguous())
    assert not _is_transposed_view(base[:, :, 1:4].transpose(1, 2))    # narrowed time axis: rows are not K*T apart
    assert not _is_transposed_view(base.double().transpose(1, 2))      # fp32 only
    assert not _is_transposed_view(torch.randn(8, 5).t())              # rank 2
    assert not _is_transposed_view(torch.randn(3, 1, 5).transpose(1, 2))
