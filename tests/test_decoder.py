"""RNNTDecoder (prediction network work-alike) against reference runs (tests/golden/ref_decoder.npz, produced by the
reference's own RNNTDecoder source, oracle/gen_golden.py::gen_decoder).  Plain torch + cuDNN module: runs on CPU here."""
import numpy as np
import pytest
import torch

from conftest import split_cases
from indic_cl_asr_b200.modules.rnnt_decoder import RNNTDecoder, label_collate


@pytest.mark.parametrize("case", ["one_layer", "two_layers_fgb"])
def test_decoder_vs_reference_run(golden, case):
    c = split_cases(golden("ref_decoder.npz"))[case]
    B, U, V, H, L = [int(x) for x in c["cfg"]]
    m = RNNTDecoder(prednet=dict(pred_hidden=H, pred_rnn_layers=L, dropout=0.0), vocab_size=V)
    sd = {k[2:]: torch.tensor(v) for k, v in c.items() if k.startswith("p.")}
    assert sorted(sd) == sorted(m.state_dict())           # same parameter names as the reference module
    m.load_state_dict(sd)
    tg = torch.tensor(c["targets"])
    g, tl, (h, cc) = m(targets=tg, target_length=torch.full((B,), U))
    assert g.shape == (B, H, U + 1)                        # [B, D, U+1] like the reference (:678-681)
    assert np.allclose(g.detach().numpy(), c["g"], atol=1e-6)
    assert np.allclose(h.detach().numpy(), c["h"], atol=1e-6) and np.allclose(cc.detach().numpy(), c["c"], atol=1e-6)
    (g * torch.tensor(c["w"])).sum().backward()
    for k, p in m.named_parameters():
        ref = c["g." + k]
        assert np.abs(p.grad.numpy() - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-6), k
    # SOS step: the first output position does not depend on the targets
    g2, _, _ = m(targets=(tg + 1) % V, target_length=torch.full((B,), U))
    assert torch.equal(g[:, :, 0], g2[:, :, 0])


def test_decoder_init_and_helpers():
    torch.manual_seed(0)
    m = RNNTDecoder(prednet=dict(pred_hidden=8, pred_rnn_layers=1, forget_gate_bias=1.0), vocab_size=5)
    b_ih, b_hh = m.prediction["dec_rnn"].lstm.bias_ih_l0, m.prediction["dec_rnn"].lstm.bias_hh_l0
    assert torch.all(b_ih[8:16] == 1.0) and torch.all(b_hh[8:16] == 0.0)   # common/parts/rnn.py:209-216
    assert m.prediction["embed"].padding_idx == 5 and m.prediction["embed"].weight.shape == (6, 8)
    assert label_collate([[1, 2, 3], [4]]).tolist() == [[1, 2, 3], [4, 0, 0]]
    with pytest.raises(ValueError):
        label_collate(3)
    g, hid = m.predict(None, None, add_sos=False, batch_size=2)            # priming call of greedy decoding
    assert g.shape == (2, 1, 8)
    with pytest.raises(TypeError):
        m(torch.zeros(1, 2, dtype=torch.long), torch.tensor([2]))          # keyword-only, like @typecheck
