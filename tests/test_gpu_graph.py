"""GraphedStep (indic_cl_asr_b200/graph.py): a captured training step (fused joint + CTC branch on its side stream + EWC
penalty sweep + backward) must reproduce the eager step, replay after replay and with new data copied into the static
inputs."""
import pytest
import torch

from helpers import rel_err, synth_batch
from indic_cl_asr_b200 import CTCLoss, ConvASRDecoder, HybridRNNTCTCLoss, RNNTJoint, RNNTLoss, cl
from indic_cl_asr_b200.graph import GraphedStep

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_graphed_step_matches_eager_and_follows_new_inputs():
    torch.manual_seed(0)
    V, H, De, Dp = 40, 64, 32, 32
    joint = RNNTJoint(jointnet=dict(encoder_hidden=De, pred_hidden=Dp, joint_hidden=H, activation="tanh", dropout=0.0),
                      num_classes=V, fuse_loss_wer=True, fused_batch_size=4).to(DEV)
    joint.set_loss(RNNTLoss(num_classes=V))
    joint.set_wer(object())
    head = ConvASRDecoder(feat_in=De, num_classes=V).to(DEV)
    model = torch.nn.ModuleDict({"joint": joint, "ctc_decoder": head})
    hybrid = HybridRNNTCTCLoss(joint, head, CTCLoss(num_classes=V, zero_infinity=True), ctc_loss_weight=0.3)
    fp = cl.flat_params(model)
    theta, star = cl.get_params(model), cl.get_params_clone(model)
    star.flat.add_(0.01 * torch.randn_like(star.flat))
    fish = cl.get_zero_params(model, DEV)
    fish.flat.uniform_(0.0, 1.0)
    cfg = {"cl_config": {"e_lambda": 10.0}}

    def step(enc, dec, tr, el, tl):
        fp.bind_grads(zero=False)
        _, avg = cl.get_penalty_grads_async(cfg, fish, theta, star, out=fp.grad)
        enc.grad = None
        dec.grad = None
        loss, _ = hybrid(enc, el, dec, tr, tl)
        loss.backward()
        return loss.detach().reshape(1), avg.reshape(1), enc.grad, fp.grad

    def inputs(seed):
        e, d, tr, el, tl = synth_batch(5, 21, 8, V, De, Dp, seed=seed, device=DEV)
        return [e.requires_grad_(True), d.requires_grad_(True), tr, el, tl]

    a, b = inputs(1), inputs(2)
    ref = {}
    for name, ins in (("a", a), ("b", b)):
        loss, avg, g_enc, g_flat = step(*ins)
        torch.cuda.synchronize()
        ref[name] = (loss.clone(), avg.clone(), g_enc.clone(), g_flat.clone())
    static = inputs(1)
    gs = GraphedStep(step, static, warmup=2)
    for name, ins in (("a", a), ("b", b), ("a", a)):   # replay, new data, back again
        loss, avg, g_enc, g_flat = gs(*[x.detach() for x in ins])
        torch.cuda.synchronize()
        r = ref[name]
        assert abs(loss.item() - r[0].item()) <= 1e-6 * abs(r[0].item()), name
        assert abs(avg.item() - r[1].item()) <= 1e-6 * abs(r[1].item()), name
        assert rel_err(g_enc.cpu().numpy(), r[2].cpu().numpy()) <= 2e-5, name      # fp32 atomics reorder low bits
        assert rel_err(g_flat.cpu().numpy(), r[3].cpu().numpy()) <= 2e-5, name
