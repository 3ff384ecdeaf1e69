"""indic_cl_asr_b200 — B200-native training hot path of the hybrid RNNT-CTC model with EWC/MAS regularisers.

Public surface mirrors the reference's (FrozenWolf-Cyber/Indic-CL-ASR, patched NeMo 1.23):

    from indic_cl_asr_b200 import RNNTJoint, RNNTLoss, RNNTLossNumba, CTCLoss, ConvASRDecoder
    from indic_cl_asr_b200.cl import get_penalty_grads, penalty, get_params, set_grads, ...

All arithmetic on the path runs in hand-written sm_100a kernels (libclasr_sm100.so, C ABI in
include/clasr_b200.h).  There is no CPU path and no fallback: a missing library or a CPU tensor raises.
"""
from . import _lib
from .losses import CTCLoss, RNNTLoss, RNNTLossNumba, rnnt_loss
from .modules import ConvASRDecoder, RNNTDecoder, RNNTJoint
from .hybrid import EncDecHybridRNNTCTCStep, HybridRNNTCTCLoss

__all__ = ["CTCLoss", "RNNTLoss", "RNNTLossNumba", "rnnt_loss", "ConvASRDecoder", "RNNTJoint", "RNNTDecoder", "HybridRNNTCTCLoss",
           "EncDecHybridRNNTCTCStep",
           "_lib"]
__version__ = "0.1.0"
