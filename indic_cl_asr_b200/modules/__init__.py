from .conv_asr import ConvASRDecoder
from .rnnt import RNNTJoint
from .rnnt_decoder import RNNTDecoder

__all__ = ["ConvASRDecoder", "RNNTJoint", "RNNTDecoder"]
