from .conv_asr import ConvASRDecoder
from .rnnt import RNNTJoint

__all__ = ["ConvASRDecoder", "RNNTJoint"]
