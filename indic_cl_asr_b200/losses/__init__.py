from .ctc import CTCLoss
from .rnnt import RNNTLoss, RNNTLossNumba, rnnt_loss

__all__ = ["CTCLoss", "RNNTLoss", "RNNTLossNumba", "rnnt_loss"]
