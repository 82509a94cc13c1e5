from .conv import MHNNConv, MHNNSConv
from .egnn_layer import EGNN
from .mlp import MLP

__all__ = ["MLP", "MHNNConv", "MHNNSConv", "EGNN"]
