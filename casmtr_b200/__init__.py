"""casmtr_b200 -- B200 (sm_100a) implementation of CasMTR's coarse-to-fine matching hot path.

Host side = this package (PyTorch for device memory / streams / torch.distributed only); compute =
hand-written CUDA in libcasmtr_b200.so behind the C ABI of include/casmtr_b200.h.  The public names
mirror the reference's module API.
"""
from .modules.quadtree_attention import QTAttA, QTAttB, QTAttGuided, CascadeQTAttB  # noqa: F401
from .modules.attention_layers import QuadtreeAttention, CascadeQuadtreeAttention  # noqa: F401
from .functions.quadtree_attention import score_computation_op, value_aggregation_op  # noqa: F401
from .cascade_matching import CascadeMatching, PostProcess, ScoreComputation  # noqa: F401
from .fine_matching import CascadeFineMatching, CascadeFinePreprocess, FineMatching  # noqa: F401
from .coarse_matching import CoarseMatching  # noqa: F401

__version__ = '0.1.0'
