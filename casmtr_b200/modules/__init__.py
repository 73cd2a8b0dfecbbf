from .quadtree_attention import QTAttA, QTAttB, QTAttGuided, CascadeQTAttB  # noqa: F401
from .attention_layers import QuadtreeAttention, CascadeQuadtreeAttention  # noqa: F401
