from .quadtree_attention import QTAttA, QTAttB, QTAttGuided, CascadeQTAttB  # noqa: F401
