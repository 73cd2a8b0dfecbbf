from .quadtree_attention import ScoreComputation, score_computation_op, value_aggregation, value_aggregation_op  # noqa: F401
