"""Import-name stand-ins for the reference's three pybind extensions.

Putting this directory on ``sys.path`` (``casmtr_b200.shims.install()``) makes
``import score_computation_cuda``, ``import value_aggregation_cuda`` and
``import fast_score_computation`` resolve to modules backed by libcasmtr_b200.so, so the
reference's unmodified Python (cuda_imp/QuadTreeAttention/QuadtreeAttention/functions/quadtree_attention.py:1-2,
src/model/functions/cascade_functions.py:1) runs on the B200 kernels, forward and backward.
"""
import os
import sys


def install():
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
