"""CPU oracle for the CasMTR coarse-to-fine matching hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``casmtr_b200/`` may import this
package; the only legitimate callers are ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``, where it is
the checker or the CPU baseline, never the thing shipped.

What it is: an independent torch-CPU (fp32) restatement of the algorithm the
reference implements in
  * cuda_imp/QuadTreeAttention/QuadtreeAttention/src/*.cu       (ops.py)
  * cuda_imp/QuadTreeAttention/QuadtreeAttention/modules/quadtree_attention.py (qtatt.py)
  * cuda_imp/score_cuda/src/score_computation_kernel.cu          (ops.py)
  * src/model/functions/cascade_matching.py, post_processing.py,
    cascade_functions.py                                         (cascade.py)
  * src/model/functions/fine_matching.py                         (fine.py)
every function citing the reference file:line it follows.

Pinning: the reference ships no tests, golden vectors or fixtures for this
path.  The oracle is therefore pinned against OUTPUTS OF
THE REFERENCE ITSELF: ``tests/golden/make_golden.py`` imports the reference's
own Python modules from /root/reference (``oracle/refload.py``; compiled
extensions replaced by the op restatements of ``ops.py``, which are in turn
checked against the reference's own pure-PyTorch formulations
``quadtree_attention_smart.torch_gather_b2`` and ``cascade_functions.torch_gather``),
runs them on seeded inputs and commits the input/output vectors under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks the oracle against
those vectors on every CPU run.  torch.topk / torch.max tie order is
unspecified in the reference; synthetic inputs carry a tie guard instead.
"""
from . import ops, qtatt, cascade, fine  # noqa: F401
