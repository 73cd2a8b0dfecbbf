"""Import the UNMODIFIED reference Python modules from /root/reference on CPU.

Test infrastructure only -- see oracle/__init__.py.  Works only in the build
container (the GPU box has no /root/reference); used by
``tests/golden/make_golden.py`` to generate the committed golden vectors and by
CPU tests that skip when the reference is absent.  No reference source is copied:
the modules are executed where they lie.  Five modules the reference imports but
that are not installed / not runnable on CPU are replaced in ``sys.modules``
(SURVEY.md Appendix C):

  score_computation_cuda, value_aggregation_cuda, fast_score_computation
      -> the op restatements in oracle/ops.py
  kornia (create_meshgrid, dsnt.spatial_expectation2d), timm.models.layers
      -> minimal stand-ins with kornia-0.6.2 / timm-0.3.2 published behaviour
"""
import importlib
import os
import sys
import types

import torch

from . import ops

REF_ROOT = os.environ.get('CASMTR_REFERENCE', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'cuda_imp'))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []            # lets "import a.b" treat it as a package
    sys.modules[name] = m
    return m


def _create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
    xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
    if normalized_coordinates:
        xs = (xs / (width - 1) - 0.5) * 2
        ys = (ys / (height - 1) - 0.5) * 2
    gy, gx = torch.meshgrid(ys, xs, indexing='ij')
    return torch.stack([gx, gy], dim=-1).unsqueeze(0)            # [1,H,W,2]


def _spatial_expectation2d(x, normalized_coordinates=True):
    B, N, H, W = x.shape
    grid = _create_meshgrid(H, W, normalized_coordinates, x.device).to(x.dtype)
    px = grid[..., 0].reshape(-1)
    py = grid[..., 1].reshape(-1)
    flat = x.reshape(B, N, -1)
    return torch.stack([(flat * px).sum(-1), (flat * py).sum(-1)], dim=-1)


def install_stubs():
    if 'score_computation_cuda' in sys.modules and getattr(sys.modules['score_computation_cuda'], '_oracle_stub', False):
        return
    _mod('score_computation_cuda', _oracle_stub=True,
         score_forward=lambda q, k, i: [ops.score5d(q, k, i)])

    def _va_fwd(score, value, index, output):
        output.copy_(ops.value_agg(score, value, index))
    _mod('value_aggregation_cuda', value_aggregation_forward=_va_fwd)
    _mod('fast_score_computation', score_forward=lambda q, k, i: [ops.score3d(q, k, i)])

    kornia = _mod('kornia')
    _mod('kornia.feature')
    utils = _mod('kornia.utils', create_meshgrid=_create_meshgrid)
    _mod('kornia.utils.grid', create_meshgrid=_create_meshgrid)
    dsnt = _mod('kornia.geometry.subpix.dsnt', spatial_expectation2d=_spatial_expectation2d)
    subpix = _mod('kornia.geometry.subpix', dsnt=dsnt)
    geometry = _mod('kornia.geometry', subpix=subpix)
    kornia.utils, kornia.geometry = utils, geometry
    kornia.feature = sys.modules['kornia.feature']

    class DropPath(torch.nn.Identity):
        def __init__(self, p=0.0):
            super().__init__()

    timm = _mod('timm')
    models = _mod('timm.models')
    layers = _mod('timm.models.layers', DropPath=DropPath,
                  trunc_normal_=torch.nn.init.trunc_normal_, to_2tuple=lambda x: (x, x))
    timm.models, models.layers = models, layers


def load():
    """Returns a namespace with the reference classes/functions the golden generator needs."""
    if not available():
        raise RuntimeError(f'reference tree not found at {REF_ROOT}')
    install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ns = types.SimpleNamespace()
    qta = importlib.import_module('cuda_imp.QuadTreeAttention.QuadtreeAttention.modules.quadtree_attention')
    smart = importlib.import_module('cuda_imp.QuadTreeAttention.QuadtreeAttention.modules.quadtree_attention_smart')
    cf = importlib.import_module('src.model.functions.cascade_functions')
    cm = importlib.import_module('src.model.functions.cascade_matching')
    fm = importlib.import_module('src.model.functions.fine_matching')
    ns.QTAttA, ns.QTAttB, ns.CascadeQTAttB, ns.QTAttGuided = qta.QTAttA, qta.QTAttB, qta.CascadeQTAttB, qta.QTAttGuided
    ns.SmartQTAttB, ns.torch_gather_b2 = smart.QTAttB, smart.torch_gather_b2
    ns.torch_gather = cf.torch_gather
    ns.CascadeMatching = cm.CascadeMatching
    ns.CascadeFineMatching, ns.FineMatching = fm.CascadeFineMatching, fm.FineMatching
    return ns
