"""Build the reference's three CUDA extensions, UNMODIFIED, from /root/reference into oracle/_ref/ (git-ignored; the
.so files travel to the GPU box with the repo snapshot).  Test infrastructure only -- see oracle/__init__.py.

    python -m oracle.build_ref            # only where /root/reference exists (the build container)

The .cpp files are compiled straight from the reference tree; each .cu is compiled through a 2-line wrapper in
oracle/ref_ext/ that adds one overload (shim.h) and then #includes the reference file where it lies.  No reference
source is copied or edited.  The resulting modules (score_computation_cuda, value_aggregation_cuda,
fast_score_computation) are what tests/test_gpu_vs_reference_ext.py compares libcasmtr_b200.so against on the B200.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('CASMTR_REFERENCE', '/root/reference')
OUT = os.path.join(HERE, '_ref')
QT = os.path.join(REF, 'cuda_imp/QuadTreeAttention/QuadtreeAttention/src')
SC = os.path.join(REF, 'cuda_imp/score_cuda/src')
MODULES = {
    'score_computation_cuda': [os.path.join(QT, 'score_computation.cpp'), os.path.join(HERE, 'ref_ext/qt_score_kernel.cu')],
    'value_aggregation_cuda': [os.path.join(QT, 'value_aggregation.cpp'), os.path.join(HERE, 'ref_ext/qt_value_kernel.cu')],
    'fast_score_computation': [os.path.join(SC, 'score_computation.cpp'), os.path.join(HERE, 'ref_ext/fast_score_kernel.cu')],
}


def available():
    return os.path.isdir(QT) and os.path.isdir(SC)


def built(name):
    return os.path.exists(os.path.join(OUT, name, name + '.so'))


def build(verbose=False):
    if not available():
        raise RuntimeError(f'reference tree not found at {REF}')
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    import concurrent.futures as cf
    from torch.utils import cpp_extension

    def one(item):
        name, srcs = item
        if built(name):
            return name
        d = os.path.join(OUT, name)
        os.makedirs(d, exist_ok=True)
        cpp_extension.load(name=name, sources=srcs, build_directory=d, verbose=verbose, is_python_module=False,
                           extra_include_paths=[QT, SC], extra_cflags=['-O2', '-w'],
                           extra_cuda_cflags=['-O2', '-w', '-gencode', 'arch=compute_100a,code=sm_100a'])
        return name
    with cf.ThreadPoolExecutor(max_workers=3) as ex:        # the three extensions build side by side (~4 min on 8 cores)
        list(ex.map(one, MODULES.items()))
    return OUT


def load(name):
    """Import a built reference extension by file path (GPU box: only the prebuilt .so exists)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    path = os.path.join(OUT, name, name + '.so')
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv))
