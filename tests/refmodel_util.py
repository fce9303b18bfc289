"""Test-side helpers for driving the UNMODIFIED reference model (tests/test_model_dropin.py, bench.py's model leg).

The reference tree is found at /root/reference (build container) or at oracle/_ref/reference (the git-ignored copy that
`make -C oracle reftree` stages so that it travels to the GPU box).  The reference's own pybind extensions, built by its
setup.py (`make -C oracle refext`), live in oracle/_ref/ref_ext and are only bound when a test asks for them.
"""
import contextlib
import glob
import importlib
import importlib.machinery
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = ("/root/reference", os.path.join(ROOT, "oracle", "_ref", "reference"))
REF_EXT_DIR = os.path.join(ROOT, "oracle", "_ref", "ref_ext")
EXT_SYMBOLS = {"_correlation_cuda": ("_correlation_forward_cuda", "_correlation_backward_cuda"),
               "_furthest_point_sampling_cuda": ("_furthest_point_sampling_cuda",),
               "_k_nearest_neighbor_cuda": ("_k_nearest_neighbor_cuda",)}
HOT_MODULES = ("models.utils", "models.RPEFlow_core", "models.pwc2d_core", "models.pwc3d_core", "models.pointconv",
               "models.losses3d", "models.RPEFlow", "models.csrc", "models.csrc.wrapper")


def reference_root():
    for c in _CANDIDATES:
        if os.path.isdir(os.path.join(c, "models")):
            return c
    return None


def reference_extensions_available():
    return all(glob.glob(os.path.join(REF_EXT_DIR, name + "*.so")) for name in EXT_SYMBOLS)


def load_reference_extensions():
    """{symbol: callable} from the reference's own compiled extensions (needs torch imported; CUDA to call them)."""
    import torch  # noqa: F401  (the extensions link against libtorch)
    out = {}
    for name, symbols in EXT_SYMBOLS.items():
        path = glob.glob(os.path.join(REF_EXT_DIR, name + "*.so"))[0]
        loader = importlib.machinery.ExtensionFileLoader(name, path)
        spec = importlib.util.spec_from_loader(name, loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        for s in symbols:
            out[s] = getattr(mod, s)
    return out


@contextlib.contextmanager
def reference_extensions_bound():
    """Inside the block, models.csrc.wrapper uses the reference's own CUDA kernels (as if its setup.py had been run)."""
    wrapper = importlib.import_module("models.csrc.wrapper")
    syms = load_reference_extensions()
    old = {s: getattr(wrapper, s) for s in syms}
    for s, fn in syms.items():
        setattr(wrapper, s, fn)
    try:
        yield
    finally:
        for s, fn in old.items():
            setattr(wrapper, s, fn)


@contextlib.contextmanager
def record_calls(names, log):
    """Wrap the named module-level functions in every reference module that imported them; log.append((name, args, out))."""
    saved = []
    for modname in HOT_MODULES:
        mod = sys.modules.get(modname)
        if mod is None:
            continue
        for name in names:
            fn = getattr(mod, name, None)
            if fn is None or getattr(fn, "_recording", False):
                continue

            def rec(*args, _fn=fn, _name=name, **kwargs):
                out = _fn(*args, **kwargs)
                log.append((_name, args, kwargs, out))
                return out
            rec._recording = True
            saved.append((mod, name, fn))
            setattr(mod, name, rec)
    try:
        yield log
    finally:
        for mod, name, fn in saved:
            setattr(mod, name, fn)
