"""oracle/ref_shim.py -- TEST INFRASTRUCTURE (build container only).

Imports the reference's OWN GAN2Shape/renderer/{renderer.py,utils.py} UNMODIFIED from /root/reference
and runs it on torch-CPU, by (a) injecting a stub module named `neural_renderer` whose `Renderer` is
oracle.nr_port.Renderer and (b) making `torch.Tensor.cuda` the identity (renderer.py:33,34,42
hard-code `.cuda()`).  Recipe from SURVEY.md App. C.  Used to pin oracle/renderer_oracle.py against
the real reference code and to generate tests/golden/*.npz (tests/golden/make_golden.py).
/root/reference does not exist on the GPU box: nothing that runs there may import this module.
"""
import importlib.util
import os
import sys
import types
import warnings

import torch

REF_DIR = "/root/reference/GAN2Shape/renderer"


def available():
    return os.path.isfile(os.path.join(REF_DIR, "renderer.py"))


_mod = None


def load():
    """Returns the reference's `GAN2Shape.renderer` package object (cached)."""
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        raise RuntimeError("reference tree not present at " + REF_DIR)
    from . import nr_port
    stub = types.ModuleType("neural_renderer")
    stub.Renderer = nr_port.Renderer
    sys.modules["neural_renderer"] = stub
    torch.Tensor.cuda = lambda self, *a, **k: self
    spec = importlib.util.spec_from_file_location(
        "_g2s_reference_renderer", os.path.join(REF_DIR, "__init__.py"),
        submodule_search_locations=[REF_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_g2s_reference_renderer"] = mod
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    _mod = mod
    return mod


def make_renderer(image_size=128, min_depth=0.9, max_depth=1.1, cfgs=None):
    """Reference Renderer constructed as GAN2Shape/model.py:68 does (config.yml:25-27 values)."""
    cfgs = cfgs if cfgs is not None else {"rot_center_depth": 1.0, "fov": 10, "tex_cube_size": 2}
    return load().Renderer(cfgs, image_size, min_depth, max_depth)


def load_with(nr_module, name):
    """Loads a SECOND, independent copy of the reference's `GAN2Shape.renderer` package whose `import neural_renderer`
    resolves to `nr_module` (e.g. g2s_b200.nr_compat: the product's drop-in for that import).  `name` = the module name
    the copy is registered under."""
    if not available():
        raise RuntimeError("reference tree not present at " + REF_DIR)
    saved = sys.modules.get("neural_renderer")
    sys.modules["neural_renderer"] = nr_module
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_DIR, "__init__.py"),
                                                      submodule_search_locations=[REF_DIR])
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(mod)
    finally:
        if saved is not None:
            sys.modules["neural_renderer"] = saved
        else:
            sys.modules.pop("neural_renderer", None)
    return mod
