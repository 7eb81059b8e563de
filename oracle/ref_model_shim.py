"""oracle/ref_model_shim.py -- TEST INFRASTRUCTURE (build container only).

Imports the reference's OWN GAN2Shape/model.py and GAN2Shape/losses.py UNMODIFIED from /root/reference so that the
caller-side formulas either side of the renderer path (get_clamped_depth, rescale_depth, get_lighting_directions,
get_shading, PhotometricLoss, SmoothLoss) can pin oracle/callers_oracle.py and generate tests/golden/callers_*.npz.
The StyleGAN2 / LPIPS / network stack those files import is replaced by empty stub modules (none of it is touched by the
functions used here); `GAN2Shape.renderer` is the module oracle/ref_shim.py loads; `.cuda()` is the identity (ref_shim).
The model class is never constructed: its methods are called unbound on a namespace that carries the few attributes
they read (min_depth, max_depth, border_depth = 0.7 max + 0.3 min as model.py:51).
/root/reference does not exist on the GPU box: nothing that runs there may import this module.
"""
import importlib.util
import os
import sys
import types

from . import ref_shim

REF_PKG = "/root/reference/GAN2Shape"
_mods = None


def available():
    return ref_shim.available() and os.path.isfile(os.path.join(REF_PKG, "model.py"))


def _load_file(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns (model module, losses module) of the reference (cached)."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("reference tree not present at " + REF_PKG)
    ren = ref_shim.load()
    pkg = types.ModuleType("GAN2Shape")
    pkg.__path__ = [REF_PKG]
    sys.modules["GAN2Shape"] = pkg
    sys.modules["GAN2Shape.renderer"] = ren
    sg = types.ModuleType("GAN2Shape.stylegan2")
    for n in ("Generator", "Discriminator", "PerceptualLoss"):
        setattr(sg, n, type(n, (), {}))
    sys.modules["GAN2Shape.stylegan2"] = sg
    nets = types.ModuleType("GAN2Shape.networks")
    sys.modules["GAN2Shape.networks"] = nets
    pkg.networks = nets
    pkg.utils = _load_file("GAN2Shape.utils", os.path.join(REF_PKG, "utils.py"))
    losses = _load_file("GAN2Shape.losses", os.path.join(REF_PKG, "losses.py"))
    model = _load_file("GAN2Shape.model", os.path.join(REF_PKG, "model.py"))
    _mods = (model, losses)
    return _mods


def model_self(min_depth=0.9, max_depth=1.1):
    """A stand-in `self` for the unbound GAN2Shape methods used here."""
    G = load()[0].GAN2Shape
    ns = types.SimpleNamespace(min_depth=min_depth, max_depth=max_depth, border_depth=0.7 * max_depth + 0.3 * min_depth)
    ns.rescale_depth = lambda d: G.rescale_depth(ns, d)
    return G, ns
