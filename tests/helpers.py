"""Shared helpers for the test-suite (oracle access, emu harness, comparisons)."""
import ctypes
import os
import subprocess

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CFGS = {"rot_center_depth": 1.0, "fov": 10, "tex_cube_size": 2}
MIN_DEPTH, MAX_DEPTH = 0.9, 1.1


def oracle_renderer(S):
    from oracle import renderer_oracle as ro
    return ro.OracleRenderer(dict(CFGS), S, MIN_DEPTH, MAX_DEPTH)


def golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def rel_err(a, b):
    """max |a-b| / max(|b|): the 'within 1e-5 relative' measure used for depth, images and gradients."""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def el_err(a, b, floor_frac=1e-2):
    """ELEMENT-WISE relative error with an absolute floor: max_i |a_i - b_i| / (|b_i| + floor_frac * max|b|).
    `rel_err` normalises every element by max|b|; this one holds small elements to (almost) their own scale.  The floor
    keeps elements that are zero up to accumulation noise (an fp32 ulp of the largest terms: ~1e-7 max|b|) (float atomics, fp32 sums) from dividing by nothing."""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    floor = floor_frac * b.abs().max().clamp_min(1e-30)
    return ((a - b).abs() / (b.abs() + floor)).max().item()


_STATS = os.path.join(ROOT, "gpurun_out", "parity_stats.jsonl")


def log_stats(test, **kw):
    """append the measured errors of a parity test to gpurun_out/parity_stats.jsonl (read back after a gpurun call)"""
    import json
    try:
        os.makedirs(os.path.dirname(_STATS), exist_ok=True)
        with open(_STATS, "a") as f:
            f.write(json.dumps(dict(test=test, **kw)) + "\n")
    except OSError:
        pass


# ---- emu harness (product arithmetic headers compiled for the host) -------------------------------------------
class EmuCam(ctypes.Structure):
    _fields_ = [("K", ctypes.c_float * 9), ("invK", ctypes.c_float * 9), ("rcd", ctypes.c_float),
                ("near_z", ctypes.c_float), ("far_z", ctypes.c_float), ("clamp_lo", ctypes.c_float),
                ("clamp_hi", ctypes.c_float), ("S", ctypes.c_int)]


_emu = None


def emu_lib():
    global _emu
    if _emu is None:
        src = os.path.join(ROOT, "tests", "emu", "emu_host.cpp")
        out = os.path.join(ROOT, "tests", "emu", "libemu.so")
        deps = [src, os.path.join(ROOT, "gan-2d-to-3d_b200", "csrc", "g2s_math.cuh"),
                os.path.join(ROOT, "gan-2d-to-3d_b200", "csrc", "g2s_raster.cuh")]
        if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
            subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, src, "-lm"])
        _emu = ctypes.CDLL(out)
    return _emu


def emu_cam(orc, S, near=0.1, far=100.0):
    cam = EmuCam()
    for i, v in enumerate(orc.K.reshape(-1).tolist()):
        cam.K[i] = v
    for i, v in enumerate(orc.inv_K.reshape(-1).tolist()):
        cam.invK[i] = v
    margin = (MAX_DEPTH - MIN_DEPTH) / 2
    cam.rcd, cam.near_z, cam.far_z = orc.rot_center_depth, near, far
    cam.clamp_lo, cam.clamp_hi, cam.S = MIN_DEPTH - margin, MAX_DEPTH + margin, S
    return cam


def vp(t):
    return ctypes.c_void_p(t.data_ptr())


def close_except_few(a, b, tol=1e-5, frac=2e-3):
    """For whole-path comparisons where R is computed by the CUDA libm (sin/cos differ from the CPU's by an ulp):
    a handful of rounding-decided sub-pixels may pick the neighbouring face, which changes isolated output pixels.
    Every other element must agree to `tol` (relative to max|b|); at most `frac` of the elements may differ more."""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    bad = (a - b).abs() > tol * b.abs().max().clamp_min(1e-30)
    return bad.double().mean().item() <= frac
