"""CPU-side checks: the C-ABI library loads and exports every symbol include/g2s_b200.h declares (no compute calls
without a GPU), argument validation returns error codes, host-side mirror of utils.py matches the oracle."""
import ctypes
import os
import re

import pytest
import torch

from helpers import CFGS, MAX_DEPTH, MIN_DEPTH, ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "g2s_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(g2s_[a-z0-9_]+)\s*\(", hdr)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    import g2s_b200
    from g2s_b200 import _lib, build
    build.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(lib, name), name
        assert name in _lib.SIGNATURES, "ctypes prototype missing for " + name
    assert set(_lib.SIGNATURES) == set(declared)
    assert lib.g2s_version() >= 100
    assert lib.g2s_error_string(0) == b"ok"


def _declared_prototypes():
    """name -> list of parameter type strings, from the header (comments stripped)"""
    hdr = open(os.path.join(ROOT, "include", "g2s_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|long|size_t|void|const char \*)\s*\*?\s*(g2s_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        params = [p.strip() for p in m.group(2).replace("\n", " ").split(",")]
        protos[m.group(1)] = [] if params == ["void"] else params
    return protos


def test_ctypes_prototypes_match_the_header_argument_for_argument():
    """a drifted ctypes signature (one pointer too few, an int where the header has a long) corrupts the call silently:
    every prototype of _lib.SIGNATURES is held against the header's parameter list -- count and kind"""
    from g2s_b200 import _lib
    protos = _declared_prototypes()
    assert set(protos) == set(_lib.SIGNATURES)

    def kind(decl):
        if "*" in decl:
            return "ptr"
        base = decl.split()[:-1]
        if "long" in base or "size_t" in base:
            return "long"
        if "float" in base:
            return "float"
        return "int"

    def ckind(t):
        if t in (ctypes.c_long, ctypes.c_ulong, ctypes.c_size_t, ctypes.c_longlong, ctypes.c_ulonglong):
            return "long"
        if t is ctypes.c_float:
            return "float"
        if t in (ctypes.c_int, ctypes.c_uint, ctypes.c_int32):
            return "int"
        return "ptr"

    for name, params in protos.items():
        argtypes = _lib.SIGNATURES[name][1]
        assert len(argtypes) == len(params), (name, len(argtypes), len(params))
        for k, (decl, t) in enumerate(zip(params, argtypes)):
            assert kind(decl) == ckind(t), (name, k, decl, t)


def test_ctypes_structs_match_the_header_field_for_field():
    """struct g2s_camera / g2s_photo_loss: field names, order, array lengths and sizes as the header declares them"""
    from g2s_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "g2s_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)

    def fields(body):
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(.*?)(\w+)(?:\[(\d+)\])?$", decl)
            out.append((m.group(2), "*" in m.group(1), int(m.group(3) or 1)))
        return out

    cam = fields(re.search(r"typedef struct g2s_camera \{(.*?)\} g2s_camera;", hdr, flags=re.S).group(1))
    loss = fields(re.search(r"typedef struct \{(.*?)\} g2s_photo_loss;", hdr, flags=re.S).group(1))
    for decl, struct in ((cam, _lib.Camera), (loss, _lib.PhotoLoss)):
        assert [f[0] for f in decl] == [f[0] for f in struct._fields_]
        for (name, is_ptr, n), (_, ctype) in zip(decl, struct._fields_):
            if is_ptr:
                assert ctype is ctypes.c_void_p, name
            else:
                assert ctypes.sizeof(ctype) == 4 * n, name
    assert ctypes.sizeof(_lib.Camera) == 4 * (9 + 9 + 5 + 1 + 9)
    assert ctypes.sizeof(_lib.PhotoLoss) == 24          # two pointers, a float, padding


def test_argument_validation_returns_error_codes_without_touching_the_gpu():
    from g2s_b200 import _lib
    lib = _lib.load()
    cam = _lib.Camera()
    cam.image_size = 16
    null = ctypes.c_void_p(None)
    # keys [3, 32, 32] + work list [3 * 4 * 15^2] + counter words (8 + 8 per lane part), 8 bytes each
    assert lib.g2s_zbuffer_bytes(3, 16) == (3 * (32 * 32 + 4 * 15 * 15) + 8 + 8 * 4) * 8
    assert lib.g2s_zbuffer_bytes(0, 16) == 0
    assert lib.g2s_zbuffer_init(null, 1, 16, 100.0, null) == -1
    assert lib.g2s_warp_depth_fwd(ctypes.byref(cam), null, 0, null, null, 1, null, null, null, null) == -1
    assert lib.g2s_normal_fwd(ctypes.byref(cam), null, 1, 4, 4, null, null) == -1
    assert lib.g2s_sample_fwd(null, 0, null, 1, 1, 4, 4, 4, 4, 0, 0, null, null) == -1
    one = ctypes.c_void_p(8)     # non-NULL dummy: shape checks come before any dereference / launch
    assert lib.g2s_warp_depth_fwd(ctypes.byref(cam), one, 0, one, one, 0, one, one, one, null) == -2
    assert lib.g2s_sample_fwd(one, 0, one, 1, 1, 4, 4, 4, 4, 7, 0, one, null) == -4
    bg = (ctypes.c_float * 4)(1, 1, 1, 1)
    assert lib.g2s_render_rgb_fwd(ctypes.byref(cam), one, one, 0, 1, 3, 1, bg, 1, one, one, null, null) == -4
    assert b"NULL" in lib.g2s_error_string(-1)


def test_host_mirror_of_utils_matches_oracle():
    import g2s_b200
    from oracle import renderer_oracle as ro
    torch.manual_seed(0)
    for w in (3, 5, 6):
        view = torch.randn(4, w) * 0.3
        R, t = g2s_b200.get_transform_matrices(view)
        Ro, to = ro.get_transform_matrices(view)
        assert torch.allclose(R, Ro, atol=1e-6) and torch.equal(t, to)
    with pytest.raises(Exception):
        g2s_b200.get_transform_matrices(torch.zeros(2, 4))
    assert torch.equal(g2s_b200.get_face_idx(2, 5, 6), ro.get_face_idx(2, 5, 6))
    assert torch.equal(g2s_b200.get_grid(2, 4, 5, normalize=False), ro.get_grid(2, 4, 5, normalize=False))
    light = torch.rand(3, 4) * 2 - 1
    for a, b in zip(g2s_b200.get_lighting_directions(light), ro.get_lighting_directions(light)):
        assert torch.equal(a, b)
    im = torch.rand(2, 3, 5, 6)
    for tx in (1, 2):
        assert torch.allclose(g2s_b200.get_textures_from_im(im, tx), ro.get_textures_from_im(im, tx), atol=1e-6)
    with pytest.raises(NotImplementedError):
        g2s_b200.get_textures_from_im(im, 3)
    x = torch.rand(10)
    assert float(g2s_b200.mm_normalize(x, 2, 5).min()) == 2.0 and abs(float(g2s_b200.mm_normalize(x, 2, 5).max()) - 5.0) < 1e-6


def test_renderer_constructs_on_cpu_and_refuses_cpu_compute():
    import g2s_b200
    from oracle import renderer_oracle as ro
    ren = g2s_b200.Renderer(dict(CFGS), 32, MIN_DEPTH, MAX_DEPTH, device="cpu")
    orc = ro.OracleRenderer(dict(CFGS), 32, MIN_DEPTH, MAX_DEPTH)
    assert torch.equal(ren.K, orc.K) and torch.equal(ren.inv_K, orc.inv_K)
    cam = ren._camera(depth_pass=True)
    assert abs(cam.clamp_lo - 0.8) < 1e-6 and abs(cam.clamp_hi - 1.2) < 1e-6 and cam.image_size == 32
    assert abs(cam.far_z - 100.0) < 1e-6 and abs(ren._camera(rgb_pass=True).far_z - 10.0) < 1e-6
    ren.downscale_K(2)
    orc.downscale_K(2)
    assert torch.allclose(ren.K, orc.K) and torch.allclose(ren.inv_K, orc.inv_K)
    assert list(ren._camera(depth_pass=True).K) == list(orc.K_origin.reshape(-1).tolist())   # rasteriser keeps K
    ren.set_transform_matrices(torch.zeros(1, 6))
    with pytest.raises(RuntimeError):
        ren.warp_canon_depth(torch.ones(1, 32, 32))
    with pytest.raises(RuntimeError):
        ren.get_normal_from_depth(torch.ones(1, 32, 32))


def test_mesh_export_writers(tmp_path):
    """mesh_export: the OBJ / PLY writers and the plotly array helpers (host-side; SURVEY.md 8f row 4, plotting.py:58-131)"""
    import numpy as np
    import g2s_b200
    from g2s_b200 import mesh_export as me
    S = 5
    faces = g2s_b200.get_face_idx(1, S, S)[0].numpy().astype(np.int32)
    rng = np.random.default_rng(0)
    mesh = dict(vertices=rng.standard_normal((S * S, 3)).astype(np.float32), faces=faces,
                colors=rng.random((S * S, 3)).astype(np.float32))
    obj, ply = tmp_path / "m.obj", tmp_path / "m.ply"
    me.write_obj(str(obj), mesh)
    me.write_ply(str(ply), mesh)
    lines = obj.read_text().splitlines()
    vs = [l for l in lines if l.startswith("v ")]
    fs = [l for l in lines if l.startswith("f ")]
    assert len(vs) == S * S and len(fs) == 2 * (S - 1) ** 2
    assert [int(t) for t in fs[0].split()[1:]] == [int(i) + 1 for i in faces[0]]
    assert np.allclose([float(t) for t in vs[3].split()[1:4]], mesh["vertices"][3], rtol=1e-6)
    raw = ply.read_bytes()
    head, body = raw.split(b"end_header\n")
    assert b"element vertex %d" % (S * S) in head and b"element face %d" % len(faces) in head
    assert len(body) == S * S * (12 + 3) + len(faces) * (1 + 12)
    kw = me.mesh3d_arrays(mesh)
    assert set(kw) == {"x", "y", "z", "i", "j", "k", "vertexcolor"} and len(kw["i"]) == len(faces)
    z, col = me.surface_arrays(torch.full((1, S, S), 0.9), torch.zeros(1, 3, S, S))
    assert z.shape == (S, S) and float(z[0, 0]) == -np.float32(0.9) and col.shape == (S, S)


def test_header_is_plain_c_and_a_c_host_links_and_runs(tmp_path):
    """include/g2s_b200.h compiles as C99 (no C++ / torch types at the boundary) and a C program linked against the library
    calls it -- the binding a non-Python host would use.  No device needed for these entry points."""
    import shutil
    import subprocess
    from g2s_b200 import build
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    build.build()
    lib = os.path.join(ROOT, "gan-2d-to-3d_b200", "csrc", "libg2s_b200.so")
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "g2s_b200.h"
int main(void) {
    g2s_camera cam;
    memset(&cam, 0, sizeof cam);
    cam.image_size = 16;
    if (g2s_version() < 200) return 1;
    if (strcmp(g2s_error_string(0), "ok") != 0) return 2;
    if (g2s_workspace_bytes(G2S_WS_TEXELS, 3, 16) != 3u * 8u * 16u * 16u * 4u) return 3;
    if (g2s_warp_depth_fwd(&cam, NULL, 0, NULL, NULL, 1, NULL, NULL, NULL, NULL) != -1) return 4;   /* NULL -> error code */
    if (g2s_chunk_views(128) != 512) return 5;
    printf("c host ok %d\n", g2s_version());
    return 0;
}
''')
    exe = tmp_path / "host"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           lib, "-Wl,-rpath," + os.path.dirname(lib)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "c host ok" in out.stdout
