import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import g2s_b200
from g2s_b200 import synthetic, _lib
S, P, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
case = synthetic.make_case(S, P, seed=1234, n_images=N)
ren = g2s_b200.Renderer({"rot_center_depth": 1.0, "fov": 10, "tex_cube_size": 2}, S, 0.9, 1.1)
d = {k: v.cuda() for k, v in case.items()}
with torch.no_grad():
    ren.render_chain(d["depth"], d["albedo"], d["view"], d["light"], views_per_image=P)
torch.cuda.synchronize()
lib = ctypes.CDLL(os.environ["G2S_LIB"])
out = (ctypes.c_ulonglong * 2)()
lib.g2s_debug_slow_counters(out)
print("S=%d views=%d: slow hits %d (%.4f per view), slow rows %d (%.4f per view)" % (S, N * P, out[0], out[0] / (N * P), out[1], out[1] / (N * P)))
