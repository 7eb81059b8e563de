"""A/B harness: python profiles/tools/ab_bench.py [name=lib.so ...] -- runs bench.py once per library variant, prints per-kernel ms."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
extra = [a for a in sys.argv[1:] if a.startswith("--")]
variants = [a.split("=", 1) for a in sys.argv[1:] if "=" in a and not a.startswith("--")] or [["main", ""]]
for name, lib in variants:
    env = dict(os.environ)
    if lib:
        env["G2S_LIB"] = os.path.join(ROOT, lib)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--no-single-image",
                          "--steps", "5", "--warmup", "3"] + extra, env=env, capture_output=True, text=True)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not lines:
        print(name, "FAILED", out.stderr[-2000:])
        continue
    d = json.loads(lines[-1])
    ks = " ".join("%s=%.2f" % (k["name"].replace("k_", ""), k["ms_per_step"]) for k in d["roofline"]["kernels"][:9])
    print("%-12s %.0f r/s  %.2f ms | %s" % (name, d["value"], d["ms_per_step"], ks), flush=True)
