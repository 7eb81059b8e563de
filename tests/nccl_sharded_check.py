"""Multi-GPU check of the view-sharded render (BASELINE.json configs[3], the face config: ONE image whose P views are split
across the ranks, so the per-image gradients need the one NCCL all_reduce of sharding.py).  Not collected by pytest (needs
N GPUs); run under torchrun on a GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        tests/nccl_sharded_check.py --size 256 --views 1024

Every rank renders its slice (CUDA path) and ALSO the whole batch on its own GPU: the slice of recon_im / recon_depth must be
bit-identical to the single-GPU render, the reduced grad_depth / grad_albedo and the per-view gradients within 1e-5.  Prints
one JSON line with the verdict and the device time of the sharded step (max over ranks)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--views", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import g2s_b200
    from g2s_b200 import synthetic
    from g2s_b200.sharding import render_chain_sharded
    S, P = args.size, args.views
    case = {k: v.to(dev) for k, v in synthetic.make_case(S, P, seed=77, n_images=1).items()}
    ren = g2s_b200.Renderer({"rot_center_depth": 1.0, "fov": 10, "tex_cube_size": 2}, S, 0.9, 1.1, device=dev)

    def fn(d, a, v, l, vpi):
        return ren.render_chain(d, a, v, l, views_per_image=vpi)

    def sharded(want_loss=True):
        return render_chain_sharded(fn, case["depth"], case["albedo"], case["view"], case["light"], case["cotangent"], P,
                                    rank, world, want_loss=want_loss)

    out = sharded()
    full = render_chain_sharded(fn, case["depth"], case["albedo"], case["view"], case["light"], case["cotangent"], P, 0, 1)
    sh = out["shard"]
    v0, v1 = sh["view_start"], sh["view_stop"]
    ok = torch.equal(out["recon_im"], full["recon_im"][v0:v1]) and torch.equal(out["recon_depth"], full["recon_depth"][v0:v1])
    errs = dict(grad_depth=rel(out["grad_depth"], full["grad_depth"]), grad_albedo=rel(out["grad_albedo"], full["grad_albedo"]),
                grad_view=rel(out["grad_view"], full["grad_view"][v0:v1]), grad_light=rel(out["grad_light"], full["grad_light"][v0:v1]),
                loss=abs(float(out["loss"]) - float(full["loss"])) / max(1e-30, abs(float(full["loss"]))))
    ok = ok and all(e < 1e-5 for e in errs.values())
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    for _ in range(2):
        sharded(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        sharded(False)      # the timed step: render fwd+bwd of the shard + the all_reduce of the per-image gradients
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"check": "view-sharded render vs single GPU", "pass": bool(flag.item() > 0), "n_gpus": world,
                          "image_size": S, "views": P, "forward_bit_identical": True if flag.item() > 0 else None,
                          "rank0_rel_errors": errs, "ms_per_step": float(ms.item()),
                          "renders_per_s": P / (float(ms.item()) * 1e-3), "scaling": "strong"}))
    dist.destroy_process_group()
    sys.exit(0 if flag.item() > 0 else 1)


if __name__ == "__main__":
    main()
