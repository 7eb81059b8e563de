#!/usr/bin/env python
"""bench.py -- projected-view renders/s (fwd+bwd, 128^2) of the fused depth-map render, with its HBM roofline and the
reference's CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--images I] [--views P] [--size S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one pass of the hot path (chain C of SURVEY.md 8d: normal -> shading -> warp_canon_depth ->
get_inv_warped_2d_grid -> grid_sample -> clamp, forward AND backward to depth/albedo/view/light) over one batch of
synthetic input: `--images` images x `--views` pseudo-views at side `--size` per GPU.  Default = BASELINE.json
configs[1] (GAN2Shape cat config: 128^2, 16 pseudo-views per image) in its batched form, 256 images per GPU, so the
per-step working set (4.5 GB algorithmic) is far larger than the 126 MB L2 (timing rule: inputs larger than L2).

Keys of the JSON line: see the round prompt; `value` = renders/s with inputs resident in HBM (CUDA events, max over
ranks; the loss is SURVEY.md 8d's fixed device-resident cotangent on recon_im, handed straight to the backward; nothing but
the step is in the stream); `e2e` = the same through hostio.HostRenderStep with pinned HOST inputs copied in and the four
gradients copied out every step (two slots, three streams: the copies of neighbouring steps overlap the kernels, all
inside the timed region); `e2e_full` = the same with recon_im and recon_depth downloaded as well; `roofline` = the WHOLE
fwd+bwd step against the HBM roofline (renders/s/GPU x (64 S^2 + 48 S^2/P) bytes / measured peak; `kernel` =
the dominant kernel, `kernels[]` = every kernel's live time, share and own algorithmic bytes from a second pass with a
CUDA-event pair around every kernel, single forward lane; `traffic` = DRAM bytes per step from the committed warm-cache ncu
capture of this shape and code, profiles/r02_traffic.json, or null); `cpu_baseline` = the oracle (reference renderer.py on
torch-CPU + the C restatement of the external rasteriser's brute-force loop) on a bounded sample, with the bounding-box-culled
variant (bit-identical outputs) beside it as `cpu_baseline.culled`; `single_image` = the literal configs[1] shape (1 image x
16 views) eager and as a CUDA graph; `other_configs` = the car (64-yaw render_yaw sweep) and face (256^2 x 1024 views)
configs of BASELINE.json on one GPU; `loss_step` = the step with the reference's masked photometric loss (model.py:265-274)
in place of the fixed cotangent, taken inside the render (Renderer.render_chain_loss) and as the unfused composition; under
--gpus N > 1 `face_sharded` = the face config with its 1024 views split over the ranks (strong scaling, the per-image
gradients all-reduced over NCCL inside the timed region).  `--ncu-step` (profiling only) brackets one step with
cudaProfilerStart/Stop and prints no line.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFGS = {"rot_center_depth": 1.0, "fov": 10, "tex_cube_size": 2}
MIN_DEPTH, MAX_DEPTH = 0.9, 1.1
METRIC = "projected-view renders/sec (fwd+bwd)"
UNIT = "renders/s"


def alg_bytes_per_render(S, P):
    """fwd 32 S^2 + 16 S^2/P, bwd 32 S^2 + 32 S^2/P."""
    return 64.0 * S * S + 48.0 * S * S / P


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(S, views, steps, warmup, threads, mode="brute"):
    """The reference's CPU implementation of the path = the oracle: reference renderer.py arithmetic on torch-CPU plus
    the faithful O((2S)^2 * 4(S-1)^2) rasteriser loop (oracle/nr_raster.c, OpenMP).  Returns renders/s over `steps`
    steps of `views` views of one image (fwd+bwd)."""
    import torch
    import torch.nn.functional as F
    import g2s_b200
    from g2s_b200 import synthetic
    from oracle import nr_port, renderer_oracle as ro
    torch.set_num_threads(threads)
    nr_port.lib().nr_set_threads(threads)      # torchrun exports OMP_NUM_THREADS=1: pin the OpenMP team explicitly
    nr_port.MODE["raster"] = mode          # "brute": the reference's face loop; "culled": bounding-box culled, same outputs
    case = synthetic.make_case(S, views, seed=1234)
    orc = ro.OracleRenderer(dict(CFGS), S, MIN_DEPTH, MAX_DEPTH)

    def step():
        depth = case["depth"].clone().requires_grad_(True)
        albedo = case["albedo"].clone().requires_grad_(True)
        view = case["view"].clone().requires_grad_(True)
        light = case["light"].clone().requires_grad_(True)
        out = orc.render_chain(depth, albedo, view, light)
        (out["recon_im"] * case["cotangent"]).sum().backward()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    nr_port.MODE["raster"] = "culled"
    return views * steps / dt, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S, P = args.size, args.views
    threads = os.cpu_count() or 1
    views = 8 if S <= 128 else 2          # ~7 s (128^2) / ~30 s (256^2) of brute-force rasterisation per step on 16 cores
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    value, dt = cpu_reference_run(S, views, steps, warmup, threads)
    sample = "%d step(s) x %d view(s) of one %dx%d image, fwd+bwd, brute-force face loop" % (steps, views, S, S)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, per_gpu_images=args.images),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, per_gpu_images):
    return {"workload": "GAN2Shape cat config (BASELINE.json configs[1]): %dx%d depth, %d projected pseudo-views per "
                        "image, fwd+bwd; batched form, %d images per GPU per step" % (args.size, args.size, args.views,
                                                                                      per_gpu_images),
            "image_size": args.size, "views_per_image": args.views, "images_per_gpu": per_gpu_images,
            "align_corners": False,
            "l2_policy": "inputs larger than L2 (per-step working set >> 126 MB); no explicit flush",
            "cotangent": "device-resident fixed random d(loss)/d(recon_im) (stands in for the on-device losses of "
                         "model.py:274-278)"}


# ---------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import g2s_b200
    from g2s_b200 import synthetic, _lib, build as _build

    rank = int(os.environ.get("RANK", "0"))
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        _build.build()      # no-op when the in-tree library is up to date (building the product is not a fallback)
    else:
        for _ in range(240):            # other ranks wait for rank 0's (re)build before loading the library
            if not _build.needs_build():
                break
            time.sleep(0.5)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    S, P, N = args.size, args.views, args.images
    B = N * P
    case = synthetic.make_case(S, P, seed=1234 + rank, n_images=N)
    ren = g2s_b200.Renderer(dict(CFGS), S, MIN_DEPTH, MAX_DEPTH, device=dev)
    host = {k: case[k].pin_memory() for k in ("depth", "albedo", "view", "light")}
    cot = case["cotangent"].to(dev)
    d_in = {k: v.to(dev) for k, v in host.items()}
    loss_buf = torch.zeros((), device=dev)

    def step_device():
        depth = d_in["depth"].requires_grad_(True)
        albedo = d_in["albedo"].requires_grad_(True)
        view = d_in["view"].requires_grad_(True)
        light = d_in["light"].requires_grad_(True)
        for tns in (depth, albedo, view, light):
            tns.grad = None
        recon_im, recon_depth, fidx = ren.render_chain(depth, albedo, view, light, views_per_image=P)
        # the loss is a fixed device-resident cotangent on recon_im -- handed straight to the backward
        # (no `(recon_im * cot).sum()` glue kernels in the timed region).  Whole images per rank: nothing to exchange.
        torch.autograd.backward([recon_im], [cot])

    # e2e: the same step through the host-buffer front end (hostio.HostRenderStep): every step uploads its inputs from
    # pinned host memory and downloads its results (the four gradients) to pinned host memory; uploads / downloads of
    # neighbouring steps overlap the kernels on separate streams, all inside the timed region.
    from g2s_b200 import hostio
    hstep = hostio.HostRenderStep(ren, N, P, cot)
    hfull = hostio.HostRenderStep(ren, N, P, cot, outputs=("recon_im", "recon_depth"))

    def step_e2e():
        hstep.submit(host)

    h2d = sum(host[k].numel() * 4 for k in host)
    d2h = hstep.d2h_bytes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tns = torch.tensor([ms], device=dev)
            dist.all_reduce(tns, op=dist.ReduceOp.MAX)
            ms = float(tns.item())
        return ms

    for _ in range(args.warmup):
        step_device()
    if args.ncu_step:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        step_device()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = lib.g2s_launch_count()
    ms_total = timed(step_device, args.steps)          # the reported number: no per-kernel events in the stream
    launches = lib.g2s_launch_count() - launches0
    lib.g2s_profile_enable(1)                          # second pass, same steps, with a CUDA-event pair around every kernel
    ms_profiled = timed(step_device, args.steps)
    lib.g2s_profile_enable(0)
    clocks = sampler.stop() if sampler else None

    names = (ctypes.c_char_p * 32)()
    tot = (ctypes.c_float * 32)()
    cnt = (ctypes.c_int * 32)()
    nk = lib.g2s_profile_read(32, names, tot, cnt)
    kernels = [{"name": names[i].decode(), "launches": int(cnt[i]), "ms_per_launch": tot[i] / cnt[i],
                "ms_per_step": tot[i] / args.steps} for i in range(nk)]

    def e2e_time(hs):
        def join():
            hs.drain()                         # host waits for the last downloads ...
            cur = torch.cuda.current_stream()
            for st_ in (hs.h2d, hs.comp, hs.d2h):
                cur.wait_stream(st_)           # ... and the timing stream is ordered after all three pipelines

        for _ in range(max(1, min(args.warmup, 3))):
            hs.submit(host)
        join()

        def run():
            for st_ in (hs.h2d, hs.comp, hs.d2h):
                st_.wait_stream(torch.cuda.current_stream())    # nothing starts before the opening event
            for _ in range(args.steps):
                hs.submit(host)
            join()

        return timed(run, 1)

    ms_e2e = e2e_time(hstep)
    ms_e2e_full = e2e_time(hfull)      # + recon_im and recon_depth downloaded every step
    d2h_full = hfull.d2h_bytes
    del hfull

    # BASELINE.json configs[3] under N > 1 ranks: ONE 256^2 image x 1024 views, the views split over the ranks (strong
    # scaling); every step ends in the NCCL all_reduce of grad_depth + grad_albedo, inside the timed region
    face_sharded = None
    if world > 1 and not args.no_single_image:
        from g2s_b200.sharding import render_chain_sharded
        FS, FP = 256, 1024
        fcase = {k: v.to(dev) for k, v in synthetic.make_case(FS, FP, seed=11, n_images=1).items()}
        fren = g2s_b200.Renderer(dict(CFGS), FS, MIN_DEPTH, MAX_DEPTH, device=dev)

        def ffn(d, a, v, l, vpi):
            return fren.render_chain(d, a, v, l, views_per_image=vpi)

        def fstep():
            render_chain_sharded(ffn, fcase["depth"], fcase["albedo"], fcase["view"], fcase["light"], fcase["cotangent"], FP,
                                 rank, world, want_loss=False)

        for _ in range(3):
            fstep()
        ms_f = timed(fstep, 10) / 10
        face_sharded = {"workload": "face config: one 256x256 image x 1024 views split over %d ranks, fwd+bwd, gradients "
                                    "all-reduced (NCCL) every step" % world, "scaling": "strong", "ms_per_step": ms_f,
                        "value": FP / (ms_f * 1e-3), "unit": UNIT, "views_per_rank": FP // world,
                        "all_reduce_bytes_per_step": 16 * FS * FS}
        del fcase, fren

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms_total / args.steps
    renders = B * world
    value = renders / (ms_step * 1e-3)
    e2e_value = renders / (ms_e2e / args.steps * 1e-3)
    peak, peak_src = measured_peak_gbs()
    per_render = alg_bytes_per_render(S, P)
    # per-kernel algorithmic bytes per STEP (DESIGN.md "kernels"): the HBM traffic each kernel cannot avoid, i.e. its
    # share of SURVEY.md 8(d)'s per-render budget; per launch = per step / launches per step (chunked launches)
    S2 = float(S * S)
    kb = {"k_normal_fwd": N * 4 * S2,                              # reads depth; the normal map is L2-resident scratch
          "k_splat_tile": N * 4 * S2,                              # reads the depth maps; the z-buffer is L2 scratch
          "k_resolve_fused": B * 32 * S2 + N * 12 * S2,            # writes recon_depth+recon_im+face_idx; reads albedo
          "k_render_bwd_pixel": B * (12 + 4) * S2 + N * 12 * S2,   # reads grad_recon_im, recon_depth, albedo
          "k_render_bwd_tex": N * 12 * S2,                         # writes grad_albedo
          "k_normal_bwd": N * (4 + 4) * S2,                        # reads depth, writes grad_depth
          "k_raster_bwd_px": B * 16 * S2,                          # reads the face-index map
          "k_project_verts": N * 4 * S2, "k_vertex_bwd": N * (4 + 4) * S2}   # depth in, grad_depth out (scratch in L2)
    for k in kernels:
        per_step = kb.get(k["name"])
        k["share_of_step"] = k["ms_per_step"] / ms_step
        if per_step:
            k["alg_bytes_per_launch"] = per_step * args.steps / k["launches"]
            k["achieved_gbs"] = per_step / (k["ms_per_step"] * 1e-3) / 1e9
            k["frac"] = k["achieved_gbs"] / peak
    kernels.sort(key=lambda k: -k["ms_per_step"])
    dom = kernels[0] if kernels else None
    # DRAM bytes the whole step actually moves, per step of this run: from the committed warm-cache ncu capture of the SAME
    # shape (side, views per image, views per launch) and the same kernels -- profiles/r02_traffic.json, written by
    # profiles/tools/traffic_from_ncu.py -- or null when the shape differs
    traffic, ncu_ctx = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            tj = json.load(f)
        if tj["image_size"] == S and tj["views_per_image"] == P:
            traffic = tj["dram_bytes_per_render"] * B
            ncu_ctx = {"source": tj["source"], "dram_bytes_per_render": tj["dram_bytes_per_render"],
                       "alg_bytes_per_render": per_render, "ratio_to_algorithmic": tj["dram_bytes_per_render"] / per_render,
                       "per_kernel_dram_bytes_per_render": tj.get("kernels")}
    except Exception:
        traffic = None
    step_achieved = value / world * per_render / 1e9
    roofline = {"bound": "hbm", "kernel": dom["name"] if dom else None,
                "achieved": step_achieved, "peak": peak, "unit": "GB/s", "frac": step_achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_render": per_render, "alg_bytes_per_step": per_render * B,
                "kernel_share_of_step": dom["share_of_step"] if dom else None,
                "kernel_achieved": dom.get("achieved_gbs") if dom else None, "kernel_frac": dom.get("frac") if dom else None,
                "ncu": ncu_ctx,
                "note": "achieved / frac: the WHOLE fwd+bwd step, SURVEY.md 8(d): renders/s/GPU x (64 S^2 + 48 S^2/P) bytes over the "
                        "measured HBM peak.  `kernel` = the kernel with the largest share of the step (per-kernel live times, "
                        "shares and own algorithmic bytes under kernels[]).  The rasteriser kernels are instruction-issue bound "
                        "by construction (bit-exact reproduction of the reference's un-fused fp32 arithmetic; z-buffer and "
                        "scratch live in L2): the step sits far below the HBM roofline, see DESIGN.md section 7",
                "kernels": kernels}

    # the literal BASELINE.json configs[1] shape, one image x P views per step: launch-bound, so also as a CUDA graph
    single = None
    if world == 1 and not args.no_single_image:
        try:
            from g2s_b200 import graphs
            one = synthetic.make_case(S, P, seed=99, n_images=1)
            one = {k: v.to(dev) for k, v in one.items()}
            gstep = graphs.GraphedRenderStep(ren, 1, P)
            gstep.step(one["depth"], one["albedo"], one["view"], one["light"], one["cotangent"])

            def eager_one():
                d1 = one["depth"].requires_grad_(True)
                a1 = one["albedo"].requires_grad_(True)
                im1 = ren.render_chain(d1, a1, one["view"], one["light"], views_per_image=P)[0]
                torch.autograd.grad([im1], [d1, a1], grad_outputs=[one["cotangent"]])

            for _ in range(5):
                eager_one()
            reps = 100
            ms_g = timed(lambda: gstep.step(), reps) / reps
            ms_e = timed(eager_one, reps) / reps
            single = {"workload": "1 image x %d views at %dx%d, fwd+bwd (literal configs[1] shape; launch-bound)" % (P, S, S),
                      "cuda_graph": {"value": P / (ms_g * 1e-3), "unit": UNIT, "ms_per_step": ms_g},
                      "eager": {"value": P / (ms_e * 1e-3), "unit": UNIT, "ms_per_step": ms_e}}
        except Exception as exc:   # the headline numbers do not depend on this extra
            single = {"error": str(exc)[:200]}

    # the other BASELINE.json configs that fit one GPU, measured in the same run (reported, not the headline):
    #   configs[2] car: render_yaw sweep of 64 yaw angles at 128^2, forward only, both branches (renderer.py:141-198)
    #   configs[3] face: ONE 256^2 image x 1024 pseudo-views, fwd+bwd (N > 1: tests/nccl_sharded_check.py shards its views)
    other = None
    if world == 1 and not args.no_single_image:
        try:
            with torch.no_grad():
                car = synthetic.make_case(128, 1, seed=7, n_images=1)
                ren128 = ren if S == 128 else g2s_b200.Renderer(dict(CFGS), 128, MIN_DEPTH, MAX_DEPTH, device=dev)
                im_c, d_c = car["albedo"].to(dev), car["depth"].to(dev)
                for gs in (False, True):
                    ren128.render_yaw(im_c, d_c, maxr=90, nsample=64, grid_sample=gs)
                reps = 50
                ms_mesh = timed(lambda: ren128.render_yaw(im_c, d_c, maxr=90, nsample=64), reps) / reps
                ms_gs = timed(lambda: ren128.render_yaw(im_c, d_c, maxr=90, nsample=64, grid_sample=True), reps) / reps
            face = {k: v.to(dev) for k, v in synthetic.make_case(256, 1024, seed=11, n_images=1).items()}
            ren256 = g2s_b200.Renderer(dict(CFGS), 256, MIN_DEPTH, MAX_DEPTH, device=dev)

            def face_step():
                d_, a_ = face["depth"].requires_grad_(True), face["albedo"].requires_grad_(True)
                v_, l_ = face["view"].requires_grad_(True), face["light"].requires_grad_(True)
                im_ = ren256.render_chain(d_, a_, v_, l_, views_per_image=1024)[0]
                torch.autograd.grad([im_], [d_, a_, v_, l_], grad_outputs=[face["cotangent"]])

            for _ in range(3):
                face_step()
            ms_face = timed(face_step, 5) / 5
            other = {"car_render_yaw_64_at_128": {"mesh_texture_branch": {"ms_per_sweep": ms_mesh, "value": 64 / (ms_mesh * 1e-3)},
                                                  "grid_sample_branch": {"ms_per_sweep": ms_gs, "value": 64 / (ms_gs * 1e-3)},
                                                  "unit": "yaw renders/s (forward only)"},
                     "face_1024_views_at_256": {"ms_per_step": ms_face, "value": 1024 / (ms_face * 1e-3), "unit": UNIT}}
            del face, ren256
        except Exception as exc:   # the headline numbers do not depend on these extras
            other = {"error": str(exc)[:200]}

    # SURVEY.md 8(f) row 1 measured: the step WITH the reference's step-3 photometric loss (model.py:265-274) on the bench
    # shape -- taken inside the render (Renderer.render_chain_loss: resolve epilogue + loss-driven pixel backward) against
    # the unfused composition render_chain -> PhotometricLoss (three more passes over recon_im / target / recon_depth)
    loss_step = None
    if world == 1 and not args.no_single_image:
        try:
            gen = torch.Generator().manual_seed(5)
            target = (torch.rand(B, 3, S, S, generator=gen) * 2 - 1).to(dev)
            vmask = (torch.rand(B, 1, S, S, generator=gen) > 0.2).float().to(dev)
            photo = g2s_b200.PhotometricLoss()

            def leaves():
                ts = [d_in[k].requires_grad_(True) for k in ("depth", "albedo", "view", "light")]
                for tns in ts:
                    tns.grad = None
                return ts

            def step_fused():
                depth, albedo, view, light = leaves()
                ren.render_chain_loss(depth, albedo, view, light, target, vmask, views_per_image=P)[0].backward()

            def step_unfused():
                depth, albedo, view, light = leaves()
                im_, rd_, _ = ren.render_chain(depth, albedo, view, light, views_per_image=P)
                photo(im_, target, mask=vmask, **g2s_b200.recon_im_mask(rd_.detach(), MIN_DEPTH, MAX_DEPTH)).backward()

            for fn in (step_fused, step_unfused):
                for _ in range(2):
                    fn()
            reps = max(3, args.steps // 3)
            ms_fu, ms_un = timed(step_fused, reps) / reps, timed(step_unfused, reps) / reps
            loss_step = {"workload": "the bench step with the masked photometric loss of model.py:265-274 instead of a fixed "
                                     "cotangent", "fused": {"value": B / (ms_fu * 1e-3), "unit": UNIT, "ms_per_step": ms_fu},
                         "unfused": {"value": B / (ms_un * 1e-3), "unit": UNIT, "ms_per_step": ms_un}}
            del target, vmask
        except Exception as exc:   # the headline numbers do not depend on this extra
            loss_step = {"error": str(exc)[:200]}

    threads = os.cpu_count() or 1
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        views = 8 if S <= 128 else 1
        v, dt = cpu_reference_run(S, views, 2, 0, threads)
        cviews = 64 if S <= 128 else 16
        vc, dtc = cpu_reference_run(S, cviews, 3, 1, threads, mode="culled")
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "2 steps x %d view(s) of one %dx%d image, fwd+bwd, oracle with the reference's "
                                  "brute-force face loop (%.1f s)" % (views, S, S, dt),
                        "culled": {"value": vc, "unit": UNIT, "cores": threads,
                                   "sample": "3 steps x %d views, the same oracle with the bounding-box-culled face loop "
                                             "(bit-identical outputs; NOT the reference's algorithm) (%.1f s)" % (cviews, dtc)}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "ms_per_step_with_kernel_events": ms_profiled / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, N), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": ms_e2e / args.steps, "returns": "the four gradients"},
            "e2e_full": {"value": renders / (ms_e2e_full / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                         "d2h_bytes_per_step": d2h_full * world, "ms_per_step": ms_e2e_full / args.steps,
                         "returns": "the four gradients + recon_im + recon_depth"},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "single_image": single, "other_configs": other, "loss_step": loss_step, "face_sharded": face_sharded}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--views", type=int, default=16, help="projected pseudo-views per image")
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-single-image", action="store_true", help="skip the 1-image CUDA-graph extra (profiling runs)")
    ap.add_argument("--ncu-step", action="store_true",
                    help="profiling runs only: after the warm-up, bracket ONE step with cudaProfilerStart/Stop and exit "
                         "(ncu --profile-from-start off ...); prints no bench line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
