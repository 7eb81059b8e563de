"""View / image sharding across the GPUs of one box: one process per GPU, no data-path collective.

Every (image, view) render is independent, so a rank renders a contiguous slice of the views.  The only exchange is
the sum of the per-image gradients (grad_depth [S,S], grad_albedo [3,S,S]) and of the scalar loss when ONE image's
views are split across ranks; when whole images are sharded (the bulk configuration) nothing is exchanged at all.
The reference has no multi-GPU path for the renderer; this is the host-side plumbing the bench and a
data-parallel caller use.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous balanced partition: the first n_items % world_size ranks get one extra item."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_views(n_images, views_per_image, rank, world_size):
    """Which slice of the flat view list [n_images * views_per_image] a rank renders.

    Returns dict(view_start, view_stop, image_start, image_stop, split_images) where split_images lists the images
    whose views are shared with another rank (their gradients need the all-reduce).  Whole images are kept together
    whenever there are at least as many images as ranks."""
    if n_images >= world_size:
        i0, i1 = shard_range(n_images, rank, world_size)
        return dict(view_start=i0 * views_per_image, view_stop=i1 * views_per_image, image_start=i0, image_stop=i1,
                    split_images=[])
    v0, v1 = shard_range(n_images * views_per_image, rank, world_size)
    if v1 == v0:
        return dict(view_start=v0, view_stop=v1, image_start=0, image_stop=0, split_images=[])
    i0, i1 = v0 // views_per_image, (v1 - 1) // views_per_image + 1
    split = [i for i in range(i0, i1) if v0 > i * views_per_image or v1 < (i + 1) * views_per_image]
    return dict(view_start=v0, view_stop=v1, image_start=i0, image_stop=i1, split_images=split)


def any_rank_splits(n_images, views_per_image, world_size):
    """Does ANY rank share an image with another rank?  Every rank can answer this locally from the partition (no
    collective, no host sync): whole images are kept together whenever there are at least as many images as ranks."""
    if n_images >= world_size:
        return False
    return any(shard_views(n_images, views_per_image, r, world_size)["split_images"] for r in range(world_size))


def reduce_image_grads(grads, n_images, group=None, extra=None, async_op=False):
    """Sum per-image gradients over the ranks that share images: ONE all_reduce of one flat buffer (the tensors of `grads`,
    whose dim 0 is the FULL image axis [n_images, ...] with zero rows for the images a rank did not touch, plus the optional
    scalar tensors in `extra`) -- 16 S^2 bytes per image, <= 1 MB at 256^2: launch latency, not bandwidth, is what it costs,
    so one call instead of one per tensor.  Returns a `finish()` callable that waits for the collective (when async_op) and
    copies the sums back into the tensors; no-op when not initialised / single rank."""
    extra = list(extra or [])
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return lambda: grads
    for g in grads:
        assert g.shape[0] == n_images
    parts = list(grads) + extra
    flat = torch.cat([p.reshape(-1).float() for p in parts])
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def finish():
        if async_op and work is not None:
            work.wait()
        off = 0
        for p in parts:
            n = p.numel()
            p.copy_(flat[off:off + n].reshape(p.shape))
            off += n
        return grads

    return finish


def render_chain_sharded(render_fn, depth, albedo, view, light, cotangent, views_per_image, rank, world_size,
                         group=None, want_loss=True):
    """Data-parallel fwd+bwd of the fused render: each rank renders its slice of the views with `render_fn`
    (= Renderer.render_chain; injected so the host logic can be tested on CPU with the oracle) and the per-image
    gradients are summed across ranks only for images whose views span ranks.

    Returns dict(recon_im, recon_depth [local views], grad_depth, grad_albedo [n_images, ...] (reduced), grad_view,
    grad_light [local views], loss (global sum of <recon_im, cotangent>; skipped with want_loss=False), shard)."""
    n_images = depth.shape[0]
    sh = shard_views(n_images, views_per_image, rank, world_size)
    v0, v1, i0, i1 = sh["view_start"], sh["view_stop"], sh["image_start"], sh["image_stop"]
    # The per-image gradients live in ONE flat buffer from the start (grad_depth | grad_albedo | loss), so the exchange is one
    # all_reduce IN PLACE and the results are views of it: at 8 ranks a face-config step is 2 ms, and every extra small
    # kernel (clone, zeros, cat, copy back) is 3-5 us of it.
    nd, na = depth.numel(), albedo.numel()
    flat = torch.zeros(nd + na + 1, device=depth.device, dtype=torch.float32)
    g_depth, g_albedo, loss = flat[:nd].view(depth.shape), flat[nd:nd + na].view(albedo.shape), flat[nd + na]
    out = dict(shard=sh, recon_im=None, recon_depth=None, grad_view=None, grad_light=None)
    if v1 > v0:
        d = depth[i0:i1].detach().requires_grad_(True)
        a = albedo[i0:i1].detach().requires_grad_(True)
        vw = view[v0:v1].detach().requires_grad_(True)
        lt = light[v0:v1].detach().requires_grad_(True)
        if sh["split_images"] or (v1 - v0) % views_per_image:
            # views of a single image (or of images cut at the shard boundary): render image by image
            ims, rds = [], []
            for i in range(i0, i1):
                a0, a1 = max(v0, i * views_per_image), min(v1, (i + 1) * views_per_image)
                im, rd = render_fn(d[i - i0:i - i0 + 1], a[i - i0:i - i0 + 1], vw[a0 - v0:a1 - v0], lt[a0 - v0:a1 - v0],
                                   a1 - a0)[:2]
                ims.append(im)
                rds.append(rd)
            recon_im, recon_depth = (torch.cat(ims, 0), torch.cat(rds, 0)) if len(ims) > 1 else (ims[0], rds[0])
        else:
            recon_im, recon_depth = render_fn(d, a, vw, lt, views_per_image)[:2]
        cot = cotangent[v0:v1]
        torch.autograd.backward([recon_im], [cot])        # the cotangent goes straight to the backward
        g_depth[i0:i1].copy_(d.grad)
        g_albedo[i0:i1].copy_(a.grad)
        if want_loss:
            loss.copy_((recon_im.detach() * cot).sum())
        out.update(recon_im=recon_im.detach(), recon_depth=recon_depth.detach(), grad_view=vw.grad, grad_light=lt.grad)
    if world_size > 1 and dist.is_initialized():
        # decided locally (round 1 all-reduced a flag and read it back on the host every step)
        if any_rank_splits(n_images, views_per_image, world_size):
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        elif want_loss:
            dist.all_reduce(flat[nd + na:], op=dist.ReduceOp.SUM, group=group)
    out.update(grad_depth=g_depth, grad_albedo=g_albedo, loss=loss)
    return out
