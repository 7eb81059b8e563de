"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes shard the views, render their slice (with the oracle
standing in for the CUDA render: test infrastructure only) and reduce the per-image gradients; the result must equal
the single-process render."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import oracle_renderer, rel_err


def test_shard_range_and_views():
    from g2s_b200.sharding import shard_range, shard_views
    for n in (0, 1, 7, 16, 33):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    # whole images per rank when there are enough images: nothing to reduce
    s = shard_views(4096, 32, 3, 8)
    assert s["split_images"] == [] and s["view_stop"] - s["view_start"] == 512 * 32
    # one image, 1024 views over 8 ranks (BASELINE face config): every rank shares image 0
    s = shard_views(1, 1024, 5, 8)
    assert (s["view_start"], s["view_stop"]) == (640, 768) and s["split_images"] == [0]
    # 3 images x 4 views over 2 ranks: image 1 is cut in the middle
    a, b = shard_views(3, 4, 0, 2), shard_views(3, 4, 1, 2)
    assert a["split_images"] == [] and b["split_images"] == []   # 3 images >= 2 ranks -> whole images
    a, b = shard_views(1, 5, 0, 2), shard_views(1, 5, 1, 2)
    assert (a["view_stop"], b["view_start"]) == (3, 3) and a["split_images"] == [0] and b["split_images"] == [0]
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _oracle_render_fn(S):
    orc = oracle_renderer(S)

    def fn(depth, albedo, view, light, vpi):
        outs_im, outs_d = [], []
        for i in range(depth.shape[0]):
            o = orc.render_chain(depth[i:i + 1], albedo[i:i + 1], view[i * vpi:(i + 1) * vpi], light[i * vpi:(i + 1) * vpi])
            outs_im.append(o["recon_im"])
            outs_d.append(o["recon_depth"])
        return torch.cat(outs_im, 0), torch.cat(outs_d, 0)
    return fn


def _worker(rank, world, port, n_images, vpi, S, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from g2s_b200 import synthetic
    from g2s_b200.sharding import render_chain_sharded
    case = synthetic.make_case(S, vpi, seed=3, n_images=n_images)
    out = render_chain_sharded(_oracle_render_fn(S), case["depth"], case["albedo"], case["view"], case["light"],
                               case["cotangent"], vpi, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, (out["shard"]["view_start"], out["recon_im"], out["grad_view"]))
    if rank == 0:
        q.put((out["grad_depth"], out["grad_albedo"], float(out["loss"]), gathered, out["shard"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images,vpi", [(1, 5), (2, 3)])
def test_world_size_2_matches_single_process(n_images, vpi):
    S, world = 16, 2
    from g2s_b200 import synthetic
    from g2s_b200.sharding import render_chain_sharded
    case = synthetic.make_case(S, vpi, seed=3, n_images=n_images)
    single = render_chain_sharded(_oracle_render_fn(S), case["depth"], case["albedo"], case["view"], case["light"],
                                  case["cotangent"], vpi, 0, 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300) + n_images
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, vpi, S, q)) for r in range(world)]
    for p in procs:
        p.start()
    g_depth, g_albedo, loss, gathered, shard = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # per-image gradients: complete (reduced) for the images rank 0 renders; images owned entirely by another rank
    # stay with that rank (no exchange when whole images are sharded)
    sl = slice(shard["image_start"], shard["image_stop"])
    assert rel_err(g_depth[sl], single["grad_depth"][sl]) < 1e-5
    assert rel_err(g_albedo[sl], single["grad_albedo"][sl]) < 1e-5
    if not shard["split_images"]:
        others = [i for i in range(n_images) if not (shard["image_start"] <= i < shard["image_stop"])]
        assert all(float(g_depth[i].abs().max()) == 0.0 for i in others)
    assert abs(loss - float(single["loss"])) <= 1e-5 * max(1.0, abs(float(single["loss"])))
    gathered.sort(key=lambda t: t[0])
    im = torch.cat([g[1] for g in gathered if g[1] is not None], 0)
    gv = torch.cat([g[2] for g in gathered if g[2] is not None], 0)
    assert torch.equal(im, single["recon_im"])          # forward outputs are bit-identical for 1 vs N ranks
    assert rel_err(gv, single["grad_view"]) < 1e-5
