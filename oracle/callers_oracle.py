"""oracle/callers_oracle.py -- TEST INFRASTRUCTURE: torch-CPU restatement of the callers either side of the renderer
path.  Each function cites the reference lines it follows.  Pinned against the reference's
own GAN2Shape/model.py and GAN2Shape/losses.py (imported unmodified through oracle/ref_model_shim.py) by
tests/test_callers_oracle.py, and against the committed golden vectors tests/golden/callers_*.npz those files produced.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import torch
import torch.nn.functional as F


def rescale_depth(depth, min_depth, max_depth):
    """GAN2Shape/model.py:85-86."""
    return (1 + depth) / 2 * max_depth + (1 - depth) / 2 * min_depth


def get_clamped_depth(depth_raw, h, w, min_depth, max_depth, border_depth=None, clamp_border=True, per_image=False):
    """GAN2Shape/model.py:337-345.  per_image=False is the reference (one mean over the whole tensor); per_image=True
    is the batched form (each map centred on its own mean)."""
    if border_depth is None:
        border_depth = 0.7 * max_depth + 0.3 * min_depth          # model.py:51
    if per_image:
        mean = depth_raw.reshape(depth_raw.shape[0], -1).mean(1).view(-1, 1, 1)
    else:
        mean = depth_raw.view(1, -1).mean(1).view(1, 1, 1)
    depth = rescale_depth(torch.tanh(depth_raw - mean), min_depth, max_depth)
    if clamp_border:
        depth_border = torch.zeros(1, h, w - 4)
        depth_border = F.pad(depth_border, (2, 2), mode='constant', value=1.02)
        depth = depth * (1 - depth_border) + depth_border * border_depth
    return depth


def recon_im_mask(recon_depth, min_depth, max_depth, masks=None):
    """GAN2Shape/model.py:146-150 (and :265-269 with the pseudo-view masks)."""
    margin = (max_depth - min_depth) / 2
    m = (recon_depth < max_depth + margin).float().unsqueeze(1).detach()
    return m if masks is None else m * masks


def photometric_loss(image1, image2, mask=None, conf_sigma=None):
    """GAN2Shape/losses.py:39-51."""
    loss = (image1 - image2).abs()
    if conf_sigma is not None:
        eps = 1e-7                                                       # losses.py:40
        loss = loss * 2 ** 0.5 / (conf_sigma + eps) + (conf_sigma + eps).log()
    if mask is not None:
        mask = mask.expand_as(loss)
        return (loss * mask).sum() / mask.sum()
    return loss.mean()


def _gradient(pred):
    """GAN2Shape/losses.py:74-79."""
    if pred.dim() == 4:
        pred = pred.reshape(-1, pred.size(2), pred.size(3))
    return pred[:, :, 1:] - pred[:, :, :-1], pred[:, 1:] - pred[:, :-1]


def smooth_loss(pred_map):
    """GAN2Shape/losses.py:56-72."""
    if type(pred_map) not in [tuple, list]:
        pred_map = [pred_map]
    loss, weight = 0, 1
    for m in pred_map:
        dx, dy = _gradient(m)
        dx2, dxdy = _gradient(dx)
        dydx, dy2 = _gradient(dy)
        loss = loss + (dx2.abs().mean() + dxdy.abs().mean() + dydx.abs().mean() + dy2.abs().mean()) * weight
        weight /= 2.3
    return loss
