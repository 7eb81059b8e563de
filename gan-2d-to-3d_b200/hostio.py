"""Host-buffer front end of the fused render step: inputs in pinned host memory, results back in pinned host memory.

`HostRenderStep` keeps two device-side slots and three streams (H2D, compute, D2H): the upload of step k+1 and the
download of step k-1 overlap the kernels of step k, so a stream of host-resident batches runs at the speed of the
kernels instead of copies + kernels.  Every step still copies its own inputs host->device and its own results
device->host; nothing is cached between steps.  This is what bench.py times as `e2e`.

    step = HostRenderStep(renderer, n_images, views_per_image)
    for batch in batches:                       # dicts of pinned CPU tensors: depth, albedo, view, light, cotangent (optional)
        done = step.submit(batch)               # enqueues this step, then returns the PREVIOUS step's results (or None);
                                                # they stay valid until the next submit()
    for res in step.drain(): ...                # the last results
"""
import torch


class _Slot:
    def __init__(self, N, B, S, dev, outputs):
        f = dict(device=dev, dtype=torch.float32)
        self.depth = torch.empty(N, S, S, **f)
        self.albedo = torch.empty(N, 3, S, S, **f)
        self.view = torch.empty(B, 6, **f)
        self.light = torch.empty(B, 4, **f)
        h = dict(dtype=torch.float32, pin_memory=True)
        self.host = {"grad_depth": torch.empty(N, S, S, **h), "grad_albedo": torch.empty(N, 3, S, S, **h),
                     "grad_view": torch.empty(B, 6, **h), "grad_light": torch.empty(B, 4, **h)}
        if "recon_im" in outputs:
            self.host["recon_im"] = torch.empty(B, 3, S, S, **h)
        if "recon_depth" in outputs:
            self.host["recon_depth"] = torch.empty(B, S, S, **h)
        self.uploaded = torch.cuda.Event()
        self.computed = torch.cuda.Event()
        self.downloaded = torch.cuda.Event()
        self.busy = False


class HostRenderStep:
    """Fused projected-view render fwd+bwd (Renderer.render_chain) fed from / drained to pinned host memory.

    `cotangent` [B,3,S,S] = d(loss)/d(recon_im) stays on the device (it stands for the losses the caller computes there);
    `outputs` names extra forward results to bring back besides the four gradients ("recon_im", "recon_depth")."""

    def __init__(self, renderer, n_images, views_per_image, cotangent, outputs=()):
        self.ren, self.N, self.P = renderer, n_images, views_per_image
        self.B, self.S = n_images * views_per_image, renderer.image_size
        self.dev = torch.device(renderer.device)
        self.cot = cotangent
        self.outputs = tuple(outputs)
        self.slots = [_Slot(self.N, self.B, self.S, self.dev, self.outputs) for _ in range(2)]
        self.h2d, self.comp, self.d2h = (torch.cuda.Stream(device=self.dev) for _ in range(3))
        self.k = 0
        self.h2d_bytes = 4 * (self.N * self.S * self.S * 4 + self.B * 10)
        self.d2h_bytes = sum(t.numel() * 4 for t in self.slots[0].host.values())

    def _finish(self, slot):
        slot.downloaded.synchronize()
        slot.busy = False
        return slot.host

    def submit(self, batch):
        slot, prev = self.slots[self.k & 1], self.slots[(self.k + 1) & 1]
        self.k += 1
        assert not slot.busy     # its previous step (two submits ago) was waited for by the last submit
        with torch.cuda.stream(self.h2d):
            slot.depth.copy_(batch["depth"], non_blocking=True)
            slot.albedo.copy_(batch["albedo"], non_blocking=True)
            slot.view.copy_(batch["view"], non_blocking=True)
            slot.light.copy_(batch["light"], non_blocking=True)
            slot.uploaded.record(self.h2d)
        with torch.cuda.stream(self.comp):
            self.comp.wait_event(slot.uploaded)
            d, a = slot.depth.requires_grad_(True), slot.albedo.requires_grad_(True)
            v, l = slot.view.requires_grad_(True), slot.light.requires_grad_(True)
            im, rd, _ = self.ren.render_chain(d, a, v, l, views_per_image=self.P)
            grads = torch.autograd.grad([im], [d, a, v, l], grad_outputs=[self.cot])
            slot.computed.record(self.comp)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(slot.computed)
            for name, g in zip(("grad_depth", "grad_albedo", "grad_view", "grad_light"), grads):
                slot.host[name].copy_(g, non_blocking=True)
                g.record_stream(self.d2h)
            if "recon_im" in self.outputs:
                slot.host["recon_im"].copy_(im.detach(), non_blocking=True)
                im.record_stream(self.d2h)
            if "recon_depth" in self.outputs:
                slot.host["recon_depth"].copy_(rd.detach(), non_blocking=True)
                rd.record_stream(self.d2h)
            slot.downloaded.record(self.d2h)
        for t in (slot.depth, slot.albedo, slot.view, slot.light):
            t.requires_grad_(False)
        slot.busy = True
        # with this step queued behind it, wait for the previous one: the GPU stays busy while the host has its results
        return self._finish(prev) if prev.busy else None

    def drain(self):
        """Waits for the steps still in flight and returns their host results, oldest first."""
        return [self._finish(slot) for slot in (self.slots[self.k & 1], self.slots[(self.k + 1) & 1]) if slot.busy]
