"""CUDA-graph capture of the fused render step for small, launch-bound batches.

One image x 16 views (the literal GAN2Shape cat configuration) is ~0.25 ms of kernels but ~0.55 ms end to end in
eager mode: ~13 kernel launches, memsets and the Python/autograd glue dominate.  `GraphedRenderStep` captures
forward + backward (torch.autograd.grad, so no AccumulateGrad nodes tied to another stream) once into a
torch.cuda.CUDAGraph over static buffers and replays it: the step becomes a single graph launch.
"""
import torch


class GraphedRenderStep:
    """Static-shape fwd+bwd of Renderer.render_chain with a fixed cotangent source.

    step(depth, albedo, view, light, cotangent) copies the inputs into the static buffers, replays the graph and
    returns (recon_im, recon_depth, face_idx, (g_depth, g_albedo, g_view, g_light)) -- views of static outputs that the
    next step overwrites."""

    def __init__(self, renderer, n_images, views_per_image, device=None, warmup=3):
        S = renderer.image_size
        dev = torch.device(device if device is not None else renderer.device)
        B = n_images * views_per_image
        self.renderer, self.P = renderer, views_per_image
        self.depth = torch.full((n_images, S, S), 1.0, device=dev, requires_grad=True)
        self.albedo = torch.zeros(n_images, 3, S, S, device=dev, requires_grad=True)
        self.view = torch.zeros(B, 6, device=dev, requires_grad=True)
        self.light = torch.zeros(B, 4, device=dev, requires_grad=True)
        self.cotangent = torch.zeros(B, 3, S, S, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):           # allocations, lazy inits (z-buffer, smem opt-in) outside the capture
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._run()

    def _run(self):
        im, rd, fidx = self.renderer.render_chain(self.depth, self.albedo, self.view, self.light,
                                                  views_per_image=self.P)
        # the loss stands on the cotangent (d loss / d recon_im), handed straight to the backward
        grads = torch.autograd.grad([im], [self.depth, self.albedo, self.view, self.light], grad_outputs=[self.cotangent])
        return im, rd, fidx, grads

    def step(self, depth=None, albedo=None, view=None, light=None, cotangent=None):
        with torch.no_grad():
            for dst, src in ((self.depth, depth), (self.albedo, albedo), (self.view, view), (self.light, light),
                             (self.cotangent, cotangent)):
                if src is not None:
                    dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.out
