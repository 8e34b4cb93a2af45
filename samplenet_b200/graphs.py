"""CUDA-graph capture of a whole SampleNet step.

At the headline size (B=32, N=1024->64) the step is ~a dozen kernels of a few microseconds each; launched one by one from
Python the host is the bottleneck.  `GraphedStep` captures `net(x)` + `net.get_simplification_loss(...)` once into a CUDA
graph over static buffers and replays it: one host call per step.  This is the B200-native replacement for what a tracing
compiler would do, and it is part of the public API:

    step = GraphedStep(net, batch_size=32, num_points=1024)            # forward + loss (no grad)
    simp, proj, loss = step(x_cuda)                                    # x already in HBM
    loss_host = step.run_from_host(x_pinned)                           # H2D copy + replay + D2H of the loss, synchronised
"""
import torch

from . import _lib


class GraphedStep:
    def __init__(self, net, batch_size, num_points, gamma=1, delta=0, device=None, warmup=2):
        self.net = net
        dev = torch.device(device) if device is not None else next(net.parameters()).device
        self.device = dev
        shape = (batch_size, num_points, 3) if net.input_shape == "bnc" else (batch_size, 3, num_points)
        self.x = torch.zeros(shape, device=dev)
        self.loss_host = torch.zeros(1).pin_memory()
        self.stream = torch.cuda.Stream(device=dev)
        m = net.num_out_points

        def body():
            simp, proj = net(self.x)
            ref_bnc = self.x if net.input_shape == "bnc" else self.x.permute(0, 2, 1).contiguous()
            simp_bnc = simp if net.output_shape == "bnc" else simp.permute(0, 2, 1).contiguous()
            loss = net.get_simplification_loss(ref_bnc, simp_bnc, m, gamma, delta)
            return simp, proj, loss

        with torch.cuda.device(dev), torch.no_grad():
            self.stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self.stream):
                for _ in range(warmup):
                    body()
            self.stream.synchronize()
            before = _lib.launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.simp, self.proj, self.loss = body()
            self.launches_per_step = _lib.launch_count() - before
        self.loss_flat = self.loss.reshape(1)

    def __call__(self, x):
        """x: CUDA tensor shaped like the capture buffer (copied device-to-device), returns the static outputs."""
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.simp, self.proj, self.loss

    def run_from_host(self, x_pinned):
        """End-to-end step: pinned host batch -> device, replay, loss back to the host (synchronised); returns float."""
        self.x.copy_(x_pinned, non_blocking=True)
        self.graph.replay()
        self.loss_host.copy_(self.loss_flat, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return float(self.loss_host[0])
