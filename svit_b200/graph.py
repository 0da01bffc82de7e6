"""CUDA-graph replay of the inference forward: the ~190 kernel launches of one SViT forward (plus the host-side
ctypes / tensor-map work of the C-ABI calls) are captured once on a fixed input buffer and replayed as a single
graph launch.  Every C-ABI entry point takes the caller's stream, allocates nothing and never synchronises, so the
whole forward is capturable; PyTorch only supplies the capture-time allocator pool."""
from __future__ import annotations

import torch

from . import ops


class GraphedForward:
    """g = GraphedForward(model, example_clip); probs, extra = g(clip)   (inference only, fixed input shape).

    The returned tensors are the graph's static outputs: they are overwritten by the next call."""

    def __init__(self, model, example_clip: torch.Tensor, warmup: int = 2):
        assert example_clip.is_cuda, "GraphedForward needs a CUDA clip"
        self.model = model.eval()
        self.static_in = example_clip.clone()
        side = torch.cuda.Stream(device=example_clip.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):  # first-call work (kernel attributes, table caches) must not be captured
                self.model([self.static_in])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.launches()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out, self.extra = self.model([self.static_in])
        self.launches_per_replay = ops.launches() - n0

    def __call__(self, clip: torch.Tensor):
        if clip.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(clip, non_blocking=True)
        self.graph.replay()
        ops._state["launches"] += self.launches_per_replay
        return self.out, self.extra
