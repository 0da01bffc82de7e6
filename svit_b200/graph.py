"""CUDA-graph replay of the hot path.

GraphedForward   the ~190 kernel launches of one SViT inference forward (plus the host-side ctypes / tensor-map work
                 of the C-ABI calls) captured once on a fixed input buffer and replayed as a single graph launch.
GraphedTrainStep one whole training step -- forward, loss, backward, bucketed gradient all-reduce (NCCL), gradient
                 clipping and the fused AdamW update -- captured once and replayed: ~510 C-ABI launches, ~700 torch glue
                 ops and 405 autograd hooks per step become one graph launch, so the step is bound by the GPU and not
                 by the host (tools/train_net.py:97-151 is the loop being replaced).

Every C-ABI entry point takes the caller's stream, allocates nothing and never synchronises, so both are capturable;
PyTorch supplies the capture-time allocator pool, the autograd tape (walked once, at capture) and the graph-safe RNG
(DropPath / head dropout draw fresh masks at every replay)."""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import ops


def _param_versions(model):
    return [p._version for p in model.parameters()]


class GraphedForward:
    """g = GraphedForward(model, example_clip); probs, extra = g(clip)   (inference only, fixed input shape).

    The returned tensors are the graph's static outputs: they are overwritten by the next call.  The graph reads the
    compute-dtype weight copies that were current at capture time; if a parameter is modified afterwards
    (load_state_dict, an optimizer step, .to()) the next call re-captures instead of replaying stale weights."""

    def __init__(self, model, example_clip: torch.Tensor, warmup: int = 2, lanes: int = 1):
        """`lanes` > 1 splits the batch into that many independent sub-batches whose forwards are captured on
        separate streams (clips are independent: no forward communication): inside the graph the kernels of one lane
        fill the SMs that the tail waves of another lane's kernels leave idle."""
        assert example_clip.is_cuda, "GraphedForward needs a CUDA clip"
        if model.training:
            raise RuntimeError("GraphedForward captures the inference forward: call model.eval() first")
        self.model = model
        self.static_in = example_clip.clone()
        self.warmup = max(1, warmup)
        self.lanes = max(1, min(int(lanes), self.static_in.shape[0]))
        self._lane_streams = [torch.cuda.Stream(device=self.static_in.device) for _ in range(self.lanes - 1)]
        self._capture()

    def _forward(self):
        if self.lanes == 1:
            return self.model([self.static_in])
        B = self.static_in.shape[0]
        bounds = [B * i // self.lanes for i in range(self.lanes + 1)]
        cur = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(cur)
        outs, done = [], []
        for i in range(self.lanes):
            part = self.static_in[bounds[i]:bounds[i + 1]]
            if i == 0:
                outs.append(self.model([part]))
                continue
            st = self._lane_streams[i - 1]
            st.wait_event(start)
            with torch.cuda.stream(st):
                outs.append(self.model([part]))
                ev = torch.cuda.Event()
                ev.record(st)
                done.append(ev)
        for ev in done:
            cur.wait_event(ev)
        probs = torch.cat([o[0] for o in outs])
        extra = {}
        for k, v in outs[0][1].items():
            same = torch.is_tensor(v) and v.ndim > 0 and v.shape[0] == bounds[1] - bounds[0]
            extra[k] = torch.cat([o[1][k] for o in outs]) if same else v
        return probs, extra

    def _capture(self):
        side = torch.cuda.Stream(device=self.static_in.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):  # first-call work (kernel attributes, table caches) must not be captured
                self._forward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.launches()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out, self.extra = self._forward()
        self.launches_per_replay = ops.launches() - n0
        # the graph holds raw pointers into the cached weight copies: keep them alive and remember their versions
        self._weights = [ent[3] for ent in ops._wcache.values()]
        self._versions = _param_versions(self.model)

    def __call__(self, clip: torch.Tensor):
        if self.model.training:
            raise RuntimeError("GraphedForward: the model was switched to train(); the graph is an inference forward")
        if _param_versions(self.model) != self._versions:
            self._capture()
        if clip.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(clip, non_blocking=True)
        self.graph.replay()
        ops._state["launches"] += self.launches_per_replay
        return self.out, self.extra


def cross_entropy_loss(preds, extra, labels):
    """The video-rank loss of the reference (models/losses.py:156-168 with is_video: CE on the class logits)."""
    return torch.nn.functional.cross_entropy(extra["logits"].float(), labels)


class GraphedTrainStep:
    """step = GraphedTrainStep(model, optimizer, clip, labels, reducer=None, max_norm=1.0); loss = step(clip, labels)

    One CUDA graph holds forward + loss + backward + gradient all-reduce + clip + AdamW.  `optimizer` is a
    svit_b200.optim.FusedAdamW (its step-dependent scalars are refreshed in device memory before every replay, so the
    learning-rate schedule keeps working: set param_groups[i]["lr"] between calls as usual); `reducer` a
    svit_b200.distributed.GradAllReducer or None (single GPU).  The returned loss is the graph's static output
    (device scalar, overwritten by the next call)."""

    def __init__(self, model, optimizer, example_clip: torch.Tensor, example_labels: torch.Tensor, reducer=None,
                 max_norm: Optional[float] = 1.0, loss_fn: Callable = cross_entropy_loss, warmup: int = 3,
                 frames_pass: bool = False):
        assert example_clip.is_cuda, "GraphedTrainStep needs CUDA tensors"
        if not model.training:
            raise RuntimeError("GraphedTrainStep captures a training step: call model.train() first")
        self.model, self.optimizer, self.reducer = model, optimizer, reducer
        self.max_norm, self.loss_fn, self.frames_pass = max_norm, loss_fn, frames_pass
        self.static_clip = example_clip.clone()
        self.static_labels = example_labels.clone()
        self.params = [p for p in model.parameters() if p.requires_grad]
        side = torch.cuda.Stream(device=example_clip.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):  # eager steps: kernel attributes, caches, NCCL communicators, optimizer state
                if optimizer is not None:
                    optimizer.upload_hyper()
                self._step_body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # compute-dtype weight copies must be re-made INSIDE the graph (the parameters change at every replay)
        ops._wcache.clear()
        if optimizer is not None:
            optimizer.upload_hyper()
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.launches()
        with torch.cuda.graph(self.graph):
            self.loss = self._step_body()
        self.launches_per_replay = ops.launches() - n0
        self._first = True  # the capture consumed one upload_hyper(): the first replay reuses it

    def _step_body(self):
        for p in self.params:
            p.grad = None
        if self.reducer is not None:
            self.reducer.prepare()
        preds, extra = self.model([self.static_clip])
        loss = self.loss_fn(preds, extra, self.static_labels)
        if self.frames_pass:
            from .distributed import consistency_loss, forward_video_frames
            _p, _e = forward_video_frames(self.model, self.static_clip)
            lam = getattr(self.model, "_lambda", {})
            for k, v in consistency_loss(lam, extra, _e).items():
                loss = loss + lam[k] * v
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        if self.optimizer is not None:
            self.optimizer.step(max_norm=self.max_norm, captured=True)
        return loss.detach()

    def __call__(self, clip: torch.Tensor, labels: torch.Tensor):
        if clip.data_ptr() != self.static_clip.data_ptr():
            self.static_clip.copy_(clip, non_blocking=True)
        if labels.data_ptr() != self.static_labels.data_ptr():
            self.static_labels.copy_(labels, non_blocking=True)
        if self.optimizer is not None and not self._first:
            self.optimizer.upload_hyper()
        self._first = False
        self.graph.replay()
        ops._state["launches"] += self.launches_per_replay
        if self.optimizer is not None:
            torch._C._increment_version(self.params)  # the replayed kernels wrote the parameters through raw pointers
        return self.loss
