"""Data-parallel plumbing for the SViT hot path: one process per GPU, clips sharded across ranks, no
communication in the forward pass, one bucketed gradient all-reduce per training step.

Replaces the reference's DistributedDataParallel wrap (slowfast/models/build.py:69-74) for this path.
The collective itself is NCCL over NVLink 5 / NVSwitch through torch.distributed (34.37 M fp32 gradients =
137.5 MB per step): buckets are filled in reverse parameter order -- head and late blocks
first, the order in which backward produces them -- and each bucket's all-reduce is launched
asynchronously as soon as its last gradient has been accumulated, so the transfers hide under the
remaining backward work (blocks 1 and 0 are both the heaviest and the last to finish).

On CPU the same class runs over gloo (tests/test_dist_gloo.py, world_size 2).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun). Returns (rank, world, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available() and backend != "gloo"
    dev = torch.device("cuda", local) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if use_cuda:
            dist.init_process_group(backend or "nccl", device_id=dev)
        else:
            dist.init_process_group(backend or "gloo")
    return rank, world, dev


def shard_batch(n_items: int, rank: int, world: int) -> range:
    """Contiguous, even split of the clip batch (the reference requires divisibility: defaults.py:1141-1149)."""
    if n_items % world != 0:
        raise ValueError(f"batch of {n_items} clips is not divisible by world size {world}")
    per = n_items // world
    return range(rank * per, (rank + 1) * per)


class GradAllReducer:
    """Bucketed, overlapped gradient averaging (sum all-reduce / world size, as DDP does).

    Usage per step:   reducer.prepare(); loss.backward(); reducer.finish()
    Parameters that received no gradient are reduced as zeros, which mirrors find_unused_parameters=False
    plus the reference's `+ sum(p) * 0` tricks (video_model_builder.py:359, 514)."""

    def __init__(self, params, bucket_bytes: int = 32 << 20, group=None, overlap: bool = True):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.overlap = overlap  # False: one pass after backward (no NCCL CTAs next to the persistent GEMM kernels)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        order = list(reversed(self.params))  # backward completion order ~ reverse registration order
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, size = [], 0
        for p in order:
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat = [torch.zeros(sum(p.numel() for p in b), dtype=torch.float32, device=b[0].device)
                     for b in self.buckets]
        self._where = {}
        for bi, b in enumerate(self.buckets):
            off = 0
            for p in b:
                self._where[p] = (bi, off)
                off += p.numel()
        self._views = []
        for bi, b in enumerate(self.buckets):
            off, vs = 0, []
            for p in b:
                vs.append(self.flat[bi][off:off + p.numel()].view_as(p))
                off += p.numel()
            self._views.append(vs)
        # NCCL averages inside the collective; gloo has no AVG: sum, then divide
        nccl = dist.is_initialized() and dist.get_backend(group) == "nccl"
        self._op = dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM
        self._pending = [0] * len(self.buckets)
        self._works = [None] * len(self.buckets)
        self._launched = [0] * len(self.buckets)
        self._next = 0
        self._prepared = False
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.bytes_per_step = sum(f.numel() * 4 for f in self.flat)

    def prepare(self):
        for i, b in enumerate(self.buckets):
            self._pending[i] = len(b)
            self._works[i] = None
            self._launched[i] = 0
        self._next = 0
        self._prepared = True

    def _launch(self, bi: int):
        """Pack the bucket with ONE multi-tensor copy (torch._foreach_copy_), then start its all-reduce."""
        b, views = self.buckets[bi], self._views[bi]
        missing = [v for v, p in zip(views, b) if p.grad is None]
        if missing:
            torch._foreach_zero_(missing)  # parameters without a gradient are reduced as zeros
        # gradients accumulated in place into last step's bucket views are already where they belong
        have = [(v, p.grad) for v, p in zip(views, b) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        if have:
            torch._foreach_copy_([v for v, _ in have], [g.view_as(v) for v, g in have])
        self._launched[bi] += 1
        if self.world > 1:
            self._works[bi] = dist.all_reduce(self.flat[bi], op=self._op, group=self.group, async_op=True)

    def _drain_ready(self):
        """Launch, IN BUCKET ORDER, every bucket whose gradients are complete.  A complete bucket behind an incomplete
        one waits (as DDP does): ranks whose sets of gradient-less parameters differ (video ranks never touch the box /
        contact heads, image ranks never touch pos_embed_temporal or head.projection) would otherwise issue the same
        collectives in different orders -- mismatched sizes, a hang or a wrong reduction."""
        while self._next < len(self.buckets) and self._pending[self._next] == 0:
            self._launch(self._next)
            self._next += 1

    def _on_grad(self, p: torch.nn.Parameter):
        if not self._prepared:
            return  # backward outside prepare()/finish(): nothing is reduced, finish() will say so
        bi = self._where[p][0]
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and self.overlap:
            self._drain_ready()

    def finish(self):
        """Launch what is left (buckets with gradient-less parameters, or everything when overlap is off) in bucket
        order, wait, average; .grad becomes a view of the bucket (no copy back: 405 attribute assignments instead of
        405 kernels)."""
        if not self._prepared:
            raise RuntimeError("GradAllReducer.finish() without prepare(): call prepare() before backward()")
        while self._next < len(self.buckets):
            self._pending[self._next] = 0
            self._launch(self._next)
            self._next += 1
        if any(n != 1 for n in self._launched):
            raise RuntimeError(f"GradAllReducer: every bucket must be reduced exactly once per step, got {self._launched}")
        for bi, b in enumerate(self.buckets):
            if self._works[bi] is not None:
                self._works[bi].wait()
            if self.world > 1 and self._op == dist.ReduceOp.SUM:
                self.flat[bi].div_(self.world)
            for p, v in zip(b, self._views[bi]):
                p.grad = v
        self._prepared = False

    def remove(self):
        for h in self._hooks:
            h.remove()


def forward_video_frames(model, clip):
    """The reference's second, gradient-free pass of every training step (tools/train_net.py:105-110,
    TRAIN.FORWARD_VIDEO_FRAMES, on by default): the B clips are re-fed as B*T single frames
    ([B, C, T, H, W] -> [B*T, C, 1, H, W]), i.e. the T = 1 shapes of every kernel (N = 3141 / 789 / 201 / 54 tokens,
    rel_pos_t interpolated 15 -> 1).  Returns (preds, extra_preds) of the frames."""
    with torch.no_grad():
        frames = clip.transpose(1, 2).flatten(0, 1).unsqueeze(2)
        return model([frames])


def consistency_loss(lambdas: dict, extra_preds: dict, frames_extra_preds: dict) -> dict:
    """VideoImageLoss._consistency_loss (models/losses.py:127-136), verbatim semantics: the terms are keyed
    'video_image_desc_l1_loss' / '..._l2_loss', while get_lambdas_dict (utils/misc.py:411-423) only ever defines
    'video_image_boxes_l1_loss' -- so with the stock config the dict is empty and the frames pass adds compute but
    no loss term.  Kept as is: a drop-in must not change the optimisation problem."""
    ret = {}
    pred = extra_preds["obj_desc"]                                   # [B, T, O, d]
    tar = frames_extra_preds["obj_desc"].reshape(pred.shape).detach()  # [B*T, 1, O, d] -> [B, T, O, d]
    if "video_image_desc_l1_loss" in lambdas:
        ret["video_image_desc_l1_loss"] = torch.nn.functional.l1_loss(pred, tar)
    if "video_image_desc_l2_loss" in lambdas:
        ret["video_image_desc_l2_loss"] = torch.nn.functional.mse_loss(pred, tar)
    return ret


def train_step(model, reducer: Optional[GradAllReducer], clip, labels, optimizer=None, max_norm: Optional[float] = 1.0,
               frames_pass: bool = False, lambdas: Optional[dict] = None):
    """One data-parallel training step on this rank's shard: forward, optional frames pass (train_net.py:105-110),
    CE loss (+ the consistency terms the lambda dict enables; losses.py:156-168, video rank), backward with
    overlapped gradient all-reduce, optional clip (tools/train_net.py:144-147) and optimizer step."""
    if reducer is not None:
        reducer.prepare()
    logits, extra = model([clip])
    loss = torch.nn.functional.cross_entropy(logits.float(), labels)
    if frames_pass and clip.size(2) > 1:
        _preds, _extra = forward_video_frames(model, clip)
        extra["frames_output"] = {"preds": _preds, "extra_preds": _extra}
        lam = lambdas if lambdas is not None else getattr(model, "_lambda", {})
        for k, v in consistency_loss(lam, extra, _extra).items():
            loss = loss + lam[k] * v
    loss.backward()
    if reducer is not None:
        reducer.finish()
    if max_norm is not None:
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
    if optimizer is not None:
        optimizer.step()
        optimizer.zero_grad(set_to_none=False)
    return loss.detach()
