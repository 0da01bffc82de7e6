"""Fused optimizer step for the SViT training loop.

Reference: `construct_optimizer` (slowfast/models/optimizer.py:15-112; configs/ssv2.yaml: AdamW, lr 2e-4, weight decay
1e-4, ZERO_WD_1D_PARAM) and the step in tools/train_net.py:133-151 (`clip_grad_norm_(params, 1.0)` then
`optimizer.step()`).  torch steps the 405 parameter tensors through foreach kernels after a separate norm pass; here a
device-side pointer table drives two launches of libsvit_sm100.so (`svit_grad_sqnorm`, `svit_adamw_step`): the clip
coefficient is computed on the device from the squared norm, so the step has no host synchronisation.

`FusedAdamW` keeps torch.optim.AdamW's interface for what the loop uses (param_groups with per-group lr /
weight_decay, `step()`, `zero_grad()`); results match torch.optim.AdamW to fp32 rounding
(tests/test_gpu_parity.py::test_fused_adamw_matches_torch).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, List, Optional

import torch

from . import _lib
from .ops import _call, _stream

CHUNK = 8192  # elements per scheduling unit


def split_weight_decay_groups(model, weight_decay: float, zero_wd_1d: bool = True):
    """The grouping rule of optimizer.py:31-58 (no BatchNorm on this path): 1-D parameters and biases get zero weight
    decay when SOLVER.ZERO_WD_1D_PARAM, as do names returned by model.no_weight_decay()."""
    skip = model.no_weight_decay() if hasattr(model, "no_weight_decay") else {}
    decay, zero = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if name in skip or ((p.dim() == 1 or name.endswith(".bias")) and zero_wd_1d):
            zero.append(p)
        else:
            decay.append(p)
    groups = [{"params": decay, "weight_decay": weight_decay}, {"params": zero, "weight_decay": 0.0}]
    return [g for g in groups if g["params"]]


def construct_optimizer(model, cfg):
    """optimizer.py:15-112 for OPTIMIZING_METHOD == 'adamw' (what configs/ssv2.yaml selects)."""
    if cfg.SOLVER.OPTIMIZING_METHOD != "adamw":
        raise NotImplementedError(f"Does not support {cfg.SOLVER.OPTIMIZING_METHOD} optimizer")
    groups = split_weight_decay_groups(model, cfg.SOLVER.WEIGHT_DECAY, cfg.SOLVER.ZERO_WD_1D_PARAM)
    return FusedAdamW(groups, lr=cfg.SOLVER.BASE_LR, eps=1e-8, weight_decay=cfg.SOLVER.WEIGHT_DECAY)


def lr_func_cosine(cfg, cur_epoch: float, base_lr=None) -> float:
    """utils/lr_policy.py:35-66: half-cosine from BASE_LR to COSINE_END_LR over MAX_EPOCH (offset by the warm-up when
    COSINE_AFTER_WARMUP)."""
    if base_lr is None:
        base_lr, end_lr = cfg.SOLVER.BASE_LR, cfg.SOLVER.COSINE_END_LR
    elif isinstance(base_lr, tuple):
        base_lr, end_lr = base_lr
    else:
        end_lr = cfg.SOLVER.COSINE_END_LR
    offset = cfg.SOLVER.WARMUP_EPOCHS if cfg.SOLVER.COSINE_AFTER_WARMUP else 0.0
    assert end_lr < base_lr
    return end_lr + (base_lr - end_lr) * (math.cos(math.pi * (cur_epoch - offset) / (cfg.SOLVER.MAX_EPOCH - offset)) + 1.0) * 0.5


def get_epoch_lr(cur_epoch: float, cfg) -> dict:
    """models/optimizer.py:115-125 -> utils/lr_policy.py:9-32 for LR_POLICY 'cosine' (what configs/ssv2.yaml uses):
    the policy value, replaced by a linear ramp from WARMUP_START_LR during the first WARMUP_EPOCHS.  Returns
    {'lr': value} like the reference."""
    if cfg.SOLVER.LR_POLICY != "cosine":
        raise NotImplementedError(f"Unknown LR policy: {cfg.SOLVER.LR_POLICY}")
    lr = lr_func_cosine(cfg, cur_epoch, base_lr=cfg.SOLVER.BASE_LR)
    if cur_epoch < cfg.SOLVER.WARMUP_EPOCHS:
        lr_start = cfg.SOLVER.WARMUP_START_LR
        lr_end = lr_func_cosine(cfg, cfg.SOLVER.WARMUP_EPOCHS)
        lr = cur_epoch * (lr_end - lr_start) / cfg.SOLVER.WARMUP_EPOCHS + lr_start
    return {"lr": lr}


def set_lr(optimizer, new_lr: dict):
    """models/optimizer.py:128-137."""
    for g in optimizer.param_groups:
        g["lr"] = new_lr["lr"]


class FusedAdamW:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        params = list(params)
        if params and not isinstance(params[0], dict):
            params = [{"params": params}]
        self.param_groups: List[dict] = []
        for g in params:
            g = dict(g)
            g["params"] = [p for p in g["params"]]
            g.setdefault("lr", lr)
            g.setdefault("betas", betas)
            g.setdefault("eps", eps)
            g.setdefault("weight_decay", weight_decay)
            self.param_groups.append(g)
        self.state = {}
        self._step = 0
        self._tables = None
        self._sqnorm = None
        self._hyper_dev = None

    # ---- table construction -----------------------------------------------------------------------------------
    def _build(self):
        self._tables = []
        for g in self.param_groups:
            ps = [p for p in g["params"] if p.requires_grad]
            if not ps:
                self._tables.append(None)
                continue
            dev = ps[0].device
            for p in ps:
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdamW: parameters must be contiguous fp32 tensors")
                if p not in self.state:
                    self.state[p] = {"exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
            tab = torch.zeros(len(ps), 6, dtype=torch.int64)
            wd_bits = torch.tensor([g["weight_decay"]], dtype=torch.float32).view(torch.int32).item()
            chunk_tensor, chunk_start = [], []
            for i, p in enumerate(ps):
                st = self.state[p]
                tab[i, 0], tab[i, 2], tab[i, 3] = p.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                tab[i, 4] = p.numel()
                tab[i, 5] = wd_bits & 0xFFFFFFFF  # float in the low 4 bytes, padding above
                for s in range(0, p.numel(), CHUNK):
                    chunk_tensor.append(i)
                    chunk_start.append(s)
            self._tables.append({
                "params": ps, "host": tab.pin_memory() if torch.cuda.is_available() else tab,
                "dev": torch.empty(len(ps), 6, dtype=torch.int64, device=dev),
                "chunk_tensor": torch.tensor(chunk_tensor, dtype=torch.int32, device=dev),
                "chunk_start": torch.tensor(chunk_start, dtype=torch.int64, device=dev),
                "nchunks": len(chunk_tensor), "wd": g["weight_decay"], "grad_ptrs": None})
        dev = next(t for t in self._tables if t is not None)["dev"].device
        self._sqnorm = torch.zeros(len(self._tables) + 1, dtype=torch.float32, device=dev)

    def _refresh(self, t):
        """Gradient pointers change whenever autograd allocates fresh .grad tensors: re-upload the table when they did.

        A parameter without a gradient gets a persistent ZERO gradient: the reference keeps every parameter in the graph
        (`+ sum(p) * 0` terms, video_model_builder.py:359, 514), so torch.optim.AdamW still applies decoupled weight
        decay and moment decay to it every step -- with or without the gradient all-reducer in front."""
        ptrs = []
        for p in t["params"]:
            g = p.grad
            if g is None:
                g = p.grad = torch.zeros_like(p)
            if g.dtype != torch.float32 or not g.is_contiguous():
                raise RuntimeError("FusedAdamW: gradients must be contiguous fp32 tensors")
            ptrs.append(g.data_ptr())
        if ptrs != t["grad_ptrs"]:
            # the pinned table is the source of an asynchronous copy: the previous upload must have executed before the
            # host rewrites it (the step never synchronises otherwise, so the host may run a full step ahead)
            capturing = t["dev"].is_cuda and torch.cuda.is_current_stream_capturing()
            if t.get("uploaded") is not None and not capturing:
                t["uploaded"].synchronize()
            t["host"][:, 1] = torch.tensor(ptrs, dtype=torch.int64)
            t["dev"].copy_(t["host"], non_blocking=True)
            if t["dev"].is_cuda and not capturing:
                t["uploaded"] = torch.cuda.Event()
                t["uploaded"].record()
            t["grad_ptrs"] = ptrs

    # ---- CUDA-graph mode: step-dependent scalars live in device memory -------------------------------------------
    def _hyper_row(self, g) -> list:
        import numpy as np
        b1, b2 = g["betas"]
        st = np.float32(self._step)
        bc1 = np.float32(1.0) - np.power(np.float32(b1), st)
        sqrt_bc2 = np.sqrt(np.float32(1.0) - np.power(np.float32(b2), st))
        return [float(g["lr"]), float(bc1), float(sqrt_bc2), 0.0]

    def upload_hyper(self):
        """Graph mode, OUTSIDE the graph and before each replay: advance the step counter and refresh
        {lr, 1 - beta1^step, sqrt(1 - beta2^step)} per group in device memory (one small H2D copy from a ring of pinned
        rows, so the host never rewrites a row whose copy may still be pending)."""
        if self._tables is None:
            self._build()
        live = [g for g, t in zip(self.param_groups, self._tables) if t is not None]
        if self._hyper_dev is None:
            dev = self._sqnorm.device
            self._hyper_dev = torch.zeros(len(live), 4, dtype=torch.float32, device=dev)
            self._hyper_host = torch.zeros(8, len(live), 4, dtype=torch.float32).pin_memory()
            self._hyper_events = [None] * 8
        self._step += 1
        slot = self._step % 8
        if self._hyper_events[slot] is not None:
            self._hyper_events[slot].synchronize()
        self._hyper_host[slot] = torch.tensor([self._hyper_row(g) for g in live], dtype=torch.float32)
        self._hyper_dev.copy_(self._hyper_host[slot], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._hyper_events[slot] = ev

    # ---- public API -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, max_norm: Optional[float] = None, captured: bool = False):
        """One AdamW step over every group; `max_norm` fuses torch.nn.utils.clip_grad_norm_(all params, max_norm).
        captured=True (inside a CUDA graph capture / GraphedTrainStep): lr and the bias corrections are read from the
        device row written by upload_hyper(), and the step counter is not touched here."""
        if self._tables is None:
            self._build()
        if captured and self._hyper_dev is None:
            raise RuntimeError("FusedAdamW.step(captured=True) needs upload_hyper() first")
        if not captured:
            self._step += 1
        live = [(g, t) for g, t in zip(self.param_groups, self._tables) if t is not None]
        for _, t in live:
            self._refresh(t)
        total = None
        if max_norm is not None and max_norm > 0:
            for i, (_, t) in enumerate(live):
                _call("svit_grad_sqnorm", t["dev"].data_ptr(), t["chunk_tensor"].data_ptr(), t["chunk_start"].data_ptr(),
                      t["nchunks"], CHUNK, self._sqnorm[i:].data_ptr(), _stream())
            total = self._sqnorm[len(live):len(live) + 1]
            torch.sum(self._sqnorm[:len(live)], dim=0, keepdim=True, out=total)  # norm over all groups
        for i, (g, t) in enumerate(live):
            b1, b2 = g["betas"]
            mn = float(max_norm) if total is not None else 0.0
            tp = total.data_ptr() if total is not None else None
            if captured:
                _call("svit_adamw_step_dev", t["dev"].data_ptr(), t["chunk_tensor"].data_ptr(), t["chunk_start"].data_ptr(),
                      t["nchunks"], CHUNK, self._hyper_dev[i].data_ptr(), float(b1), float(b2), float(g["eps"]), mn, tp,
                      _stream())
            else:
                _call("svit_adamw_step", t["dev"].data_ptr(), t["chunk_tensor"].data_ptr(), t["chunk_start"].data_ptr(),
                      t["nchunks"], CHUNK, float(g["lr"]), float(b1), float(b2), float(g["eps"]), self._step, mn, tp,
                      _stream())
            # the kernel writes the parameters through raw pointers: tell autograd / the bf16 weight caches
            # (ops.cast_weight keys on param._version) that they changed
            torch._C._increment_version(t["params"])  # takes an iterable of tensors: one call, not one per tensor

    def grad_norm(self) -> torch.Tensor:
        """Total gradient norm seen by the last clipped step (device scalar)."""
        n = sum(1 for t in self._tables if t is not None)
        return self._sqnorm[n].sqrt()

    # ---- checkpointing: the torch.optim.AdamW layout, so optimizer states saved by the reference's
    # utils/checkpoint.py (`optimizer.state_dict()`) resume here and vice versa ----------------------------------------
    def state_dict(self) -> dict:
        index, groups = {}, []
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                index.setdefault(id(p), len(index))
                ids.append(index[id(p)])
            b1, b2 = g["betas"]
            groups.append({"lr": g["lr"], "betas": (b1, b2), "eps": g["eps"], "weight_decay": g["weight_decay"],
                           "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                           "differentiable": False, "fused": None, "params": ids})
        state = {}
        for g in self.param_groups:
            for p in g["params"]:
                st = self.state.get(p)
                if st is not None:
                    state[index[id(p)]] = {"step": torch.tensor(float(self._step)), "exp_avg": st["exp_avg"],
                                           "exp_avg_sq": st["exp_avg_sq"]}
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd: dict):
        groups = sd["param_groups"]
        if len(groups) != len(self.param_groups) or any(len(a["params"]) != len(b["params"])
                                                         for a, b in zip(groups, self.param_groups)):
            raise ValueError("loaded state dict has a different number of parameter groups / parameters")
        by_index = {}
        for saved, g in zip(groups, self.param_groups):
            for k in ("lr", "eps", "weight_decay"):
                g[k] = saved[k]
            g["betas"] = tuple(saved["betas"])
            for idx, p in zip(saved["params"], g["params"]):
                by_index[idx] = p
        steps = set()
        for idx, st in sd["state"].items():
            p = by_index[int(idx)]
            mine = self.state.setdefault(p, {"exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)})
            mine["exp_avg"].copy_(st["exp_avg"])     # in place: the device pointer table keeps pointing at these buffers
            mine["exp_avg_sq"].copy_(st["exp_avg_sq"])
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): the fused step keeps one counter")
        self._step = steps.pop() if steps else 0
        self._tables = None  # weight decay / membership may have changed: rebuild the tables at the next step

    def zero_grad(self, set_to_none: bool = True):
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is not None:
                    if set_to_none:
                        p.grad = None
                    else:
                        p.grad.zero_()
