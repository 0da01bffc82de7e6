"""Configuration tree for the SViT hot path.

The reference reads an fvcore/yacs ``CfgNode`` (slowfast/config/defaults.py) merged with
configs/ssv2.yaml.  Only the keys the model constructor touches are needed for the hot
path, so this module provides a small attribute-dict with exactly those keys and the
values of the shipped SSv2 recipe:

  MVIT.*      configs/ssv2.yaml:58-164 (+ defaults.py:345-471 for keys the YAML omits)
  SVIT.*      configs/ssv2.yaml:185-188, defaults.py:20-28
  DATA.*      configs/ssv2.yaml (NUM_FRAMES 16, crop 224), defaults.py
  DETECTION.* defaults.py:720-732
  MODEL.*     configs/ssv2.yaml:53-57, defaults.py:334

A YAML file in the reference's format can be merged on top with ``merge_yaml``; tuple
valued keys arrive as strings there ("(3, 7, 7)") and are literal_eval'ed.
"""
from __future__ import annotations

import ast
import copy


class CfgNode(dict):
    """dict with attribute access, nested."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        return CfgNode({k: copy.deepcopy(v, memo) for k, v in self.items()})


_SSV2 = {
    "NUM_GPUS": 8,
    "DATA": {
        "NUM_FRAMES": 16,
        "TRAIN_CROP_SIZE": 224,
        "TEST_CROP_SIZE": 224,
        "INPUT_CHANNEL_NUM": [3],
        "MEAN": [0.45, 0.45, 0.45],   # config/defaults.py DATA.MEAN / DATA.STD (configs/ssv2.yaml does not override)
        "STD": [0.225, 0.225, 0.225],
    },
    "MODEL": {
        "NUM_CLASSES": 174,
        "DROPOUT_RATE": 0.5,
        "HEAD_ACT": "softmax",
        "MODEL_NAME": "SViT",
        "ROI_HEAD_ACT_DURING_TRAINING": False,
    },
    "MVIT": {
        "MODE": "conv",
        "POOL_FIRST": False,
        "CLS_EMBED_ON": True,
        "PATCH_KERNEL": [3, 7, 7],
        "PATCH_STRIDE": [2, 4, 4],
        "PATCH_PADDING": [1, 3, 3],
        "PATCH_2D": False,
        "EMBED_DIM": 96,
        "NUM_HEADS": 1,
        "MLP_RATIO": 4.0,
        "QKV_BIAS": True,
        "DROPPATH_RATE": 0.4,
        "DROPOUT_RATE": 0.0,
        "DEPTH": 16,
        "NORM": "layernorm",
        "DIM_MUL": [[1, 2.0], [3, 2.0], [14, 2.0]],
        "HEAD_MUL": [[1, 2.0], [3, 2.0], [14, 2.0]],
        "POOL_KVQ_KERNEL": [3, 3, 3],
        "POOL_KV_STRIDE": None,
        "POOL_KV_STRIDE_ADAPTIVE": [1, 8, 8],
        "POOL_Q_STRIDE": [[0, 1, 1, 1], [1, 1, 2, 2], [2, 1, 1, 1], [3, 1, 2, 2]]
        + [[i, 1, 1, 1] for i in range(4, 14)]
        + [[14, 1, 2, 2], [15, 1, 1, 1]],
        "NORM_STEM": False,
        "SEP_POS_EMBED": False,
        "USE_ABS_POS": False,
        "REL_POS_SPATIAL": True,
        "REL_POS_TEMPORAL": True,
        "REL_POS_ZERO_INIT": False,
        "RESIDUAL_POOLING": True,
        "DIM_MUL_IN_ATT": True,
        "SEPARATE_QKV": False,
        "ZERO_DECAY_POS_CLS": False,
    },
    "SVIT": {"O": 4, "LAMBDA_CON": 1.5, "LAMBDA_EDGES": 0.3, "LAMBDA_NODES": 3.7,
             # svit_b200 extension: "" | "replace" | "add" -- scatter the box-conditioned RoI tokens into the
             # object rows of the sequence when SViT.forward is given bboxes
             "BOX_TOKENS": ""},
    "DETECTION": {
        "ENABLE": False,
        "ALIGNED": True,
        "SPATIAL_SCALE_FACTOR": 16,
        "ROI_XFORM_RESOLUTION": 7,
    },
    "TRAIN": {"DATASET": "ssv2", "FORWARD_VIDEO_FRAMES": True, "BATCH_SIZE": 63},
    "TEST": {"BATCH_SIZE": 64},
    "SOLVER": {"CLIP_GRAD_L2NORM": 1.0, "BASE_LR": 2e-4, "WEIGHT_DECAY": 1e-4, "OPTIMIZING_METHOD": "adamw",
               "ZERO_WD_1D_PARAM": True,
               # configs/ssv2.yaml:169-182 (learning-rate policy, svit_b200/optim.py::get_epoch_lr)
               "LR_POLICY": "cosine", "COSINE_END_LR": 2e-6, "COSINE_AFTER_WARMUP": True, "MAX_EPOCH": 50,
               "WARMUP_EPOCHS": 0.0, "WARMUP_START_LR": 2e-6},
}


def ssv2_cfg() -> CfgNode:
    """The MViTv2-S 16x224 + 4 object tokens recipe (configs/ssv2.yaml)."""
    return CfgNode(copy.deepcopy(_SSV2))


def tiny_cfg(frames=4, crop=32, depth=4, classes=10) -> CfgNode:
    """A reduced recipe with the same structure (q-stride-2 + dim doubling at blocks 1 and 3,
    adaptive kv stride) used by the parity tests at sizes the CPU oracle finishes quickly."""
    c = ssv2_cfg()
    c.DATA.NUM_FRAMES = frames
    c.DATA.TRAIN_CROP_SIZE = crop
    c.DATA.TEST_CROP_SIZE = crop
    c.MODEL.NUM_CLASSES = classes
    c.MVIT.DEPTH = depth
    c.MVIT.DIM_MUL = [[1, 2.0], [3, 2.0]][: max(0, (depth) // 2)]
    c.MVIT.HEAD_MUL = [[1, 2.0], [3, 2.0]][: max(0, (depth) // 2)]
    c.MVIT.DIM_MUL = [d for d in c.MVIT.DIM_MUL if d[0] < depth]
    c.MVIT.HEAD_MUL = [d for d in c.MVIT.HEAD_MUL if d[0] < depth]
    qs = []
    for i in range(depth):
        qs.append([i, 1, 2, 2] if i in (1, 3) else [i, 1, 1, 1])
    c.MVIT.POOL_Q_STRIDE = qs
    c.MVIT.POOL_KV_STRIDE_ADAPTIVE = [1, 4, 4]
    c.MVIT.DROPPATH_RATE = 0.1
    return c


_TUPLE_KEYS = ("PATCH_KERNEL", "PATCH_STRIDE", "PATCH_PADDING")


def merge_yaml(cfg: CfgNode, path: str) -> CfgNode:
    """Merge a reference-format YAML over ``cfg`` (unknown sections are kept as data)."""
    import yaml

    with open(path, "r") as f:
        y = yaml.safe_load(f)

    def rec(dst, src):
        for k, v in src.items():
            if isinstance(v, dict):
                if k not in dst or not isinstance(dst[k], dict):
                    dst[k] = CfgNode()
                rec(dst[k], v)
            else:
                if k in _TUPLE_KEYS and isinstance(v, str):
                    v = list(ast.literal_eval(v))
                dst[k] = v

    rec(cfg, y)
    return cfg


def round_width(width, multiplier, min_width=1, divisor=1):
    """Channel/head rounding rule of the reference (slowfast/models/utils.py:16-29)."""
    if not multiplier:
        return width
    width *= multiplier
    min_width = min_width or divisor
    out = max(min_width, int(width + divisor / 2) // divisor * divisor)
    if out < 0.9 * width:
        out += divisor
    return int(out)


def block_specs(cfg: CfgNode):
    """Per-block geometry derived exactly like SViT.__init__ (video_model_builder.py:133-232).

    Returns a list of dicts: dim, dim_out, num_heads, input_size, kernel_q, kernel_kv,
    stride_q, stride_kv, drop_path.  Does not mutate ``cfg``.
    """
    import torch

    depth = cfg.MVIT.DEPTH
    patch_stride = list(cfg.MVIT.PATCH_STRIDE)
    if cfg.MVIT.PATCH_2D:
        patch_stride = [1] + patch_stride
    input_dims = [cfg.DATA.NUM_FRAMES, cfg.DATA.TRAIN_CROP_SIZE, cfg.DATA.TRAIN_CROP_SIZE]
    patch_dims = [input_dims[i] // patch_stride[i] for i in range(3)]
    dim_mul = [1.0] * (depth + 1)
    head_mul = [1.0] * (depth + 1)
    for i, m in cfg.MVIT.DIM_MUL:
        dim_mul[i] = m
    for i, m in cfg.MVIT.HEAD_MUL:
        head_mul[i] = m
    pool_q = [[] for _ in range(depth)]
    pool_kv = [[] for _ in range(depth)]
    stride_q = [[] for _ in range(depth)]
    stride_kv = [[] for _ in range(depth)]
    for ent in cfg.MVIT.POOL_Q_STRIDE:
        stride_q[ent[0]] = list(ent[1:])
        if cfg.MVIT.POOL_KVQ_KERNEL is not None:
            pool_q[ent[0]] = list(cfg.MVIT.POOL_KVQ_KERNEL)
        else:
            pool_q[ent[0]] = [s + 1 if s > 1 else s for s in ent[1:]]
    kv_list = cfg.MVIT.POOL_KV_STRIDE
    if cfg.MVIT.POOL_KV_STRIDE_ADAPTIVE is not None:
        _s = list(cfg.MVIT.POOL_KV_STRIDE_ADAPTIVE)
        kv_list = []
        for i in range(depth):
            if len(stride_q[i]) > 0:
                _s = [max(_s[d] // stride_q[i][d], 1) for d in range(len(_s))]
            kv_list.append([i] + _s)
    for ent in kv_list or []:
        stride_kv[ent[0]] = list(ent[1:])
        if cfg.MVIT.POOL_KVQ_KERNEL is not None:
            pool_kv[ent[0]] = list(cfg.MVIT.POOL_KVQ_KERNEL)
        else:
            pool_kv[ent[0]] = [s + 1 if s > 1 else s for s in ent[1:]]
    dpr = [x.item() for x in torch.linspace(0, cfg.MVIT.DROPPATH_RATE, depth)]
    embed_dim = cfg.MVIT.EMBED_DIM
    num_heads = cfg.MVIT.NUM_HEADS
    input_size = list(patch_dims)
    specs = []
    for i in range(depth):
        num_heads = round_width(num_heads, head_mul[i])
        if cfg.MVIT.DIM_MUL_IN_ATT:
            dim_out = round_width(embed_dim, dim_mul[i], divisor=round_width(num_heads, head_mul[i]))
        else:
            dim_out = round_width(embed_dim, dim_mul[i + 1], divisor=round_width(num_heads, head_mul[i + 1]))
        specs.append(dict(dim=embed_dim, dim_out=dim_out, num_heads=num_heads, input_size=list(input_size),
                          kernel_q=pool_q[i], kernel_kv=pool_kv[i], stride_q=stride_q[i],
                          stride_kv=stride_kv[i], drop_path=dpr[i]))
        if len(stride_q[i]) > 0:
            input_size = [s // st for s, st in zip(input_size, stride_q[i])]
        embed_dim = dim_out
    return specs, patch_dims, embed_dim


def attn_param_shapes(dim, dim_out, num_heads, input_size, stride_q, stride_kv, prefix=""):
    """Parameter names/shapes of one MultiScaleAttention (attention.py:228-327)."""
    hd = dim_out // num_heads
    size = input_size[1]
    q_size = size // stride_q[1] if len(stride_q) > 0 else size
    kv_size = size // stride_kv[1] if len(stride_kv) > 0 else size
    rel = 2 * max(q_size, kv_size) - 1
    s = {
        prefix + "rel_pos_h": (rel, hd), prefix + "rel_pos_w": (rel, hd),
        prefix + "rel_pos_t": (2 * input_size[0] - 1, hd),
        prefix + "qkv.weight": (3 * dim_out, dim), prefix + "qkv.bias": (3 * dim_out,),
        prefix + "proj.weight": (dim_out, dim_out), prefix + "proj.bias": (dim_out,),
    }
    for n in "qkv":
        s[prefix + f"pool_{n}.weight"] = (hd, 1, 3, 3, 3)
        s[prefix + f"norm_{n}.weight"] = (hd,)
        s[prefix + f"norm_{n}.bias"] = (hd,)
    return s


def block_param_shapes(spec, prefix=""):
    """Parameter names/shapes of one MultiScaleBlock (attention.py:498-555)."""
    dim, dout = spec["dim"], spec["dim_out"]
    s = {prefix + "norm1.weight": (dim,), prefix + "norm1.bias": (dim,)}
    s.update(attn_param_shapes(dim, dout, spec["num_heads"], spec["input_size"], spec["stride_q"],
                               spec["stride_kv"], prefix + "attn."))
    s.update({prefix + "norm2.weight": (dout,), prefix + "norm2.bias": (dout,),
              prefix + "mlp.fc1.weight": (4 * dout, dout), prefix + "mlp.fc1.bias": (4 * dout,),
              prefix + "mlp.fc2.weight": (dout, 4 * dout), prefix + "mlp.fc2.bias": (dout,)})
    if dim != dout:
        s.update({prefix + "proj.weight": (dout, dim), prefix + "proj.bias": (dout,)})
    return s


def state_shapes(cfg: CfgNode):
    """All state_dict entries of SViT for ``cfg`` (video_model_builder.py:37-256, 408-470)."""
    specs, patch_dims, final_dim = block_specs(cfg)
    E = cfg.MVIT.EMBED_DIM
    s = {"cls_token": (1, 1, E), "pos_embed_temporal": (1, cfg.DATA.NUM_FRAMES, E),
         "object_queries": (1, cfg.SVIT.O, E),
         "patch_embed.proj.weight": (E, cfg.DATA.INPUT_CHANNEL_NUM[0], *cfg.MVIT.PATCH_KERNEL),
         "patch_embed.proj.bias": (E,)}
    for i, sp in enumerate(specs):
        s.update(block_param_shapes(sp, f"blocks.{i}."))
    s.update({"norm.weight": (final_dim,), "norm.bias": (final_dim,),
              "head.projection.weight": (cfg.MODEL.NUM_CLASSES, final_dim),
              "head.projection.bias": (cfg.MODEL.NUM_CLASSES,),
              "head.boxes_mlp.0.weight": (4, final_dim), "head.boxes_mlp.0.bias": (4,),
              "head.boxes_bce_mlp.weight": (1, final_dim), "head.boxes_bce_mlp.bias": (1,),
              "head.contact_mlp.weight": (5, final_dim), "head.contact_mlp.bias": (5,)})
    return s
