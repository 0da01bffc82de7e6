"""svit_b200 -- B200-native (sm_100a) implementation of SViT's pooled-attention block with object tokens.

Drop-in modules: MultiScaleAttention, MultiScaleBlock (reference slowfast/models/attention.py) and SViT
(reference slowfast/models/video_model_builder.py).  All compute goes through libsvit_sm100.so
(include/svit_b200.h); there is no CPU or torch-op fallback.
"""
from .config import CfgNode, block_specs, merge_yaml, ssv2_cfg, state_shapes, tiny_cfg  # noqa: F401
from .graph import GraphedForward, GraphedTrainStep  # noqa: F401
from .model import PatchEmbed, SViT, SViTHead  # noqa: F401
from .msa import DropPath, Mlp, MultiScaleAttention, MultiScaleBlock, attention_pool  # noqa: F401

__all__ = ["SViT", "SViTHead", "PatchEmbed", "MultiScaleAttention", "MultiScaleBlock", "Mlp", "DropPath",
           "attention_pool", "GraphedForward", "GraphedTrainStep", "ssv2_cfg", "tiny_cfg", "block_specs", "state_shapes", "merge_yaml", "CfgNode"]
