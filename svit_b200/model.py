"""Drop-in SViT model (reference: slowfast/models/video_model_builder.py:24-551) on the sm_100a kernels.

Same constructor (`SViT(cfg)`), same state_dict (405 entries for configs/ssv2.yaml), same forward
contract `model([clip], metadata=None, bboxes=None) -> (preds, extra_preds)`.  The stem (PatchEmbed as an
im2col GEMM whose epilogue writes patch tokens straight into the token sequence), the cls/object-token
assembly, the 16 pooled-attention blocks, the final norm and the cls/object split all run in
libsvit_sm100.so.  The box-conditioned RoIAlign path (DETECTION.ENABLE; reference call sites :385-392,
472-491, whose head_helper.py is missing upstream) is provided as `roi_object_tokens`.

`compute_dtype` selects fp32 (parity mode, CUDA-core fp32 math) or bf16 (tcgen05 tensor-core mode);
parameters stay fp32 nn.Parameters in both.
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn
from torch.nn.init import trunc_normal_

from . import ops
from .config import block_specs
from .msa import MultiScaleBlock


def get_lambdas_dict(cfg):
    """utils/misc.py:411-423."""
    d = {"loss_ce": 1, "boxes_l1_loss": 5 * cfg.SVIT.LAMBDA_NODES, "boxes_bce_loss": 1 * cfg.SVIT.LAMBDA_NODES,
         "boxes_giou_loss": 2 * cfg.SVIT.LAMBDA_NODES, "loss_contact_state": cfg.SVIT.LAMBDA_EDGES}
    if cfg.TRAIN.FORWARD_VIDEO_FRAMES:
        d["video_image_boxes_l1_loss"] = cfg.SVIT.LAMBDA_CON
    return d


class PatchEmbed(nn.Module):
    """Parameter container for the stem conv (stem_helper.py:290-320); compute happens in SViT.forward."""

    def __init__(self, dim_in=3, dim_out=768, kernel=(1, 16, 16), stride=(1, 4, 4), padding=(1, 7, 7), conv_2d=False):
        super().__init__()
        if conv_2d:
            raise NotImplementedError("PATCH_2D")
        self.proj = nn.Conv3d(dim_in, dim_out, kernel_size=tuple(kernel), stride=tuple(stride), padding=tuple(padding))


class SViTHead(nn.Module):
    """Classification + HAOG box / contact heads (video_model_builder.py:408-551)."""

    def __init__(self, cfg, dim_in, num_classes, dropout_rate=0.0, act_func="softmax"):
        super().__init__()
        self.cfg = cfg
        self.T = cfg.DATA.NUM_FRAMES
        if dropout_rate > 0.0:
            self.dropout = nn.Dropout(dropout_rate)
        if cfg.DETECTION.ENABLE:
            raise NotImplementedError("DETECTION.ENABLE: use SViT.roi_object_tokens (reference head_helper.py is absent)")
        if isinstance(num_classes, dict) or num_classes == 0:
            raise NotImplementedError("dict / zero num_classes heads")
        self.projection = nn.Linear(dim_in, num_classes, bias=True)
        if act_func not in ("softmax", "sigmoid"):
            raise NotImplementedError(f"{act_func} is not supported as an activation function.")
        self.act_func = act_func
        self.boxes_mlp = nn.Sequential(nn.Linear(dim_in, 4, bias=True), nn.Sigmoid())
        self.boxes_bce_mlp = nn.Linear(dim_in, 1, bias=True)
        self.contact_mlp = nn.Linear(dim_in, 5, bias=True)
        self._lambdas = get_lambdas_dict(cfg)

    def forward(self, x, T=None, patches=None, bboxes=None):
        """x [B, 1 + T*O, C] = [cls ; object tokens] (any float dtype; head math in fp32)."""
        if T is None:
            T = self.T
        extra = {}
        if hasattr(self, "dropout") and self.training:
            x = self.dropout(x)
        B = x.size(0)
        if not torch.is_grad_enabled():
            # no autograd (inference, the no-grad frames pass): the whole head is one launch (csrc/head_loss.cu)
            O = (x.size(1) - 1) // T
            logits, probs, xobj, pred_bboxes, contact = ops.head_forward(
                x, self.projection.weight, self.projection.bias, self.boxes_mlp[0].weight, self.boxes_mlp[0].bias,
                self.boxes_bce_mlp.weight, self.boxes_bce_mlp.bias, self.contact_mlp.weight, self.contact_mlp.bias, T, O,
                act_sigmoid=self.act_func == "sigmoid", eval_mode=not self.training, want_probs=not self.training)
            extra.update(obj_desc=xobj, logits=logits, pred_bboxes=pred_bboxes, pred_contact_state=contact)
            return (logits if self.training else probs), extra
        x = x.float()
        cls, xobj = x[:, 0].contiguous(), x[:, 1:]
        xobj = xobj.reshape(B, T, -1, xobj.size(-1)).contiguous()
        extra["obj_desc"] = xobj
        logits = ops.linear(cls, self.projection.weight, self.projection.bias)
        extra["logits"] = logits
        out = logits
        if not self.training:
            out = logits.softmax(dim=1) if self.act_func == "softmax" else logits.sigmoid()
        boxes = ops.linear(xobj, self.boxes_mlp[0].weight, self.boxes_mlp[0].bias).sigmoid()
        bce = ops.linear(xobj, self.boxes_bce_mlp.weight, self.boxes_bce_mlp.bias)
        contact = ops.linear(xobj[:, :, :2].contiguous(), self.contact_mlp.weight, self.contact_mlp.bias)
        if not self.training:
            bce = bce.sigmoid()
            contact = contact.softmax(dim=-1)
        extra["pred_bboxes"] = torch.cat((bce, boxes), dim=-1)
        extra["pred_contact_state"] = contact
        return out, extra


class SViT(nn.Module):
    def __init__(self, cfg, compute_dtype: torch.dtype = torch.bfloat16):
        super().__init__()
        assert cfg.DATA.TRAIN_CROP_SIZE == cfg.DATA.TEST_CROP_SIZE
        self.cfg = cfg
        self.compute_dtype = compute_dtype
        if cfg.MVIT.POOL_FIRST or cfg.MVIT.PATCH_2D or cfg.MVIT.USE_ABS_POS or cfg.MVIT.NORM_STEM \
                or cfg.MVIT.DROPOUT_RATE > 0 or not cfg.MVIT.CLS_EMBED_ON:
            raise NotImplementedError("svit_b200 supports the configs/ssv2.yaml family (see DESIGN.md)")
        if cfg.MVIT.NORM != "layernorm":
            raise NotImplementedError("Only supports layernorm.")
        norm_layer = partial(nn.LayerNorm, eps=1e-6)
        embed_dim = cfg.MVIT.EMBED_DIM
        self.patch_stride = list(cfg.MVIT.PATCH_STRIDE)
        self.cls_embed_on = True
        self.num_classes = cfg.MODEL.NUM_CLASSES
        self.patch_embed = PatchEmbed(dim_in=cfg.DATA.INPUT_CHANNEL_NUM[0], dim_out=embed_dim,
                                      kernel=cfg.MVIT.PATCH_KERNEL, stride=cfg.MVIT.PATCH_STRIDE,
                                      padding=cfg.MVIT.PATCH_PADDING)
        specs, self.patch_dims, final_dim = block_specs(cfg)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed_temporal = nn.Parameter(torch.zeros(1, cfg.DATA.NUM_FRAMES, embed_dim))
        self.O = cfg.SVIT.O
        self.object_queries = nn.Parameter(torch.zeros(1, self.O, embed_dim))
        self.blocks = nn.ModuleList()
        for sp in specs:
            self.blocks.append(MultiScaleBlock(
                dim=sp["dim"], dim_out=sp["dim_out"], num_heads=sp["num_heads"], input_size=sp["input_size"],
                mlp_ratio=cfg.MVIT.MLP_RATIO, qkv_bias=cfg.MVIT.QKV_BIAS, drop_rate=cfg.MVIT.DROPOUT_RATE,
                drop_path=sp["drop_path"], norm_layer=norm_layer, kernel_q=sp["kernel_q"], kernel_kv=sp["kernel_kv"],
                stride_q=sp["stride_q"], stride_kv=sp["stride_kv"], mode=cfg.MVIT.MODE, has_cls_embed=True,
                pool_first=False, rel_pos_spatial=cfg.MVIT.REL_POS_SPATIAL, rel_pos_temporal=cfg.MVIT.REL_POS_TEMPORAL,
                rel_pos_zero_init=cfg.MVIT.REL_POS_ZERO_INIT, residual_pooling=cfg.MVIT.RESIDUAL_POOLING,
                dim_mul_in_att=cfg.MVIT.DIM_MUL_IN_ATT, separate_qkv=cfg.MVIT.SEPARATE_QKV))
        self.norm = norm_layer(final_dim)
        self.enable_detection = cfg.DETECTION.ENABLE
        self.head = SViTHead(cfg, final_dim, self.num_classes, dropout_rate=cfg.MODEL.DROPOUT_RATE,
                             act_func=cfg.MODEL.HEAD_ACT)
        trunc_normal_(self.cls_token, std=0.02)
        trunc_normal_(self.pos_embed_temporal, std=0.02)
        trunc_normal_(self.object_queries, std=0.02)
        self.apply(self._init_weights)
        self._lambda = get_lambdas_dict(cfg)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return []  # MVIT.ZERO_DECAY_POS_CLS is false in configs/ssv2.yaml (video_model_builder.py:268-289)

    def forward_tokens(self, clip):
        """clip [B,3,T,H,W] (or [B,3,H,W] frame mode) -> (normed tokens [B, N, C], thw, Tx)."""
        x = clip
        u8 = x.dtype == torch.uint8  # decoded frames [B, T, H, W, 3]: normalised on the device (8f N4)
        if not u8 and x.ndim == 4:
            x = x.unsqueeze(2)
        Tx = x.shape[1] if u8 else x.shape[2]
        Hin, Win = (x.shape[2], x.shape[3]) if u8 else (x.shape[-2], x.shape[-1])
        pe = self.patch_embed.proj
        x = ops.patch_embed_tokens(x, pe.weight, pe.bias, self.cls_token, self.object_queries, self.pos_embed_temporal,
                                   pe.kernel_size, pe.stride, pe.padding, self.compute_dtype,
                                   mean=self.cfg.DATA.MEAN, std=self.cfg.DATA.STD)
        T = self.cfg.DATA.NUM_FRAMES // self.patch_stride[0] if Tx > 1 else Tx
        H = (Hin + 2 * pe.padding[1] - pe.kernel_size[1]) // pe.stride[1] + 1
        W = (Win + 2 * pe.padding[2] - pe.kernel_size[2]) // pe.stride[2] + 1
        thw = [T, H, W]
        for blk in self.blocks:
            x, thw = blk(x, thw)
        x = ops.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        return x, thw, Tx

    def forward(self, x, metadata=None, bboxes=None):
        clip = x[0]
        tokens, thw, Tx = self.forward_tokens(clip)
        O_tot = Tx * self.O
        roi = None
        if bboxes is not None:
            roi = self.roi_object_tokens(tokens, thw, bboxes)
            mode = getattr(self.cfg.SVIT, "BOX_TOKENS", "")  # "", "replace" or "add" (svit_b200 extension)
            if mode:
                # the box-conditioned tokens take the place of (or are added to) the learned object tokens in the
                # sequence, rows 1 + T'H'W' + t*K + k, so the head reads them
                det = self.cfg.DETECTION
                tokens, _ = ops.roi_scatter_tokens(tokens, thw, bboxes, self.patch_stride[0], 1.0 / det.SPATIAL_SCALE_FACTOR,
                                                   det.ROI_XFORM_RESOLUTION, mode)
        cls_obj = ops.gather_cls_obj(tokens, O_tot)
        out, extra = self.head(cls_obj, T=Tx)
        if roi is not None:
            extra["roi_tokens"], extra["roi_assign"] = roi
        return out, extra

    def roi_object_tokens(self, tokens, thw, bboxes):
        """Per-frame box-conditioned tokens: RoIAlign(7x7, scale 1/SPATIAL_SCALE_FACTOR, aligned) of each
        (frame, box) on temporal slice t // patch_stride_t of the final patch grid, max over bins."""
        det = self.cfg.DETECTION
        return ops.roi_tokens(tokens, thw, bboxes, self.patch_stride[0], 1.0 / det.SPATIAL_SCALE_FACTOR,
                              det.ROI_XFORM_RESOLUTION)
