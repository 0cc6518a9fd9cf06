"""Multi-level ResNet backbone + sine position encoding for ParSeDA.

north_star keeps the backbone on torch/cuDNN; this file only restates the thin wrapper the model
needs offline (no hard-coded weight paths): /root/reference/models/DDETR_backbone.py:31-169 and
models/position_encoding.py:22-58.  Parameter/buffer names match (`backbone.0.body.*`).
"""
import math
import os
from typing import Dict, List

import torch
import torch.nn.functional as F
import torchvision
from torch import nn
from torchvision.models._utils import IntermediateLayerGetter

from . import dense, streams
from .nested import NestedTensor


class FrozenBatchNorm2d(nn.Module):
    """BatchNorm2d with fixed statistics and affine parameters (buffers), eps inside the rsqrt
    (DDETR_backbone.py:31-68)."""

    def __init__(self, n, eps=1e-5):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))
        self.eps = eps

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        state_dict.pop(prefix + "num_batches_tracked", None)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)

    def forward(self, x):
        scale = (self.weight * (self.running_var + self.eps).rsqrt()).reshape(1, -1, 1, 1)
        bias = self.bias.reshape(1, -1, 1, 1) - self.running_mean.reshape(1, -1, 1, 1) * scale
        return x * scale + bias


_FUSED_CONV = os.environ.get("RLIPV2_FUSED_CONV", "1") != "0"      # A/B switch for measurements
_POS_STREAM = os.environ.get("RLIPV2_POS_STREAM", "1") != "0"       # masks + position embeddings beside the backbone
_CONV_WGRAD_STREAM = os.environ.get("RLIPV2_CONV_WGRAD_STREAM", "0") != "0"   # conv weight gradients on the side stream (measured r01s4e: 29.95 vs 29.67 ms/step without - off)
# channels_last activations through the backbone: cuDNN's TF32 kernels are NHWC-native, so NCHW tensors cost a
# nchwToNhwc / nhwcToNchw pair around every convolution (3.3 ms per step); with BN folded and bias / residual /
# ReLU fused into the convolutions there is no NCHW-favouring elementwise pass left (measured: 44.0 -> 39.5 ms).
_BACKBONE_NHWC = os.environ.get("RLIPV2_BACKBONE_NHWC", "1") != "0"


class _ConvBiasReLU(torch.autograd.Function):
    """relu(conv(x, w) + b [+ residual]) as one cuDNN call (bias / residual / ReLU in the convolution's
    epilogue) with a hand-written backward: ReLU mask, then cuDNN dgrad / wgrad.  `b` is the frozen-BN
    shift (a constant), so no bias gradient is produced.  Replaces conv + broadcast add (+ add) + relu:
    3-4 kernels -> 1 in the forward of every trainable bottleneck convolution."""

    @staticmethod
    def forward(ctx, x, w, b, residual, stride, padding, dilation, groups, leaf=None, leaf_scale=None):
        if residual is not None:
            y = torch.cudnn_convolution_add_relu(x, w, residual, 1.0, b, stride, padding, dilation, groups)
        else:
            y = torch.cudnn_convolution_relu(x, w, b, stride, padding, dilation, groups)
        ctx.save_for_backward(x, w, y)
        ctx.conf = (stride, padding, dilation, groups)
        ctx.has_residual = residual is not None
        # `w` = leaf * leaf_scale (BN folded in); when the leaf's .grad is a view of the step's flat gradient buffer the
        # weight gradient is formed and accumulated on the parameter-gradient side stream (dense.py), off the chain of
        # input gradients that the previous block is waiting for
        ctx.leaf, ctx.leaf_scale = leaf, leaf_scale
        return y

    @staticmethod
    def backward(ctx, grad_out):
        x, w, y = ctx.saved_tensors
        stride, padding, dilation, groups = ctx.conf
        g = torch.ops.aten.threshold_backward(grad_out, y, 0)
        leaf = ctx.leaf
        side = (ctx.needs_input_grad[1] and leaf is not None and dense._WGRAD_STREAM and _CONV_WGRAD_STREAM
                and getattr(leaf, "_fuse_grad", False)
                and leaf.grad is not None and g.is_cuda)
        if side:
            with dense._ParamGradSide(g.device, g, x):
                _, gw, _ = torch.ops.aten.convolution_backward(
                    g, x, w, None, stride, padding, dilation, False, [0, 0], groups, [False, True, False])
                leaf.grad.addcmul_(gw, ctx.leaf_scale.view(-1, 1, 1, 1))
            gx = None
            if ctx.needs_input_grad[0]:
                gx, _, _ = torch.ops.aten.convolution_backward(
                    g, x, w, None, stride, padding, dilation, False, [0, 0], groups, [True, False, False])
            gw = None
        else:
            gx, gw, _ = torch.ops.aten.convolution_backward(
                g, x, w, None, stride, padding, dilation, False, [0, 0], groups,
                [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False])
        gres = g if (ctx.has_residual and ctx.needs_input_grad[3]) else None
        return gx, gw, None, gres, None, None, None, None, None, None


class PositionEmbeddingSine(nn.Module):
    """Normalised 2-D sine embedding over the un-padded extent (position_encoding.py:22-58).  `offset` is subtracted
    from the cumulative cell index before normalising: 0 for the ResNet wrapper, 0.5 (cell centres) for the Swin
    wrapper's copy (swin/position_encoding.py:34-35)."""

    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None, offset=0.0):
        super().__init__()
        self.offset = float(offset)
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature
        self.normalize = normalize
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.scale = 2 * math.pi if scale is None else scale

    def forward(self, tensor_list: NestedTensor):
        mask = tensor_list.mask
        assert mask is not None
        not_mask = ~mask
        y_embed = not_mask.cumsum(1, dtype=torch.float32)
        x_embed = not_mask.cumsum(2, dtype=torch.float32)
        if self.normalize:
            eps = 1e-6
            if self.offset:
                y_embed = (y_embed - self.offset) / (y_embed[:, -1:, :] + eps) * self.scale
                x_embed = (x_embed - self.offset) / (x_embed[:, :, -1:] + eps) * self.scale
            else:
                y_embed = y_embed / (y_embed[:, -1:, :] + eps) * self.scale
                x_embed = x_embed / (x_embed[:, :, -1:] + eps) * self.scale
        dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32, device=mask.device)
        dim_t = self.temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / self.num_pos_feats)
        pos_x = x_embed[:, :, :, None] / dim_t
        pos_y = y_embed[:, :, :, None] / dim_t
        pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
        pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
        return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


class Backbone(nn.Module):
    """ResNet with FrozenBatchNorm returning C3, C4, C5 (strides 8/16/32); stem + layer1 frozen
    (DDETR_backbone.py:71-138)."""

    def __init__(self, name="resnet50", train_backbone=True, return_interm_layers=True, dilation=False,
                 weights_path=None):
        super().__init__()
        assert name not in ("resnet18", "resnet34"), "number of channels are hard coded"
        backbone = getattr(torchvision.models, name)(
            replace_stride_with_dilation=[False, False, dilation], weights=None, norm_layer=FrozenBatchNorm2d)
        if weights_path:
            backbone.load_state_dict(torch.load(weights_path, map_location="cpu"), strict=True)
        for pname, parameter in backbone.named_parameters():
            if not train_backbone or ("layer2" not in pname and "layer3" not in pname and "layer4" not in pname):
                parameter.requires_grad_(False)
        if return_interm_layers:
            return_layers = {"layer2": "0", "layer3": "1", "layer4": "2"}
            self.strides = [8, 16, 32]
            self.num_channels = [512, 1024, 2048]
        else:
            return_layers = {"layer4": "0"}
            self.strides = [32]
            self.num_channels = [2048]
        if dilation:
            self.strides[-1] = self.strides[-1] // 2
        self.body = IntermediateLayerGetter(backbone, return_layers=return_layers)
        self.fold_bn = os.environ.get("RLIPV2_FOLD_BN", "1") != "0"    # A/B switch for measurements

    # ---- frozen-BN folding ----------------------------------------------------------------------------
    # The reference applies FrozenBatchNorm2d as separate broadcast mul / add kernels after every
    # convolution (DDETR_backbone.py:59-68) and ReLU as a third pass: on the 200x334 and 400x667 maps of
    # an 800x1333 image those three elementwise passes cost 4-5x the convolution itself (profiles/
    # train_step_r01_v3_kernels.txt: 96 + 95 + 40 us vs a 43-53 us conv).  The statistics are constants,
    # so BN folds into the convolution exactly: conv(x, w) * s + t == conv(x, w * s) + t.  The folded
    # weight is recomputed from the live parameters every call (a [Cout,Cin,k,k] multiply), so gradients
    # still reach `conv.weight` through the product and checkpoints are untouched.
    @staticmethod
    def _bn_affine(bn):
        """(scale, shift) of a FrozenBatchNorm2d.  The four buffers are constants between checkpoint loads,
        so the pair is computed once and cached on the module, keyed on the buffers' version counters and
        storage (load_state_dict / .to() invalidate it): 6 tiny kernels per convolution per step saved."""
        bufs = (bn.weight, bn.bias, bn.running_mean, bn.running_var)
        key = tuple((b._version, b.data_ptr()) for b in bufs)
        cache = getattr(bn, "_affine_cache", None)
        if cache is None or cache[0] != key:
            with torch.no_grad():
                scale = bn.weight * (bn.running_var + bn.eps).rsqrt()
                shift = bn.bias - bn.running_mean * scale
            cache = (key, scale, shift)
            bn._affine_cache = cache
        return cache[1], cache[2]

    @classmethod
    def _fold(cls, conv, bn):
        scale, shift = cls._bn_affine(bn)
        w = conv.weight
        if not w.requires_grad:
            # frozen stem / layer1: the folded weight is a constant too
            key = (w._version, w.data_ptr(), scale.data_ptr())
            cache = getattr(conv, "_folded_cache", None)
            if cache is None or cache[0] != key:
                with torch.no_grad():
                    cache = (key, w * scale.view(-1, 1, 1, 1))
                conv._folded_cache = cache
            return cache[1], shift
        return w * scale.view(-1, 1, 1, 1), shift

    @classmethod
    def _conv_bn(cls, x, conv, bn, relu, residual=None, extra_bias=None, bias=True):
        """conv + folded frozen BN (+ residual) (+ ReLU).  On CUDA every conv that ends in a ReLU runs as ONE
        cuDNN call with the bias / residual / ReLU epilogue fused (`_ConvBiasReLU` when gradients are
        needed); `bias=False` returns the bare convolution and hands the shift to the caller, which adds
        it to the bias of the fused conv that consumes the result (`extra_bias`)."""
        w, b = cls._fold(conv, bn)
        if extra_bias is not None:
            b = b + extra_bias
        if not bias:
            return F.conv2d(x, w, None, conv.stride, conv.padding, conv.dilation, conv.groups), b
        fused = x.is_cuda and conv.groups == 1 and (relu or residual is not None) and _FUSED_CONV
        if fused:
            if torch.is_grad_enabled() and (w.requires_grad or x.requires_grad):
                leaf = conv.weight if conv.weight.requires_grad else None
                return _ConvBiasReLU.apply(x, w, b, residual, conv.stride, conv.padding, conv.dilation, conv.groups,
                                           leaf, cls._bn_affine(bn)[0] if leaf is not None else None)
            # cuDNN's fused conv + bias (+ residual) + ReLU epilogue (no autograd needed here)
            if residual is not None:
                return torch.cudnn_convolution_add_relu(x, w, residual, 1.0, b, conv.stride, conv.padding,
                                                        conv.dilation, conv.groups)
            return torch.cudnn_convolution_relu(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
        y = F.conv2d(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
        if residual is not None:
            y = y + residual
        return F.relu_(y) if (relu or residual is not None) else y

    @classmethod
    def _bottleneck(cls, x, blk):
        out = cls._conv_bn(x, blk.conv1, blk.bn1, True)
        out = cls._conv_bn(out, blk.conv2, blk.bn2, True)
        if blk.downsample is None:
            return cls._conv_bn(out, blk.conv3, blk.bn3, True, residual=x)
        if x.is_cuda and _FUSED_CONV:
            # the projection shortcut's BN shift rides on conv3's fused bias: one kernel fewer per block
            identity, shift = cls._conv_bn(x, blk.downsample[0], blk.downsample[1], False, bias=False)
            return cls._conv_bn(out, blk.conv3, blk.bn3, True, residual=identity, extra_bias=shift)
        identity = cls._conv_bn(x, blk.downsample[0], blk.downsample[1], False)
        return cls._conv_bn(out, blk.conv3, blk.bn3, True, residual=identity)

    def _forward_folded(self, x):
        body = self.body
        if _BACKBONE_NHWC and x.is_cuda:
            x = x.contiguous(memory_format=torch.channels_last)
        frozen_stem = not any(p.requires_grad for p in body.layer1.parameters())
        with torch.set_grad_enabled(torch.is_grad_enabled() and not frozen_stem):
            x = self._conv_bn(x, body.conv1, body.bn1, True)         # stem + layer1 are frozen (:75-77)
            x = body.maxpool(x)
            for blk in body.layer1:
                x = self._bottleneck(x, blk)
        outs = {}
        for name in ("layer2", "layer3", "layer4"):
            if not hasattr(body, name):
                break
            for blk in getattr(body, name):
                x = self._bottleneck(x, blk)
            if name in body.return_layers:
                outs[body.return_layers[name]] = x
        return outs

    def forward(self, tensor_list: NestedTensor, defer_masks: bool = False) -> Dict[str, NestedTensor]:
        xs = self._forward_folded(tensor_list.tensors) if self.fold_bn else self.body(tensor_list.tensors)
        out = {}
        for name, x in xs.items():
            m = tensor_list.mask
            assert m is not None
            # `defer_masks` (Joiner on a GPU): the masks only need the feature maps' shapes, so the caller forms them
            # (and the position embeddings) on a side stream beside the convolutions
            mask = None if defer_masks else F.interpolate(m[None].float(), size=x.shape[-2:]).to(torch.bool)[0]
            out[name] = NestedTensor(x, mask)
        return out


class Joiner(nn.Sequential):
    """(backbone, position embedding) -> (feature NestedTensors, position embeddings)."""

    def __init__(self, backbone, position_embedding):
        super().__init__(backbone, position_embedding)
        self.strides = backbone.strides
        self.num_channels = backbone.num_channels

    def forward(self, tensor_list: NestedTensor):
        x_in = tensor_list.tensors
        if _POS_STREAM and x_in.is_cuda and hasattr(self[0], "strides"):
            # The per-level padding masks and sine position embeddings (position_encoding.py:22-58: ~30 small kernels
            # per level, no gradients) depend on the input mask and on the feature maps' SHAPES only.  They are issued
            # on a side stream that forks before the backbone's convolutions and joins after them.
            cur = torch.cuda.current_stream(x_in.device)
            if getattr(self, "_pos_stream", None) is None:
                self._pos_stream = streams.get(x_in.device, "pos")
            side = self._pos_stream
            side.wait_stream(cur)                               # fork point: before the convolutions are issued
            xs = self[0](tensor_list, defer_masks=True)
            out = [x for _, x in sorted(xs.items())]
            with torch.cuda.stream(side):
                m = tensor_list.mask[None].float()
                for x in out:
                    x.mask = F.interpolate(m, size=x.tensors.shape[-2:]).to(torch.bool)[0]
                pos = [self[1](x).to(x.tensors.dtype) for x in out]
            cur.wait_stream(side)
            for t in [x.mask for x in out] + pos:
                t.record_stream(cur)
            return out, pos
        xs = self[0](tensor_list)
        out: List[NestedTensor] = [x for _, x in sorted(xs.items())]
        pos = [self[1](x).to(x.tensors.dtype) for x in out]
        return out, pos


def build_backbone(args):
    """build_DDETR_backbone (DDETR_backbone.py:163-169) minus the hard-coded weight path: pass
    `args.backbone_weights` (a resnet50 state_dict file) to start from pretrained weights."""
    if getattr(args, "position_embedding", "sine") not in ("v2", "sine"):
        raise NotImplementedError("ParSeDA scripts use the sine position embedding")
    if "swin" in args.backbone:
        # detr.py:326-327 -> models/swin/backbone.py:194-205 (BASELINE config 4: --backbone swin_large)
        from .swin import SwinBackbone
        backbone = SwinBackbone(args.backbone, args.num_feature_levels, getattr(args, "pretrained_swin", ""),
                                getattr(args, "use_checkpoint", False), getattr(args, "drop_path_rate", 0.2), args.dilation)
        return Joiner(backbone, PositionEmbeddingSine(args.hidden_dim // 2, normalize=True, offset=0.5))
    position_embedding = PositionEmbeddingSine(args.hidden_dim // 2, normalize=True)
    backbone = Backbone(args.backbone, args.lr_backbone > 0, args.masks or (args.num_feature_levels > 1),
                        args.dilation, getattr(args, "backbone_weights", None))
    return Joiner(backbone, position_embedding)
