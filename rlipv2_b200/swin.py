"""Swin Transformer backbone for ParSeDA (BASELINE config 4: `--backbone swin_large`).

north_star keeps the backbone on torch / cuDNN; this file restates the hierarchical shifted-window
network the reference vendors (/root/reference/models/swin/swin_transformer.py:158-763, wrapper
models/swin/backbone.py:63-205, position embedding models/swin/position_encoding.py:14-46) without timm,
with the reference's parameter / buffer names (`backbone.0.body.patch_embed.proj.weight`,
`...layers.2.blocks.17.attn.relative_position_bias_table`, `...norm3.weight`, ...), so checkpoints load
unchanged.

What is done differently from the reference's module code (same arithmetic, fp32 rounding aside):
  * the shifted-window region mask depends on the padded feature-map size only: it is built once per
    (Hp, Wp, window, shift, device) and cached, merged with the head's relative-position bias into ONE additive
    [nW, heads, T, T] term, and the window attention is a single `scaled_dot_product_attention` call
    (no [windows, heads, T, T] score tensor round trips through HBM for scale / bias / mask / softmax / dropout);
  * window partition and its inverse are one view + permute each, shared by the shifted and plain blocks;
  * stochastic depth is a per-sample Bernoulli mask scaled by 1 / keep, drawn as timm's DropPath draws it.
"""
from typing import Dict

import torch
import torch.nn.functional as F
from torch import nn

from .nested import NestedTensor

# (depths, embed_dim, heads) per `--backbone` name substring (models/swin/backbone.py:105-163)
_VARIANTS = {
    "tiny": ((2, 2, 6, 2), 96, (3, 6, 12, 24)),
    "small": ((2, 2, 18, 2), 96, (3, 6, 12, 24)),
    "base": ((2, 2, 18, 2), 128, (4, 8, 16, 32)),
    "large": ((2, 2, 18, 2), 192, (6, 12, 24, 48)),
}


def swin_variant(name: str):
    """-> dict(depths, embed_dim, num_heads, window_size, pretrain_img_size) for a `--backbone swin_*` name"""
    key = next((k for k in ("small", "base", "large") if k in name), "tiny")
    depths, dim, heads = _VARIANTS[key]
    big = "384" in name and key in ("base", "large")
    return dict(depths=depths, embed_dim=dim, num_heads=heads, window_size=12 if big else 7,
                pretrain_img_size=384 if big else 224)


class DropPath(nn.Module):
    """Stochastic depth per sample: x * Bernoulli(keep) / keep on the residual branch (timm.models.layers.DropPath,
    imported at swin_transformer.py:21)."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        if keep > 0.0:
            mask.div_(keep)
        return x * mask


def to_windows(x, ws: int):
    """[B, Hp, Wp, C] -> [B * nW, ws * ws, C]  (window_partition, swin_transformer.py:181-194)"""
    B, Hp, Wp, C = x.shape
    x = x.view(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, ws * ws, C)


def from_windows(w, ws: int, Hp: int, Wp: int):
    """inverse of `to_windows`  (window_reverse, swin_transformer.py:197-211)"""
    C = w.shape[-1]
    x = w.view(-1, Hp // ws, Wp // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, Hp, Wp, C)


_REGION_MASKS = {}


def shifted_window_mask(Hp: int, Wp: int, ws: int, shift: int, device):
    """[nW, T, T] additive mask: 0 where two cells of a (cyclically shifted) window come from the same image region,
    -100 otherwise (swin_transformer.py:401-421).  A function of the padded size only - cached."""
    key = (Hp, Wp, ws, shift, str(device))
    m = _REGION_MASKS.get(key)
    if m is None:
        region = torch.zeros((1, Hp, Wp, 1), device=device)
        bands = (slice(0, -ws), slice(-ws, -shift), slice(-shift, None))
        for i, hs in enumerate(bands):
            for j, wsl in enumerate(bands):
                region[:, hs, wsl, :] = 3 * i + j
        r = to_windows(region, ws).squeeze(-1)                       # [nW, T]
        m = (r[:, None, :] != r[:, :, None]).to(torch.float32) * -100.0
        _REGION_MASKS[key] = m
    return m


class Mlp(nn.Module):
    def __init__(self, dim, hidden, drop=0.0):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class WindowAttention(nn.Module):
    """Multi-head self-attention inside one window with a learned relative-position bias
    (swin_transformer.py:214-297)."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, tuple(window_size), num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        wh, ww = self.window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * wh - 1) * (2 * ww - 1), num_heads))
        # index of (dy, dx) between every pair of cells of a window into the (2wh-1) x (2ww-1) table
        ys, xs = torch.meshgrid(torch.arange(wh), torch.arange(ww), indexing="ij")
        ys, xs = ys.flatten(), xs.flatten()
        dy = ys[:, None] - ys[None, :] + (wh - 1)
        dx = xs[:, None] - xs[None, :] + (ww - 1)
        self.register_buffer("relative_position_index", dy * (2 * ww - 1) + dx)
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)

    def position_bias(self):
        """[heads, T, T]"""
        T = self.window_size[0] * self.window_size[1]
        return self.relative_position_bias_table[self.relative_position_index.view(-1)].view(T, T, -1).permute(2, 0, 1)

    def forward(self, x, mask=None):
        """x [B * nW, T, C]; mask [nW, T, T] additive or None"""
        Bw, T, C = x.shape
        H = self.num_heads
        q, k, v = self.qkv(x).view(Bw, T, 3, H, C // H).permute(2, 0, 3, 1, 4)        # each [Bw, H, T, hd]
        bias = self.position_bias()[None]                                               # [1, H, T, T]
        if mask is not None:
            nW = mask.shape[0]
            bias = (bias + mask[:, None]).to(q.dtype)                                   # [nW, H, T, T]
            q, k, v = (t.reshape(Bw // nW, nW * H, T, C // H) for t in (q, k, v))
            bias = bias.reshape(1, nW * H, T, T)
        p = self.attn_drop.p if self.training else 0.0
        if q.is_cuda and q.dtype == torch.float32 and not torch.backends.cuda.matmul.allow_tf32:
            # IEEE fp32 products requested (dense.set_matmul_precision('fp32'), the parity tests): the fused attention
            # kernels may take TF32 products for fp32 operands, the math backend follows the global switch
            from torch.nn.attention import SDPBackend, sdpa_kernel
            with sdpa_kernel(SDPBackend.MATH):
                out = F.scaled_dot_product_attention(q, k, v, attn_mask=bias.to(q.dtype), dropout_p=p, scale=self.scale)
        else:
            out = F.scaled_dot_product_attention(q, k, v, attn_mask=bias.to(q.dtype), dropout_p=p, scale=self.scale)
        out = out.reshape(Bw, H, T, C // H).transpose(1, 2).reshape(Bw, T, C)
        return self.proj_drop(self.proj(out))


class SwinTransformerBlock(nn.Module):
    """LN -> (shifted) window attention -> residual; LN -> MLP -> residual (swin_transformer.py:300-402)."""

    def __init__(self, dim, num_heads, window_size=7, shift_size=0, mlp_ratio=4.0, qkv_bias=True, qk_scale=None,
                 drop=0.0, attn_drop=0.0, drop_path=0.0):
        super().__init__()
        assert 0 <= shift_size < window_size, "shift_size must in 0-window_size"
        self.dim, self.num_heads, self.window_size, self.shift_size = dim, num_heads, window_size, shift_size
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, (window_size, window_size), num_heads, qkv_bias, qk_scale, attn_drop, drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), drop)

    def forward(self, x, H: int, W: int):
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        ws, s = self.window_size, self.shift_size
        y = self.norm1(x).view(B, H, W, C)
        pad_r, pad_b = (-W) % ws, (-H) % ws
        if pad_r or pad_b:
            y = F.pad(y, (0, 0, 0, pad_r, 0, pad_b))          # zero cells take part in the windows, as in the reference
        Hp, Wp = H + pad_b, W + pad_r
        mask = None
        if s > 0:
            y = torch.roll(y, shifts=(-s, -s), dims=(1, 2))
            mask = shifted_window_mask(Hp, Wp, ws, s, x.device)
        y = from_windows(self.attn(to_windows(y, ws), mask), ws, Hp, Wp)
        if s > 0:
            y = torch.roll(y, shifts=(s, s), dims=(1, 2))
        if pad_r or pad_b:
            y = y[:, :H, :W, :]
        x = x + self.drop_path(y.reshape(B, L, C))
        return x + self.drop_path(self.mlp(self.norm2(x)))


class PatchMerging(nn.Module):
    """2 x 2 neighbourhood -> 4C -> LN -> Linear 2C (swin_transformer.py:405-441)."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(4 * dim)

    def forward(self, x, H: int, W: int):
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        x = x.view(B, H, W, C)
        if H % 2 or W % 2:
            x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
        H2, W2 = (H + 1) // 2, (W + 1) // 2
        # channel blocks ordered (row 0, col 0), (1, 0), (0, 1), (1, 1): one permute of the 2 x 2 view
        x = x.view(B, H2, 2, W2, 2, C).permute(0, 1, 3, 4, 2, 5).reshape(B, H2 * W2, 4 * C)
        return self.reduction(self.norm(x))


class BasicLayer(nn.Module):
    """One resolution stage: `depth` blocks alternating plain / shifted windows, then patch merging
    (swin_transformer.py:444-537)."""

    def __init__(self, dim, depth, num_heads, window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop=0.0,
                 attn_drop=0.0, drop_path=0.0, downsample=True, use_checkpoint=False):
        super().__init__()
        self.window_size, self.shift_size, self.depth, self.use_checkpoint = window_size, window_size // 2, depth, use_checkpoint
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim, num_heads, window_size, 0 if i % 2 == 0 else window_size // 2, mlp_ratio,
                                 qkv_bias, qk_scale, drop, attn_drop,
                                 drop_path[i] if isinstance(drop_path, (list, tuple)) else drop_path)
            for i in range(depth)])
        self.downsample = PatchMerging(dim) if downsample else None

    def forward(self, x, H: int, W: int):
        for blk in self.blocks:
            if self.use_checkpoint and torch.is_grad_enabled():
                from torch.utils.checkpoint import checkpoint
                x = checkpoint(blk, x, H, W, use_reentrant=False)
            else:
                x = blk(x, H, W)
        if self.downsample is None:
            return x, x, H, W
        return x, self.downsample(x, H, W), (H + 1) // 2, (W + 1) // 2


class PatchEmbed(nn.Module):
    """4 x 4 stride-4 convolution (+ LayerNorm over channels), input padded up to a multiple of the patch
    (swin_transformer.py:540-582)."""

    def __init__(self, patch_size=4, in_chans=3, embed_dim=96, norm=True):
        super().__init__()
        self.patch_size = (patch_size, patch_size)
        self.in_chans, self.embed_dim = in_chans, embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.LayerNorm(embed_dim) if norm else None

    def forward(self, x):
        """-> tokens [B, Wh * Ww, C], Wh, Ww"""
        ph, pw = self.patch_size
        H, W = x.shape[-2:]
        if H % ph or W % pw:
            x = F.pad(x, (0, (-W) % pw, 0, (-H) % ph))
        x = self.proj(x)
        Wh, Ww = x.shape[-2:]
        x = x.flatten(2).transpose(1, 2)
        return (x if self.norm is None else self.norm(x)), Wh, Ww


class SwinTransformer(nn.Module):
    """Returns {'layer<i>': [B, C_i, H_i, W_i]} for i in out_indices (swin_transformer.py:585-763)."""

    def __init__(self, pretrain_img_size=224, patch_size=4, in_chans=3, embed_dim=96, depths=(2, 2, 6, 2),
                 num_heads=(3, 6, 12, 24), window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.2, ape=False, patch_norm=True, out_indices=(0, 1, 2, 3),
                 frozen_stages=-1, use_checkpoint=False):
        super().__init__()
        self.num_layers, self.embed_dim, self.ape = len(depths), embed_dim, ape
        self.out_indices, self.frozen_stages = tuple(out_indices), frozen_stages
        self.patch_embed = PatchEmbed(patch_size, in_chans, embed_dim, norm=patch_norm)
        if ape:
            side = pretrain_img_size // patch_size
            self.absolute_pos_embed = nn.Parameter(torch.zeros(1, embed_dim, side, side))
            nn.init.trunc_normal_(self.absolute_pos_embed, std=0.02)
        self.pos_drop = nn.Dropout(drop_rate)
        total = sum(depths)
        rates = torch.linspace(0, drop_path_rate, total, device="cpu").tolist()         # stochastic-depth decay rule
        self.layers = nn.ModuleList()
        for i, depth in enumerate(depths):
            first = sum(depths[:i])
            self.layers.append(BasicLayer(int(embed_dim * 2 ** i), depth, num_heads[i], window_size, mlp_ratio, qkv_bias,
                                          qk_scale, drop_rate, attn_drop_rate, rates[first:first + depth],
                                          downsample=i < self.num_layers - 1, use_checkpoint=use_checkpoint))
        self.num_features = [int(embed_dim * 2 ** i) for i in range(self.num_layers)]
        for i in self.out_indices:
            self.add_module(f"norm{i}", nn.LayerNorm(self.num_features[i]))
        self._freeze_stages()

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            self.patch_embed.eval()
            for p in self.patch_embed.parameters():
                p.requires_grad = False
        if self.frozen_stages >= 1 and self.ape:
            self.absolute_pos_embed.requires_grad = False
        if self.frozen_stages >= 2:
            self.pos_drop.eval()
            for i in range(self.frozen_stages - 1):
                self.layers[i].eval()
                for p in self.layers[i].parameters():
                    p.requires_grad = False

    def init_weights(self, pretrained=None):
        """trunc-normal Linear weights, unit LayerNorms; then optionally a checkpoint whose relative-position tables are
        bicubically resized to this window size (swin_transformer.py:101-155, 694-719)."""
        def _init(m):
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

        if pretrained is not None and not isinstance(pretrained, str):
            raise TypeError("pretrained must be a str or None")
        self.apply(_init)
        if pretrained:
            load_swin_checkpoint(self, pretrained)

    def forward(self, x) -> Dict[str, torch.Tensor]:
        x, H, W = self.patch_embed(x)
        if self.ape:
            ape = F.interpolate(self.absolute_pos_embed, size=(H, W), mode="bicubic")
            x = x + ape.flatten(2).transpose(1, 2)
        x = self.pos_drop(x)
        outs = {}
        for i, layer in enumerate(self.layers):
            x_out, x, H2, W2 = layer(x, H, W)
            if i in self.out_indices:
                y = getattr(self, f"norm{i}")(x_out)
                outs[f"layer{i}"] = y.view(-1, H, W, self.num_features[i]).permute(0, 3, 1, 2).contiguous()
            H, W = H2, W2
        return outs


def adapt_swin_state_dict(model: SwinTransformer, state_dict):
    """Checkpoint -> this model's shapes: strips 'module.', reshapes a token-major absolute position embedding and
    bicubically resizes relative-position tables trained at another window size (swin_transformer.py:123-153)."""
    for key in ("state_dict", "model"):
        if key in state_dict and isinstance(state_dict[key], dict):
            state_dict = state_dict[key]
            break
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
    own = model.state_dict()
    ape = sd.get("absolute_pos_embed")
    if ape is not None and ape.dim() == 3 and "absolute_pos_embed" in own:
        n2, c2, h, w = own["absolute_pos_embed"].shape
        if ape.shape[0] == n2 and ape.shape[2] == c2 and ape.shape[1] == h * w:
            sd["absolute_pos_embed"] = ape.view(n2, h, w, c2).permute(0, 3, 1, 2)
    for k in [k for k in sd if "relative_position_bias_table" in k and k in own]:
        src, dst = sd[k], own[k]
        if src.shape[1] != dst.shape[1]:
            sd.pop(k)                                                   # head count differs: keep the initialisation
        elif src.shape[0] != dst.shape[0]:
            s1, s2 = int(src.shape[0] ** 0.5), int(dst.shape[0] ** 0.5)
            r = F.interpolate(src.t().reshape(1, -1, s1, s1), size=(s2, s2), mode="bicubic")
            sd[k] = r.reshape(dst.shape[1], dst.shape[0]).t()
    return sd


def load_swin_checkpoint(model: SwinTransformer, path: str, strict: bool = False):
    ckpt = torch.load(path, map_location="cpu")
    if not isinstance(ckpt, dict):
        raise RuntimeError(f"No state_dict found in checkpoint file {path}")
    result = model.load_state_dict(adapt_swin_state_dict(model, ckpt), strict=strict)
    if result.missing_keys or result.unexpected_keys:
        print("The model and loaded state dict do not match exactly\n"
              f"unexpected key in source state_dict: {', '.join(result.unexpected_keys)}\n"
              f"missing keys in source state_dict: {', '.join(result.missing_keys)}")
    return ckpt


class SwinBackbone(nn.Module):
    """`Backbone` of models/swin/backbone.py:63-169 for swin_* names: the last `num_feature_levels` of the stride
    8 / 16 / 32 stages with their interpolated padding masks; absolute / relative position tables and every LayerNorm
    frozen (:68-70)."""

    def __init__(self, name: str, num_feature_levels: int = 3, pretrained: str = "", use_checkpoint=False,
                 drop_path_rate: float = 0.2, dilation=False):
        super().__init__()
        assert "swin" in name
        cfg = swin_variant(name)
        body = SwinTransformer(out_indices=[1, 2, 3][-num_feature_levels:], use_checkpoint=use_checkpoint,
                               drop_path_rate=drop_path_rate, **cfg)
        for pname, p in body.named_parameters():
            if "absolute_pos_embed" in pname or "relative_position_bias_table" in pname or "norm" in pname:
                p.requires_grad_(False)
        if pretrained:
            body.init_weights(pretrained)
        self.body = body
        self.strides = [8, 16, 32][-num_feature_levels:]
        self.num_channels = [cfg["embed_dim"] * m for m in (2, 4, 8)][-num_feature_levels:]
        if dilation:
            self.strides[-1] //= 2

    def forward(self, tensor_list: NestedTensor, defer_masks: bool = False) -> Dict[str, NestedTensor]:
        xs = self.body(tensor_list.tensors)
        m = tensor_list.mask
        assert m is not None
        out = {}
        for name, x in xs.items():
            mask = None if defer_masks else F.interpolate(m[None].float(), size=x.shape[-2:]).to(torch.bool)[0]
            out[name] = NestedTensor(x, mask)
        return out
