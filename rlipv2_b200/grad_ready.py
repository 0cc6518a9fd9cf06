"""Markers that tell the data-parallel step WHEN a range of parameter gradients is final during the backward.

The reference's DistributedDataParallel (main.py:515-517) overlaps its bucketed gradient all-reduce with the backward
through per-parameter hooks.  The graphed step of this repo keeps all gradients in one flat buffer whose producing kernels
write into it directly (dense.py), so there are no AccumulateGrad hooks to hang buckets on.  Instead a few identity nodes
are placed on activations in the forward; autograd runs a node's backward only after every consumer of its output has
run, and the engine pops ready nodes in decreasing creation order, so when the marker on an activation fires, the
gradients of every parameter used downstream of it (and created after it) are complete:

  'image<i>'   the backbone's feature maps (parseda.py)      -> with 'text': everything but backbone and text tower
  'text'       the pooled label embeddings (parseda_transformer.py)
  'text_mid'   input of text-tower layer `split`             -> tower layers split.. + pooler
  'text_emb'   input of text-tower layer 0                   -> tower layers 0..split-1

`flat_dp.EarlyReducer` turns fired tags into all-reduces of the corresponding flat-gradient ranges on a communication
stream.  With no callback installed `mark` returns its argument untouched (single GPU, eval, tests).
"""
import torch

_callback = None
_applied = None


def set_callback(cb):
    """cb(tag) is called from the backward pass; None uninstalls"""
    global _callback, _applied
    _callback = cb
    _applied = set() if cb is not None else None


def applied_tags():
    """tags marked during the forward passes since `set_callback` / `reset_applied`"""
    return set(_applied or ())


def reset_applied():
    if _applied is not None:
        _applied.clear()


class _Marker(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, tag):
        ctx.tag = tag
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad):
        cb = _callback
        if cb is not None:
            cb(ctx.tag)
        return grad, None


def mark(x, tag):
    if _callback is None or not (torch.is_grad_enabled() and x.requires_grad):
        return x
    _applied.add(tag)
    return _Marker.apply(x, tag)


def install_text_tower_markers(text_encoder, split=6):
    """forward pre-hooks on the HF tower's layers 0 and `split` that pass the incoming hidden states through `mark`;
    -> the hook handles (remove() uninstalls)"""
    layers = text_encoder.encoder.layer

    def hook(tag):
        def pre(module, args, kwargs):
            if args:
                return (mark(args[0], tag),) + tuple(args[1:]), kwargs
            if "hidden_states" in kwargs:
                kwargs = dict(kwargs)
                kwargs["hidden_states"] = mark(kwargs["hidden_states"], tag)
            return args, kwargs
        return pre

    handles = [layers[0].register_forward_pre_hook(hook("text_emb"), with_kwargs=True)]
    if 0 < split < len(layers):
        handles.append(layers[split].register_forward_pre_hook(hook("text_mid"), with_kwargs=True))
    return handles
