"""Inference step: the per-batch body of the reference's evaluation loop as one call.

/root/reference/engine.py:360-430 (`evaluate_hoi_with_text`): the label strings of the dataset are tokenised and encoded
ONCE (`:367-391`), every batch then runs phase A + phase B with the pre-encoded `text` tuple (`:409-421`), drops the
trailing "no verb" column if the model has one (`:425-426`) and post-processes (`:427`).  `ParSeDAInference` holds the
encoded label set and does exactly that; `capture()` additionally records the two model phases for a fixed padded
image shape into one CUDA graph (the eager forward is ~2 700 launches of mostly 5-8 us kernels - launch-bound at batch
1), leaving only the input copy, the replay and the post-processor's single device->host read per batch.
"""
import torch

from . import streams
from .nested import NestedTensor, nested_tensor_from_tensor_list


class ParSeDAInference:
    def __init__(self, model, postprocessor, object_text, verb_text, batch_size, use_no_obj_token=True, device=None,
                 fusion_type="GLIP_attn"):
        self.model = model.eval()
        self.post = postprocessor
        self.batch_size = int(batch_size)
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.n_verbs = len(verb_text)
        tr = model.transformer
        flat = list(object_text) + (["no objects"] if use_no_obj_token else []) + list(verb_text)
        sums = torch.tensor([[len(object_text) + int(bool(use_no_obj_token)), len(verb_text)]])
        with torch.no_grad():                                                      # engine.py:373-378
            tok = tr.tokenizer.batch_encode_plus(flat, padding="longest", return_tensors="pt").to(self.device)
            memory = tr.text_encoder(**tok).pooler_output
            if fusion_type != "GLIP_attn":                                          # engine.py:382-385
                memory = tr.resizer(memory)
        self.text_memory = memory.unsqueeze(1).repeat(1, self.batch_size, 1)       # [n_text, bs, C]
        self.text_mask = torch.zeros(self.text_memory.shape[:2], dtype=torch.bool, device=self.device)
        self.sums = sums
        self.graph = None

    def _text(self, bs):
        if bs == self.batch_size:
            return (self.text_mask, self.text_memory, self.sums)
        return (self.text_mask[:, :bs], self.text_memory[:, :bs], self.sums)      # short last batch, engine.py:415-419

    def _forward(self, samples):
        text = self._text(samples.tensors.shape[0])
        cache = self.model(samples, encode_and_save=True, text=text)
        out = self.model(samples, encode_and_save=False, memory_cache=cache, text=text)
        if out["pred_verb_logits"].shape[2] == self.n_verbs + 1:                   # engine.py:425-426
            out["pred_verb_logits"] = out["pred_verb_logits"][:, :, :-1]
        return out

    @torch.no_grad()
    def __call__(self, samples, orig_target_sizes):
        """samples: NestedTensor or list of [3, H, W] images; orig_target_sizes [bs, 2] (h, w) -> per-image result dicts"""
        if not isinstance(samples, NestedTensor):
            samples = nested_tensor_from_tensor_list([s.to(self.device) for s in samples])
        if (self.graph is not None and tuple(samples.tensors.shape) == tuple(self.s_samples.tensors.shape)):
            self.s_samples.tensors.copy_(samples.tensors, non_blocking=True)
            self.s_samples.mask.copy_(samples.mask, non_blocking=True)
            self.graph.replay()
            out = self.s_out
        else:
            out = self._forward(samples)
        return self.post(out, orig_target_sizes)

    @torch.no_grad()
    def capture(self, height, width, batch=None, warmup=2):
        """record phase A + phase B for [batch, 3, height, width] inputs (any padding mask) into one CUDA graph"""
        assert self.device.type == "cuda", "CUDA graphs need a GPU (there is no CPU fallback)"
        bs = self.batch_size if batch is None else int(batch)
        self.s_samples = NestedTensor(torch.zeros(bs, 3, height, width, device=self.device),
                                      torch.zeros(bs, height, width, dtype=torch.bool, device=self.device))
        side = streams.get(self.device, "capture")
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._forward(self.s_samples)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            out = self._forward(self.s_samples)
        self.s_out = {k: v for k, v in out.items() if torch.is_tensor(v)}
        return self
