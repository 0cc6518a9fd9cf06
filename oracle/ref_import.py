"""TEST INFRASTRUCTURE - import the reference's own Python modules from /root/reference (read-only).

Only usable in the build container (the GPU box has no /root/reference).  It installs the shim set
of SURVEY.md Appendix A so that the reference's `models.*` import and run on CPU under
torch 2.11 / transformers 5.5 without network access:

  1. a stub module named `MultiScaleDeformableAttention` (ms_deform_attn_func.py:22 imports it);
  2. transformers 4.5-era helper names expected by models/modeling_roberta.py:24,26;
  3. a `timm` stub (models/fuse_helper.py:13 and models/swin/swin_transformer.py:21 import DropPath, to_2tuple,
     trunc_normal_);
  4. offline RoBERTa: `from_pretrained` of tokenizer/model/config return a random-init roberta-base
     shaped model and a deterministic fake tokenizer (dab_deformable/deformable_transformer.py:296,334-335);
  5. `torch.load` of the hard-coded ResNet-50 path returns torchvision's random state_dict
     (models/DDETR_backbone.py:121);
  6. RobertaLayer.get_extended_attention_mask with transformers-4.5.1 semantics
     (modeling_roberta.py:380 passes `device` where transformers 5 expects `dtype`);
  7. the reference's CPU route for the op: MSDeformAttnFunction.apply ->
     ms_deform_attn_core_pytorch (the native CPU path raises, cpu/ms_deform_attn_cpu.cpp:17-41);
  8. cwd = /root/reference while constructing the criterion (datasets/priors/*.npz, hoi.py:3678).

Nothing here is product code; it exists to generate the fixtures under tests/golden/.
"""
import contextlib
import os
import sys
import types

import torch

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "models"))


class FakeTokenizer:
    """Deterministic stand-in for RobertaTokenizerFast: ids from a string hash, <s>=0 </s>=2 pad=1.
    The same class is used by the product's offline text path (rlipv2_b200.text_encoder), so the
    reference and the product see identical token ids."""

    def batch_encode_plus(self, texts, padding="longest", return_tensors="pt"):
        from rlipv2_b200.text_encoder import hash_tokenize
        from transformers import BatchEncoding
        ids, mask = hash_tokenize(texts)
        return BatchEncoding({"input_ids": ids, "attention_mask": mask})

    @classmethod
    def from_pretrained(cls, *a, **k):
        return cls()


@contextlib.contextmanager
def chdir(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


_installed = False


def install():
    global _installed
    if _installed:
        return
    assert available(), "/root/reference is not present (this only runs in the build container)"
    # 1. stub extension module
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    # 2. transformers helper names
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    mu.prune_linear_layer = pu.prune_linear_layer
    mu.find_pruneable_heads_and_indices = getattr(pu, "find_pruneable_heads_and_indices", lambda *a, **k: None)
    # 3. timm stub
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        timm_models = types.ModuleType("timm.models")
        timm_layers = types.ModuleType("timm.models.layers")

        class DropPath(torch.nn.Module):
            """timm is absent: per-sample stochastic depth with timm's published semantics (identity in eval mode,
            which is all the fixtures use; Swin backbones are built with drop_path_rate > 0)"""

            def __init__(self, p=0.0):
                super().__init__()
                self.drop_prob = float(p)

            def forward(self, x):
                if self.drop_prob == 0.0 or not self.training:
                    return x
                keep = 1.0 - self.drop_prob
                m = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
                return x * (m / keep if keep > 0 else m)

        timm_layers.DropPath = DropPath
        timm_layers.to_2tuple = lambda x: (x, x)
        timm_layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models = timm_models
        timm_models.layers = timm_layers
        sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.layers": timm_layers})
    class _Permissive(types.ModuleType):      # any attribute resolves to a dummy class
        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            return type(item, (), {})

    for name in ("pycocotools", "pycocotools.mask", "pycocotools.coco", "pycocotools.cocoeval",
                 "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "h5py", "prettytable",
                 "submitit", "panopticapi", "panopticapi.utils", "cv2"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Permissive(name)
    # 4. offline RoBERTa
    import transformers
    from rlipv2_b200.text_encoder import roberta_base_config

    transformers.RobertaTokenizerFast.from_pretrained = classmethod(lambda cls, *a, **k: FakeTokenizer())
    transformers.RobertaConfig.from_pretrained = classmethod(lambda cls, *a, **k: roberta_base_config())
    transformers.RobertaModel.from_pretrained = classmethod(
        lambda cls, *a, **k: transformers.RobertaModel(roberta_base_config()))
    # 5. ResNet-50 weights
    _orig_load = torch.load

    def _load(f, *a, **k):
        if isinstance(f, str) and f.endswith("resnet50-19c8e357.pth"):
            import torchvision
            return torchvision.models.resnet50().state_dict()
        return _orig_load(f, *a, **k)

    torch.load = _load
    if REF not in sys.path:
        sys.path.insert(0, REF)
    # 6. extended attention mask (transformers 4.5.1: (1 - mask) * -10000)
    import models.modeling_roberta as mr
    mr.RobertaLayer.get_extended_attention_mask = \
        lambda self, m, shape, device=None: (1.0 - m[:, None, None, :].float()) * -10000.0
    # 7. CPU route for the op
    import models.ops.functions.ms_deform_attn_func as f1
    import models.dab_deformable.ops.functions.ms_deform_attn_func as f2
    import models.ops.modules.ms_deform_attn as m1
    import models.dab_deformable.ops.modules.ms_deform_attn as m2

    class _CPUFunction:
        @staticmethod
        def apply(value, shapes, level_start, loc, attn, im2col_step):
            return f1.ms_deform_attn_core_pytorch(value, shapes, loc, attn)

    for mod in (f1, f2, m1, m2):
        mod.MSDeformAttnFunction = _CPUFunction
    _installed = True


def parse_args(argv):
    """argparse namespace of the reference's main.py for the given flag list."""
    install()
    with chdir(REF):
        import main as ref_main
        args = ref_main.get_args_parser().parse_args(argv + ["--device", "cpu"])
    args.distributed = False
    return args


# flags of scripts/RLIP_ParSeDA/fine_tune_RLIP_ParSeDA_v2_hico.sh:17-59 that shape the model
PARSEDA_FLAGS = [
    "--hoi", "--backbone", "resnet50", "--load_backbone", "supervised", "--set_cost_bbox", "2.5",
    "--set_cost_giou", "1", "--bbox_loss_coef", "2.5", "--giou_loss_coef", "1", "--obj_loss_type",
    "cross_entropy", "--verb_loss_type", "focal", "--enc_layers", "6", "--dec_layers", "3",
    "--RLIP_ParSeDA_v2", "--with_box_refine", "--num_feature_levels", "4", "--num_patterns", "0",
    "--pe_temperatureH", "20", "--pe_temperatureW", "20", "--dim_feedforward", "2048", "--dropout", "0.0",
    "--use_no_obj_token", "--sampling_stategy", "freq", "--fusion_type", "GLIP_attn",
    "--gating_mechanism", "VXAc", "--verb_query_tgt_type", "vanilla_MBF", "--fusion_interval", "2",
    "--fusion_last_vis", "--lang_aux_loss", "--giou_verb_label", "--subject_class",
]


def build_reference_model(num_queries=16, extra=()):
    """(model, criterion, postprocessors, args) of the reference for the ParSeDA fine-tune flags."""
    install()
    args = parse_args(PARSEDA_FLAGS + ["--num_queries", str(num_queries)] + list(extra))
    with chdir(REF):
        from models import build_model
        model, criterion, post = build_model(args)
    return model, criterion, post, args
