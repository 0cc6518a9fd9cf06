"""One RLIPv2-ParSeDA training step as the reference's engine runs it.

Restates the body of the hot loop of /root/reference/engine.py:68-172 (`train_one_epoch`) and the
optimizer construction of main.py:514-539 as a reusable object:

    H2D of the batch -> model(samples, encode_and_save=True) -> model(..., encode_and_save=False)
    -> criterion -> weighted sum -> zero_grad -> backward -> clip_grad_norm_(0.1) -> AdamW.step

Data-parallel: one process per GPU, the model wrapped in DistributedDataParallel over NCCL
(main.py:515-517).  The reference needs `find_unused_parameters=True` because the verb decoder's
box heads only feed detached anchors (dab_deformable/deformable_transformer.py:1511-1541); the set of
unused parameters is static, so `static_graph=True` gives the same result without the per-step graph
walk.
"""
import torch
import torch.distributed as dist

from . import dense, models
from .nested import NestedTensor


def synthetic_text(n_obj=170, n_verb=85):
    """256 label strings: n_obj object names + 'no objects' + n_verb relation names."""
    objs = [f"object kind {i}" for i in range(n_obj)] + ["no objects"]
    verbs = [f"relation {i} with" for i in range(n_verb)]
    return [(objs, verbs)]


def synthetic_batch(batch, height=800, width=1333, n_obj=170, n_verb=85, triplets=5, seed=0, pin=True):
    """Host-side batch: images [B,3,H,W] fp32 + per-image targets (SURVEY.md section 8d recipe)."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, height, width, generator=g)
    targets = []
    for _ in range(batch):
        def boxes():
            return torch.cat([torch.rand(triplets, 2, generator=g) * 0.4 + 0.3,
                              torch.rand(triplets, 2, generator=g) * 0.2 + 0.1], 1)
        verbs = torch.zeros(triplets, n_verb)
        verbs[torch.arange(triplets), torch.randint(0, n_verb, (triplets,), generator=g)] = 1
        targets.append({"obj_labels": torch.randint(0, n_obj, (triplets,), generator=g),
                        "sub_labels": torch.zeros(triplets, dtype=torch.long), "verb_labels": verbs,
                        "sub_boxes": boxes(), "obj_boxes": boxes()})
    if pin and torch.cuda.is_available():
        images = images.pin_memory()
        targets = [{k: v.pin_memory() for k, v in t.items()} for t in targets]
    return images, targets


class ParSeDATrainStep:
    """model + criterion + optimizer; `step(images_host, targets_host, text)` runs one iteration and
    returns the (device) total loss."""

    def __init__(self, args=None, device="cuda", precision="tf32", seed=0, ddp=None, clip_max_norm=0.1,
                 lr=1.41e-4, lr_backbone=1.41e-5, text_encoder_lr=1.41e-5, weight_decay=1e-4):
        self.device = torch.device(device)
        dense.set_matmul_precision(precision)
        torch.manual_seed(seed)
        if args is None:
            args = models.default_args(device=str(self.device), num_queries=300, synthetic_text_encoder=True)
        self.args = args
        model, criterion, _ = models.build_model(args)
        self.model = model.to(self.device).train()
        self.criterion = criterion.to(self.device).train()
        self.module = self.model
        if ddp is None:
            ddp = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if ddp:
            self.model = torch.nn.parallel.DistributedDataParallel(
                self.model, device_ids=[self.device.index], static_graph=True, gradient_as_bucket_view=True)
        # main.py:523-539: three groups selected by name
        named = list(self.module.named_parameters())
        groups = [
            {"params": [p for n, p in named if "backbone" not in n and "text_encoder" not in n and p.requires_grad]},
            {"params": [p for n, p in named if "backbone" in n and p.requires_grad], "lr": lr_backbone},
            {"params": [p for n, p in named if "text_encoder" in n and p.requires_grad], "lr": text_encoder_lr},
        ]
        self.optimizer = torch.optim.AdamW(groups, lr=lr, weight_decay=weight_decay, fused=self.device.type == "cuda")
        self.clip_max_norm = clip_max_norm
        self.params = [p for g in groups for p in g["params"]]

    def to_device(self, images_host, targets_host):
        images = images_host.to(self.device, non_blocking=True)
        mask = torch.zeros(images.shape[0], images.shape[2], images.shape[3], dtype=torch.bool, device=self.device)
        targets = [{k: v.to(self.device, non_blocking=True) for k, v in t.items()} for t in targets_host]
        return NestedTensor(images, mask), targets

    def step_device(self, samples, targets, text):
        """One optimisation step on a batch that is already resident in HBM."""
        memory_cache = self.model(samples, encode_and_save=True, text=text, targets=targets)
        outputs = self.model(samples, encode_and_save=False, memory_cache=memory_cache, text=text, targets=targets)
        loss_dict = self.criterion(outputs, targets)
        wd = self.criterion.weight_dict
        losses = sum(loss_dict[k] * wd[k] for k in loss_dict.keys() if k in wd)
        self.optimizer.zero_grad(set_to_none=True)
        losses.backward()
        if self.clip_max_norm > 0:
            torch.nn.utils.clip_grad_norm_(self.params, self.clip_max_norm, foreach=True)
        self.optimizer.step()
        return losses.detach()

    def step(self, images_host, targets_host, text):
        samples, targets = self.to_device(images_host, targets_host)
        return self.step_device(samples, targets, text)
