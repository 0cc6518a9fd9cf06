"""Count aten ops per module during one ParSeDA forward (+ criterion) on CPU - tells where the small kernels
of the train step come from (every aten op is ~one kernel launch on the GPU).
    python tools/op_census.py [depth]"""
import collections
import os
import sys

import torch
from torch.utils._python_dispatch import TorchDispatchMode

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.msda_torch_oracle import CPUFunctionStub  # noqa: E402
import rlipv2_b200.ms_deform_attn as msda_mod  # noqa: E402
from rlipv2_b200 import models, train_step  # noqa: E402
from rlipv2_b200.nested import NestedTensor  # noqa: E402

msda_mod.MSDeformAttnFunction = CPUFunctionStub
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 3
stack = ["<top>"]
counts = collections.Counter()
ops = collections.defaultdict(collections.Counter)
VIEW = ("view", "reshape", "transpose", "permute", "expand", "slice", "select", "unsqueeze", "squeeze", "t.default",
        "alias", "detach", "as_strided", "unbind", "split", "_unsafe_view", "unflatten", "chunk", "narrow", "size", "stride",
        "is_", "sym_", "empty", "lift_fresh", "_local_scalar")


class Census(TorchDispatchMode):
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = str(func)
        if not any(v in name for v in VIEW):
            counts[stack[-1]] += 1
            ops[stack[-1]][name.replace("aten.", "")] += 1
        return func(*args, **(kwargs or {}))


args = models.default_args(device="cpu", num_queries=300, synthetic_text_encoder=True)
torch.manual_seed(0)
model, criterion, _ = models.build_model(args)
model.train()
for name, mod in model.named_modules():
    key = ".".join(name.split(".")[:depth]) or "<model>"
    mod.register_forward_pre_hook(lambda m, i, key=key: (stack.append(key), None)[1])
    mod.register_forward_hook(lambda m, i, o: (stack.pop(), None)[1])
text = train_step.synthetic_text(170, 85)
images, targets = train_step.synthetic_batch(2, 128, 160, seed=0, pin=False)
samples = NestedTensor(images, torch.zeros(2, 128, 160, dtype=torch.bool))
with Census():
    cache = model(samples, encode_and_save=True, text=text, targets=targets)
    out = model(samples, encode_and_save=False, memory_cache=cache, text=text, targets=targets)
    fwd = sum(counts.values())
    stack.append("<criterion>")
    loss_dict = criterion(out, targets)
    total = sum(loss_dict[k] * criterion.weight_dict[k] for k in loss_dict if k in criterion.weight_dict)
    stack.pop()
    stack.append("<backward>")
    total.backward()
print("forward ops", fwd, "total", sum(counts.values()))
for k, v in counts.most_common(40):
    top = ", ".join(f"{o}x{n}" for o, n in ops[k].most_common(8))
    print(f"{v:6d}  {k:55s} {top}")
