"""Diagnose graph-vs-eager differences of ParSeDAInference (tests/test_zz3_infer_gpu.py): raw model outputs of the eager
call, a second eager call, and the graph replay on the same inputs; per-key max |diff| and the object-logit margins of the
queries whose argmax differs.     python tools/debug_infer_graph.py [fp32|tf32]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    from rlipv2_b200 import dense
    from rlipv2_b200.infer_step import ParSeDAInference
    from rlipv2_b200.nested import nested_tensor_from_tensor_list
    from tests.test_infer_step import OBJ, VERB, _build
    model, post = _build("cuda")
    if len(sys.argv) > 1:
        dense.set_matmul_precision(sys.argv[1])
    g = torch.Generator().manual_seed(9)
    imgs = [torch.randn(3, 160, 192, generator=g).cuda(), torch.randn(3, 160, 192, generator=g).cuda()]
    infer = ParSeDAInference(model, post, OBJ, VERB, batch_size=2)
    samples = nested_tensor_from_tensor_list(imgs)
    with torch.no_grad():
        e1 = {k: v.clone() for k, v in infer._forward(samples).items() if torch.is_tensor(v)}
        e2 = {k: v.clone() for k, v in infer._forward(samples).items() if torch.is_tensor(v)}
    infer.capture(160, 192)
    infer.s_samples.tensors.copy_(samples.tensors)
    infer.s_samples.mask.copy_(samples.mask)
    infer.graph.replay()
    torch.cuda.synchronize()
    gr = {k: v.clone() for k, v in infer.s_out.items()}
    infer.graph.replay()
    torch.cuda.synchronize()
    gr2 = {k: v.clone() for k, v in infer.s_out.items()}
    for k in e1:
        print(f"{k:24s} eager-eager {float((e1[k] - e2[k]).abs().max()):.3e}  graph-eager {float((gr[k] - e1[k]).abs().max()):.3e}"
              f"  graph-graph {float((gr[k] - gr2[k]).abs().max()):.3e}  scale {float(e1[k].abs().max()):.3e}")
    lo_e, lo_g = e1["pred_obj_logits"], gr["pred_obj_logits"]
    ae, ag = lo_e[..., :-1].argmax(-1), lo_g[..., :-1].argmax(-1)
    for b, q in (ae != ag).nonzero().tolist():
        te = lo_e[b, q, :-1].topk(2)
        tg = lo_g[b, q, :-1].topk(2)
        print(f"flip b={b} q={q}: eager top2 {te.values.tolist()} idx {te.indices.tolist()} | graph {tg.values.tolist()} idx {tg.indices.tolist()}")
    top2 = lo_e[..., :-1].topk(2).values
    print("smallest eager top-2 margins:", sorted((top2[..., 0] - top2[..., 1]).flatten().tolist())[:6])


if __name__ == "__main__":
    main()
