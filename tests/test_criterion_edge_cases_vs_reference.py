"""Edge cases of the criterion / matcher against the reference's OWN classes, live (needs /root/reference): images without
any ground-truth triplet, a batch with no triplet at all, triplets without an object box (all-zero box: the `exist` masks of
/root/reference/models/hoi.py:4162-4193 and matcher.py:160-179), more triplets than queries (scipy then solves the
untransposed problem), for the fine-tune and the pre-train (pseudo relation label) flag sets.  Same random predictions on
both sides; every loss within 1e-5 relative, matcher indices exact."""
import pytest
import torch

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present")


def _reference_criterion(extra_flags, n_obj, n_verb, nq2):
    ref_import.install()
    flags = [f for f in ref_import.PARSEDA_FLAGS]
    if "--cross_modal_pretrain" in extra_flags:
        flags = [f for f in flags if f != "--hoi"]
    args = ref_import.parse_args(flags + list(extra_flags) + ["--num_queries", str(nq2)])
    with ref_import.chdir(ref_import.REF):
        from models.hoi import SetCriterionHOI
        from models.matcher import build_matcher
        from rlipv2_b200.models import build_weight_dict
        crit = SetCriterionHOI(n_obj, nq2, n_verb, matcher=build_matcher(args), weight_dict=build_weight_dict(args),
                               eos_coef=args.eos_coef, losses=["obj_labels", "verb_labels", "sub_obj_boxes", "obj_cardinality"],
                               verb_loss_type=args.verb_loss_type, obj_loss_type=args.obj_loss_type,
                               matching_symmetric=args.matching_symmetric, RLIP_ParSe=args.RLIP_ParSe,
                               subject_class=args.subject_class, use_no_verb_token=args.use_no_verb_token,
                               giou_verb_label=args.giou_verb_label, verb_curing=args.verb_curing, pseudo_verb=args.pseudo_verb,
                               triplet_filtering=args.triplet_filtering, naive_obj_smooth=args.naive_obj_smooth,
                               naive_verb_smooth=args.naive_verb_smooth, args=args)
    return crit.eval()


@pytest.mark.parametrize("sizes", [(3, 0, 2), (0, 0, 0), (1, 14, 0)])
@pytest.mark.parametrize("extra", [(), ("--cross_modal_pretrain", "--pseudo_verb")])
def test_losses_and_indices_match_reference_on_edge_cases(sizes, extra):
    from tests.test_criterion_stacked import _setup
    nq = 10
    flags = dict(hoi=False, cross_modal_pretrain=True, pseudo_verb=True) if extra else {}
    mine, outputs, targets = _setup(seed=sum(sizes) + len(extra), bs=len(sizes), nq=nq, sizes=sizes, **flags)
    mine.eval()
    ref = _reference_criterion(extra, 7, 5, 2 * nq)
    det = lambda d: {k: (v.detach() if torch.is_tensor(v) else v) for k, v in d.items() if k != "aux_outputs"}
    ref_out = det(outputs)
    ref_out["aux_outputs"] = [det(a) for a in outputs["aux_outputs"]]
    ref_targets = [{k: v.clone() for k, v in t.items()} for t in targets]
    want = ref(ref_out, ref_targets)
    got = mine(outputs, targets)
    assert sorted(got.keys()) == sorted(want.keys())
    for k in want:
        torch.testing.assert_close(got[k].detach().reshape(()), want[k].detach().reshape(()).to(got[k].dtype), rtol=1e-5,
                                   atol=1e-6, msg=lambda m, k=k: f"{k}: {m}")
    last = {k: v for k, v in ref_out.items() if k != "aux_outputs"}
    for (i, j), (ri, rj) in zip(mine.matcher(det(outputs), targets), ref.matcher(last, ref_targets)):
        assert torch.equal(i, ri) and torch.equal(j, rj)
