"""The reference's own ``validate()`` loop (mainExplain_gat.py:675-945, UNMODIFIED, imported from /root/reference over
the shims) driving this repo's ``PipelineModel`` surface and ``SceneGraphBatch`` -- the "drops in unchanged" claim of the
boundary, exercised where the reference tree exists (build container; no GPU there).

What is real here: the reference's loop (``model.eval()``, ``torch.no_grad()``, ``datum.to(device=cuda,
non_blocking=True)`` on the graph batch, ``model(questions, gt_scene_graphs, None, None, SAMPLE_FLAG=True)``, its
``accuracy`` / ``program_string_exact_match_acc`` on the returned tensors, its meters), this repo's ``PipelineModel``
class (constructor, text stack, greedy ``sample``, ``state_dict``) and this repo's collate output.  What is swapped: the
CUDA graph side -- there is no GPU in the container, and the product has no CPU fallback -- is evaluated by the CPU
oracle with the SAME parameters (test-only subclass below), so the logits the loop scores are the oracle's.  The GPU
parity of that graph side is covered by tests/test_graph_side_gpu.py / test_pipeline_gpu.py."""
import argparse
import sys
import types

import pytest
import torch

from oracle import graphvqa_oracle as orc
from oracle import run_reference as rr

pytestmark = pytest.mark.skipif(not rr.available(), reason="reference tree not present (GPU box)")


def _reference_main():
    if "util.misc" not in sys.modules:      # DETR's util/misc.py does not import with current torchvision; validate()
        util, misc = types.ModuleType("util"), types.ModuleType("util.misc")        # only asks is_main_process()
        misc.is_main_process = lambda: True
        util.misc = misc
        sys.modules["util"], sys.modules["util.misc"] = util, misc
    main = rr.load("mainExplain_gat")
    main.GQATorchDataset.indices_to_string = staticmethod(
        lambda idx, flag=True: (" ".join(str(int(t)) for t in idx if int(t) > 3), None))
    return main


def _oracle_backed(model_cls):
    class OracleBackedPipeline(model_cls):
        calls = 0

        def _run(self, questions, gt_scene_graphs, programs_input, mode):
            type(self).calls += 1
            assert mode == "sample" and programs_input is None
            q_enc, programs_output, instr = self._text_side(questions, programs_input, mode)
            side = orc.GraphSide(self.vocab.sg_vocab_size, self.vocab.sg_pad_idx).eval()
            own = self.state_dict()
            side.load_state_dict({k: own[k] for k in side.state_dict()})
            return programs_output, side(gt_scene_graphs, instr, q_enc[0])
    return OracleBackedPipeline


def test_reference_validate_loop_drives_the_engine_surface(golden, capsys):
    from graphvqa_b200.collate import SceneGraphVocab, collate_scene_graphs
    from graphvqa_b200.pipeline_model_gat import PipelineModel, VocabSpec
    main = _reference_main()
    fx = golden("collate_debug")
    vocab = SceneGraphVocab(fx["itos"])
    sgs = list(fx["scene_graphs"].values())
    # 4 debug graphs, one of them twice, one empty: the loop's group meter divides by batch // 5 (mainExplain_gat.py
    # :790), so a batch below 5 questions makes the reference itself divide by zero
    graphs = collate_scene_graphs(sgs + [sgs[1], {"objects": {}}], vocab, pin_memory=False)
    b = graphs.num_graphs
    assert b == 6
    torch.manual_seed(0)
    model = _oracle_backed(PipelineModel)(VocabSpec(text_vocab_size=rr.TEXT_VOCAB_SIZE, sg_vocab_size=len(vocab)))
    g = torch.Generator().manual_seed(1)
    questions = torch.randint(4, rr.TEXT_VOCAB_SIZE, (9, b), generator=g)
    programs = torch.randint(4, rr.TEXT_VOCAB_SIZE, (8, b * 5), generator=g)
    programs[0] = 2
    programs[5:] = 1                                        # padded tail
    full_answers = torch.randint(4, rr.TEXT_VOCAB_SIZE, (6, b), generator=g)
    labels = torch.randint(0, 1842, (b,), generator=g)
    batch = (["q%d" % i for i in range(b)], questions, graphs, programs, full_answers, labels, ["t"] * b)
    args = argparse.Namespace(print_freq=1, output_dir=".")
    main.validate([batch, batch], model, None, args)         # two batches through the unmodified loop
    out = capsys.readouterr().out
    assert type(model).calls == 2 and not model.training
    assert "Acc@Short" in out and "Acc@Program" in out
    # the same call outside the loop: shapes / dtypes the loop relies on
    with torch.no_grad():
        pred, logits = model(questions, graphs.to(device=main.cuda, non_blocking=True), None, None, SAMPLE_FLAG=True)
    assert pred.dtype == torch.long and pred.shape == (16, b * 5)
    assert logits.shape == (b, 1842) and torch.isfinite(logits).all()
    acc1 = main.accuracy(logits, labels, topk=(1,))[0]
    assert 0.0 <= float(acc1) <= 100.0
