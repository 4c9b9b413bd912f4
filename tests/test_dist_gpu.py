"""Multi-GPU engine test (needs >= 2 GPUs, skipped otherwise): PipelineModel.answer_logits sharded by
graph range over 2 ranks (NCCL all-gather of the logits) equals the single-GPU result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from graphvqa_b200.dist import distributed_answer_logits, init_distributed
    from graphvqa_b200.graph_batch import SceneGraphBatch, synthetic_topology
    from graphvqa_b200.pipeline_model_gat import PipelineModel, VocabSpec
    from oracle.golden_utils import deterministic_fill
    init_distributed("nccl")
    try:
        dev = torch.device("cuda", rank)
        model = deterministic_fill(PipelineModel(VocabSpec(text_vocab_size=500, sg_vocab_size=400)).eval(), 3).to(dev)
        b = 10
        ei, batch, _ = synthetic_topology(b, 12, 30, seed=4, jitter=4)
        g = torch.Generator().manual_seed(5)
        graphs = SceneGraphBatch(x=torch.randint(2, 400, (batch.numel(), 12), generator=g), edge_index=ei,
                                 edge_attr=torch.randint(2, 400, (ei.size(1), 1), generator=g), batch=batch,
                                 added_sym_edge=torch.randint(0, ei.size(1) // 3, (2 * b,), generator=g),
                                 num_graphs=b).to(device=dev)
        questions = torch.randint(4, 500, (7, b), generator=g).to(dev)
        with torch.no_grad():
            gathered = distributed_answer_logits(model, questions, graphs)
            single = model.answer_logits(questions, graphs)
        ret[rank] = float((gathered - single).abs().max())
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_answer_logits_equal_single_gpu():
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert all(ret[r] <= 2e-5 for r in range(2)), dict(ret)
