"""The task-level mirror (`ultra_torchdrug_b200/task.py`) against the reference's own functions: strict negative
sampling and the filter masks from golden vectors made by the unmodified reference ultra/task.py:65-118
(tests/golden/make_task_golden.py), the loss against a literal restatement, sharded steps over gloo (world size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ultra_torchdrug_b200 import nbf, sharding, synthetic, task
from ultra_torchdrug_b200.compat.torchdrug import data

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "task_strict_negative.npz")


def _golden_graph():
    g = np.load(GOLDEN)
    num_node, num_relation, num_negative, seed = (int(x) for x in g["shape"])
    graph = data.Graph(torch.from_numpy(g["triples"]), num_node=num_node, num_relation=num_relation)
    return g, graph, num_negative, seed


def test_strict_negative_equals_reference_function():
    g, graph, num_negative, seed = _golden_graph()
    batch = torch.from_numpy(g["batch"])
    sampler = task.StrictNegativeSampler(graph)
    torch.manual_seed(seed)                                  # the reference draws torch.rand for the tails, then the heads
    negative = sampler(batch[:, 0], batch[:, 1], batch[:, 2], num_negative)
    assert torch.equal(negative, torch.from_numpy(g["negative"]))
    # and never a known answer, whatever the uniform numbers (0 and the largest float below 1 included)
    half = len(batch) // 2
    edge = torch.tensor([[0.0, 1.0 - 2 ** -24]]).expand(half, 2)
    extreme = sampler(batch[:, 0], batch[:, 1], batch[:, 2], 2, rand=(edge, edge))
    t_mask, h_mask = torch.from_numpy(g["t_mask"]), torch.from_numpy(g["h_mask"])
    assert t_mask[:half].gather(1, extreme[:half]).all() and h_mask[half:].gather(1, extreme[half:]).all()
    assert t_mask[:half].gather(1, negative[:half]).all() and h_mask[half:].gather(1, negative[half:]).all()


def test_filter_mask_equals_reference_masks():
    g, graph, _, _ = _golden_graph()
    batch = torch.from_numpy(g["batch"])
    model, rel_model = nbf.ultra_models(graph.num_relation, hidden=8, num_layers=1)
    evaluator = task.ShardedEvaluator(nbf.UltraRanker(model, rel_model, graph))
    mask = evaluator.filter_mask(batch)
    assert torch.equal(mask[:, 0], torch.from_numpy(g["t_mask"])) and torch.equal(mask[:, 1], torch.from_numpy(g["h_mask"]))
    assert torch.equal(mask, nbf.UltraRanker(model, rel_model, graph).filter_mask(batch))      # the graph.match route


def test_training_loss_is_the_reference_formula():
    torch.manual_seed(0)
    pred = torch.randn(6, 9)
    target = torch.zeros_like(pred)
    target[:, 0] = 1
    loss = torch.nn.functional.binary_cross_entropy_with_logits(pred, target, reduction="none")
    weight = torch.ones_like(pred)
    weight[:, 1:] = torch.softmax(pred[:, 1:] / 0.5, dim=-1)
    want = (loss * weight).sum(-1) / weight.sum(-1)
    torch.testing.assert_close(task.training_loss(pred, 0.5), want)
    weight[:, 1:] = 1 / 8
    torch.testing.assert_close(task.training_loss(pred, 0), (loss * weight).sum(-1) / weight.sum(-1))


def test_training_indices_layout():
    batch = torch.tensor([[1, 2, 0], [3, 4, 1], [5, 6, 0], [7, 8, 1]])
    negative = torch.arange(100, 112).view(4, 3)
    h, t, r = task.training_indices(batch, negative)
    assert h[:, 0].tolist() == [1, 3, 5, 7] and t[:, 0].tolist() == [2, 4, 6, 8] and (r == batch[:, 2:]).all()
    assert torch.equal(t[:2, 1:], negative[:2]) and (h[:2] == batch[:2, :1]).all()          # first half: corrupted tails
    assert torch.equal(h[2:, 1:], negative[2:]) and (t[2:] == batch[2:, 1:2]).all()          # second half: corrupted heads


# ---- world size 2 over gloo: the sharded fine-tuning step and the sharded evaluation ----------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(seed=3, patch=True):
    from oracle.rspmm_oracle import generalized_rspmm_oracle
    if patch:
        nbf.generalized_rspmm = generalized_rspmm_oracle        # CPU worker processes: the layers call the oracle operator
    num_node, num_relation = 40, 3
    triples = synthetic.triples(num_node, num_relation, 260, seed=seed)
    graph = data.Graph(triples, num_node=num_node, num_relation=num_relation)
    torch.manual_seed(seed)
    model, rel_model = nbf.ultra_models(num_relation, hidden=8, num_layers=2)
    return triples, graph, model, rel_model


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        triples, graph, model, rel_model = _setup()
        batch = triples[:6]
        # evaluation: ranks of the global batch from per-rank slabs == single-process ranks
        evaluator = task.ShardedEvaluator(nbf.UltraRanker(model.eval(), rel_model.eval(), graph), rank=rank, world_size=world)
        ranks = evaluator(batch)
        # training: the all-reduced gradients are the gradients of the global-batch mean loss
        step = task.FinetuneStep(model.train(), rel_model.train(), graph, num_negative=4, rank=rank, world_size=world)
        torch.manual_seed(100 + rank)
        step(batch)
        results[rank] = {"ranks": ranks.clone(), "grads": [None if p.grad is None else p.grad.clone() for p in step.parameters],
                         "weights": [p.detach().clone() for p in step.parameters]}
    finally:
        dist.destroy_process_group()


def test_sharded_evaluation_and_finetune_step_gloo(monkeypatch):
    from oracle.rspmm_oracle import generalized_rspmm_oracle
    monkeypatch.setattr(nbf, "generalized_rspmm", generalized_rspmm_oracle)     # this process: restored after the test
    world = 2
    manager = mp.get_context("spawn").Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    triples, graph, model, rel_model = _setup(patch=False)
    batch = triples[:6]
    single = task.ShardedEvaluator(nbf.UltraRanker(model.eval(), rel_model.eval(), graph))(batch)
    assert single.shape == (6, 2) and (single >= 1).all()
    for rank in range(world):
        assert torch.equal(results[rank]["ranks"], single), "sharded ranks differ from the single-process ranks"
    # both ranks end the step with identical gradients and weights (DDP invariant)
    for a, b in zip(results[0]["grads"], results[1]["grads"]):
        assert (a is None) == (b is None) and (a is None or torch.equal(a, b))
    for a, b in zip(results[0]["weights"], results[1]["weights"]):
        assert torch.equal(a, b)
    assert any(g is not None and g.abs().sum() > 0 for g in results[0]["grads"])


@pytest.mark.gpu
def test_captured_finetune_step_equals_eager(cuda, monkeypatch):
    """`FinetuneStep.capture()`: the whole step (strict negatives, forward, backward, AdamW) replayed as one CUDA graph gives
    the weights of the eager step, bit for bit - every kernel on the path is deterministic - when both draw the same
    uniform numbers (torch.rand is pinned to a constant here; the graph-safe generator would otherwise differ)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    monkeypatch.setattr(torch, "rand", lambda *shape, device=None, **unused: torch.full(shape, 0.37, device=device))
    num_node, num_relation = 300, 5
    triples = synthetic.triples(num_node, num_relation, 2500, seed=9)
    graph = data.Graph(triples, num_node=num_node, num_relation=num_relation).to(cuda)
    batches = [triples[16 * i:16 * i + 16].to(cuda) for i in range(4)]

    def build():
        torch.manual_seed(5)
        model, rel_model = nbf.ultra_models(num_relation, hidden=64, num_layers=2)
        return task.FinetuneStep(model.to(cuda).train(), rel_model.to(cuda).train(), graph, num_negative=8)

    eager, captured = build(), build().capture(16)
    assert captured._graph is not None
    for a, b in zip(eager.parameters, captured.parameters):
        assert torch.equal(a, b), "capture() must leave the weights as they were"
    for batch in batches:
        loss_eager = eager(batch)
        loss_captured = captured(batch).clone()
        torch.testing.assert_close(loss_captured, loss_eager, rtol=1e-6, atol=1e-7)
    for a, b in zip(eager.parameters, captured.parameters):
        torch.testing.assert_close(b, a, rtol=1e-6, atol=1e-7)
    assert torch.isfinite(loss_captured)


@pytest.mark.gpu
def test_single_node_layers_give_the_gradients_of_the_three_node_layers(cuda, monkeypatch):
    """Model level: the fine-tuning loss and every parameter gradient with each NBFNet layer as one autograd node
    (`functional.nbf_layer`, the default) against the same step with the three separate nodes (ULTRA_NBF_SINGLE_NODE=0):
    the same arithmetic up to summation order (row statistics of the fused forward kernel, the three contributions to a
    layer input's gradient)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    monkeypatch.setattr(torch, "rand", lambda *shape, device=None, **unused: torch.full(shape, 0.61, device=device))
    num_node, num_relation = 400, 6
    triples = synthetic.triples(num_node, num_relation, 3000, seed=11)
    graph = data.Graph(triples, num_node=num_node, num_relation=num_relation).to(cuda)
    batch = triples[:12].to(cuda)

    def gradients(single):
        monkeypatch.setenv("ULTRA_NBF_SINGLE_NODE", "1" if single else "0")
        torch.manual_seed(3)
        model, rel_model = nbf.ultra_models(num_relation, hidden=64, num_layers=3)
        step = task.FinetuneStep(model.to(cuda).train(), rel_model.to(cuda).train(), graph, num_negative=6)
        loss = step.loss(batch)
        loss.backward()
        return loss.detach(), [None if p.grad is None else p.grad.clone() for p in step.parameters]

    loss_single, grads_single = gradients(True)
    loss_three, grads_three = gradients(False)
    torch.testing.assert_close(loss_single, loss_three, rtol=1e-5, atol=1e-7)
    assert sum(g is not None for g in grads_single) == sum(g is not None for g in grads_three) > 0
    for a, b in zip(grads_single, grads_three):
        if b is None:
            assert a is None or float(a.abs().max()) == 0.0
            continue
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()) + 1e-9)

