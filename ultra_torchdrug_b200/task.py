"""Host-side mirror of the task-level callers either side of the rspmm hot path.

What the multi-GPU BASELINE configurations need around the operator, restated with the reference's semantics:

  StrictNegativeSampler     reference ultra/task.py:65-118  (`_calculate_t_mask` / `_calculate_h_mask` / `_strict_negative`)
  training_loss             reference ultra/task.py:160-195 (BCE with self-adversarial negative weights)
  FinetuneStep              reference ultra/task.py:264-275 (train branch of `predict`) + ultra/engine.py:71-86
                            (forward, backward, DDP gradient all-reduce, optimizer step)
  ShardedEvaluator          reference ultra/task.py:228-263, 279-315 + ultra/engine.py:130-150 (predict, filtered ranks,
                            gather of the per-query ranks over the ranks of the job)

The query batch is sharded over the GPUs of one box (`sharding.query_slab`): every rank holds the full graph and both
networks; the only collectives are the gradient all-reduce (training) and the gather of per-query ranks (evaluation).
"""
import torch
from torch.nn import functional as F

from . import nbf, sharding


class StrictNegativeSampler(object):
    """Negatives that are not known answers, drawn on the device without materialising the (B, N) candidate masks.

    The reference builds a boolean mask of non-answers per positive (`graph.match` + scatter, task.py:65-100), lists
    the candidates with `nonzero()` (a host synchronisation and B x N bytes) and draws `num_negative` of them per row
    with `variadic_sample` (task.py:102-118).  Drawing index k = floor(rand * #candidates) from the ascending list of
    non-answers is the same as taking entity  k + #{answers a_i : a_i - i <= k}  (a_i ascending): the answers of every
    (h, r) / (t, r) pair are one sorted range of a key array built once per fact graph, and the count is a binary search
    over that range - B x num_negative independent searches, no mask, no synchronisation.  Given the same uniform numbers
    the result equals the reference's (tests/test_task.py, golden from the reference function)."""

    def __init__(self, graph):
        edge = graph.edge_list
        self.num_node, self.num_relation = int(graph.num_node), int(graph.num_relation)
        h, t, r = edge[:, 0], edge[:, 1], edge[:, 2]
        n, big_r = self.num_node, self.num_relation
        self.tail_keys = torch.unique((h * big_r + r) * n + t)          # answers t of (h, r, ?), ascending per (h, r)
        self.head_keys = torch.unique((t * big_r + r) * n + h)          # answers h of (?, r, t)
        longest = 1
        for keys in (self.tail_keys, self.head_keys):                    # one-off: longest answer list -> search depth
            if len(keys):
                longest = max(longest, int(torch.unique_consecutive(keys // n, return_counts=True)[1].max()))
        self.search_steps = max(1, int(longest).bit_length())

    def _draw(self, keys, pair, rand):
        """pair (B,) = h * R + r (or t * R + r); rand (B, S) uniform in [0, 1).  Returns (B, S) entity ids."""
        n = self.num_node
        base = pair * n
        start = torch.searchsorted(keys, base)
        count = torch.searchsorted(keys, base + n) - start               # known answers of the pair
        size = (n - count).unsqueeze(-1)
        k = (rand * size).long()                                         # index into the ascending list of non-answers
        # number of answers skipped: answers a_0 < a_1 < ... of the pair; a_i - i is non-decreasing
        low = torch.zeros_like(k)
        high = count.unsqueeze(-1).expand_as(k).clone()
        first = start.unsqueeze(-1)
        origin = base.unsqueeze(-1)
        limit = max(len(keys) - 1, 0)
        for _ in range(self.search_steps):
            middle = (low + high) // 2
            probe = keys[(first + middle).clamp(max=limit)] - origin - middle if len(keys) else middle
            right = (probe <= k) & (middle < high)
            low = torch.where(right, middle + 1, low)
            high = torch.where(right, high, middle)
        return k + low

    @torch.no_grad()
    def __call__(self, pos_h_index, pos_t_index, pos_r_index, num_negative, rand=None):
        """(B, num_negative) negatives: tails for the first half of the batch, heads for the second (task.py:105-116).
        `rand`: optional pair of uniform tensors (B // 2, S), (B - B // 2, S) - otherwise drawn with torch.rand in the
        reference's order (tails first)."""
        batch_size = len(pos_h_index)
        half = batch_size // 2
        device = pos_h_index.device
        if rand is None:
            rand = (torch.rand(half, num_negative, device=device), torch.rand(batch_size - half, num_negative, device=device))
        neg_t = self._draw(self.tail_keys, pos_h_index[:half] * self.num_relation + pos_r_index[:half], rand[0])
        neg_h = self._draw(self.head_keys, pos_t_index[half:] * self.num_relation + pos_r_index[half:], rand[1])
        return torch.cat([neg_t, neg_h])


def training_indices(batch, negative):
    """(h, t, r) index matrices (B, 1 + S) of the train branch of `predict` (reference task.py:268-273): column 0 is the
    positive; the first half of the batch corrupts tails, the second half heads."""
    pos_h, pos_t, pos_r = batch.t()
    batch_size, width = len(batch), negative.shape[1] + 1
    h_index = pos_h.unsqueeze(-1).repeat(1, width)
    t_index = pos_t.unsqueeze(-1).repeat(1, width)
    r_index = pos_r.unsqueeze(-1).repeat(1, width)
    t_index[:batch_size // 2, 1:] = negative[:batch_size // 2]
    h_index[batch_size // 2:, 1:] = negative[batch_size // 2:]
    return h_index, t_index, r_index


def training_loss(pred, adversarial_temperature=1.0):
    """BCE over (B, 1 + S) logits, positives in column 0, negatives weighted self-adversarially (reference
    task.py:166-180; `sample_weight: no` in the shipped configs).  Returns the per-row losses (B,)."""
    target = torch.zeros_like(pred)
    target[:, 0] = 1
    loss = F.binary_cross_entropy_with_logits(pred, target, reduction="none")
    neg_weight = torch.ones_like(pred)
    if adversarial_temperature > 0:
        with torch.no_grad():
            neg_weight[:, 1:] = F.softmax(pred[:, 1:] / adversarial_temperature, dim=-1)
    else:
        neg_weight[:, 1:] = 1 / (pred.shape[1] - 1)
    return (loss * neg_weight).sum(dim=-1) / neg_weight.sum(dim=-1)


class FinetuneStep(object):
    """One fine-tuning step with the query batch sharded over the ranks (BASELINE.json configs[2]).

    Every rank receives the same global batch of triples, takes its slab, draws strict negatives, runs the relation
    model and the entity model (train branch, `remove_easy_edges`), back-propagates the mean loss of the GLOBAL batch,
    all-reduces the ~0.78 MB of gradients (one flat NCCL call, `sharding.all_reduce_gradients`) and steps AdamW - what
    DistributedSampler + DDP + `Engine.train` do in the reference (engine.py:48-86).  The tail / head split of the
    negatives follows the position of a triple in the global batch, so the union over ranks is the reference's batch."""

    def __init__(self, model, rel_model, graph, num_negative=128, adversarial_temperature=1.0, lr=5e-4, rank=0, world_size=1):
        self.model, self.rel_model, self.graph = model, rel_model, graph
        self.rel_graph = nbf.construct_relation_graph(graph)
        self.sampler = StrictNegativeSampler(graph)
        self.num_negative, self.temperature = num_negative, adversarial_temperature
        self.rank, self.world_size = rank, world_size
        self.parameters = list(model.parameters()) + list(rel_model.parameters())
        on_cuda = any(p.is_cuda for p in self.parameters)
        self.optimizer = torch.optim.AdamW(self.parameters, lr=lr, capturable=on_cuda)   # capturable: see `capture`
        self._graph = None

    def loss(self, global_batch, rand=None):
        total = len(global_batch)
        start, stop = sharding.query_slab(total, self.rank, self.world_size)
        batch = global_batch[start:stop]
        pos_h, pos_t, pos_r = batch.t()
        device = batch.device
        # negatives: rows before the middle of the GLOBAL batch corrupt tails, the others heads (task.py:105-116)
        cut = min(max(total // 2 - start, 0), stop - start)
        if rand is None:
            rand = (torch.rand(cut, self.num_negative, device=device),
                    torch.rand(stop - start - cut, self.num_negative, device=device))
        sampler = self.sampler
        neg_t = sampler._draw(sampler.tail_keys, pos_h[:cut] * sampler.num_relation + pos_r[:cut], rand[0])
        neg_h = sampler._draw(sampler.head_keys, pos_t[cut:] * sampler.num_relation + pos_r[cut:], rand[1])
        width = self.num_negative + 1
        h_index, t_index, r_index = (x.unsqueeze(-1).repeat(1, width) for x in (pos_h, pos_t, pos_r))
        t_index[:cut, 1:] = neg_t
        h_index[cut:, 1:] = neg_h
        rel_input = self.rel_model(self.rel_graph, pos_r)
        pred = self.model(self.graph, [rel_input], h_index, t_index, r_index, remove_easy_edges=True)
        return training_loss(pred, self.temperature).sum() / total     # this rank's share of the global mean

    def _step(self, global_batch, rand=None):
        loss = self.loss(global_batch, rand)
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        sharding.all_reduce_gradients(self.parameters, average=False)   # sum of shares = gradient of the global mean
        self.optimizer.step()
        return loss.detach()

    def __call__(self, global_batch, rand=None):
        if self._graph is not None and rand is None and global_batch.shape == self._static_batch.shape:
            self._static_batch.copy_(global_batch)
            self._graph.replay()
            return self._static_loss
        return self._step(global_batch, rand)

    def capture(self, global_batch_size, warmup=3):
        """Record the whole step - negatives, both models forward, backward, the NCCL all-reduce, AdamW - as one CUDA graph
        for a fixed global batch size; later calls replay it (the batch is copied into the graph's input).  At 8 triples
        per GPU a step is ~600 launches of ~20 us kernels: eager launch gaps are a sixth of the step.  Nothing in the step
        synchronises with the host (strict negatives, easy-edge masking and the index derive are device-only), which is
        what makes it capturable.  The warm-up steps are real optimisation steps on zero-filled batches' worth of
        triple (0, 0, 0) - call this before training starts or restore the weights afterwards."""
        device = next(p.device for p in self.parameters if p.is_cuda)
        self._static_batch = torch.zeros(global_batch_size, 3, dtype=torch.long, device=device)
        state = [p.detach().clone() for p in self.parameters]
        stream = torch.cuda.Stream(device=device)
        stream.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(stream):
            for _ in range(warmup):                       # creates the optimizer state, the NCCL communicator, the indexes
                self._step(self._static_batch)
        torch.cuda.current_stream(device).wait_stream(stream)
        graph = torch.cuda.CUDAGraph()
        # thread_local: the NCCL watchdog thread polls events while this thread captures the all-reduce
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            self._static_loss = self._step(self._static_batch)
        with torch.no_grad():                             # undo the warm-up / capture-time updates of the weights
            for p, saved in zip(self.parameters, state):
                p.copy_(saved)
            for group_state in self.optimizer.state.values():
                for key, value in group_state.items():
                    if torch.is_tensor(value):
                        value.zero_()
        self._graph = graph
        return self


class ShardedEvaluator(object):
    """Evaluation of one global batch of test triples with the batch sharded over the ranks (BASELINE.json configs[3]):
    `predict` (reference task.py:228-263) on this rank's triples, filtered ranks of the true tail and head on the device
    (task.py:279-315 without the `.cpu()` copies of the (B, 2, N) scores and masks), then one gather of the (B, 2) ranks
    (engine.py:148-150).  The filter (known answers of every (h, r, ?) and (?, r, t)) is a sorted key array per graph."""

    def __init__(self, ranker, filter_graph=None, rank=0, world_size=1):
        self.ranker = ranker
        self.answers = StrictNegativeSampler(filter_graph or ranker.graph)
        self.rank, self.world_size = rank, world_size
        self._captured = {}

    def _known_mask(self, keys, pair, num_node):
        """(B, N) True where the entity is NOT a known answer of the pair (the reference's t_mask / h_mask)."""
        base = pair * num_node
        start = torch.searchsorted(keys, base)
        stop = torch.searchsorted(keys, base + num_node)
        mask = torch.ones(len(pair), num_node, dtype=torch.bool, device=pair.device)
        if len(keys) == 0:
            return mask
        longest = 1 << self.answers.search_steps
        offset = torch.arange(longest, device=pair.device).unsqueeze(0)
        position = (start.unsqueeze(-1) + offset).clamp(max=len(keys) - 1)
        valid = offset < (stop - start).unsqueeze(-1)
        entity = torch.where(valid, keys[position] - base.unsqueeze(-1), torch.full_like(position, num_node))
        padded = torch.cat([mask, torch.ones(len(pair), 1, dtype=torch.bool, device=pair.device)], dim=1)
        padded.scatter_(1, entity, False)
        return padded[:, :num_node]

    def filter_mask(self, batch):
        pos_h, pos_t, pos_r = batch.t()
        a = self.answers
        t_mask = self._known_mask(a.tail_keys, pos_h * a.num_relation + pos_r, a.num_node)
        h_mask = self._known_mask(a.head_keys, pos_t * a.num_relation + pos_r, a.num_node)
        return torch.stack([t_mask, h_mask], dim=1)

    def local_ranks(self, batch, predict=None):
        pred = (predict or self.ranker.predict)(batch)                         # (B, 2, N)
        target = torch.stack([batch[:, 1], batch[:, 0]], dim=1)                # true tail, true head
        return nbf.UltraRanker.rank(pred, target, self.filter_mask(batch))

    @torch.no_grad()
    def __call__(self, global_batch, predict=None):
        """(B_global, 2) filtered ranks on every rank."""
        start, stop = sharding.query_slab(len(global_batch), self.rank, self.world_size)
        ranks = self.local_ranks(global_batch[start:stop], predict)
        return sharding.gather_queries(ranks, len(global_batch))
