"""`torchdrug.data` stand-in: the `Graph` container fields the hot path reads
(SURVEY.md section 8 row a9): edge_list (E,3)=[node_in,node_out,rel], edge_weight, num_node,
num_relation, adjacency, degree_out, undirected(add_inverse), match, edge_mask, context managers.
"""
import contextlib

import torch


def _memoise_transpose(sparse):
    """Make `sparse.transpose(0, 1)` return the same tensor object every time.

    The reference layers call `graph.adjacency.transpose(0, 1)` once per layer (layer.py:127,328); handing
    back one object lets the operator find the graph index it attached to that object on the first call,
    without fingerprinting the indices again (ultra_torchdrug_b200.functional.graph_index)."""
    plain_transpose = sparse.transpose
    memo = {}

    def transpose(dim0, dim1):
        key = (min(dim0, dim1), max(dim0, dim1))
        if key not in memo:
            memo[key] = plain_transpose(dim0, dim1)
        return memo[key]

    sparse.transpose = transpose
    return sparse


class Graph(object):
    """Relational graph container.  `edge_list[:, 0]` is the source (node_in), `[:, 1]` the
    destination (node_out), `[:, 2]` the relation type, as consumed at reference layer.py:56,82."""

    def __init__(self, edge_list=None, edge_weight=None, num_node=None, num_relation=None,
                 node_feature=None, edge_feature=None, graph_feature=None, meta_dict=None, **kwargs):
        if edge_list is None:
            edge_list = torch.zeros(0, 3 if num_relation else 2, dtype=torch.long)
        edge_list = torch.as_tensor(edge_list, dtype=torch.long)
        if edge_list.dim() != 2 or edge_list.shape[1] not in (2, 3):
            raise ValueError("`edge_list` should be (E, 2) or (E, 3), got %s" % (tuple(edge_list.shape),))
        if edge_weight is None:
            edge_weight = torch.ones(len(edge_list), device=edge_list.device)
        else:
            edge_weight = torch.as_tensor(edge_weight, dtype=torch.float, device=edge_list.device)
        if num_node is None:
            num_node = int(edge_list[:, :2].max()) + 1 if len(edge_list) else 0
        if num_relation is None and edge_list.shape[1] == 3:
            num_relation = int(edge_list[:, 2].max()) + 1 if len(edge_list) else 0
        object.__setattr__(self, "meta_dict", dict(meta_dict or {}))
        object.__setattr__(self, "_scope", None)
        self._edge_list = edge_list
        self._edge_weight = edge_weight
        self.num_node = int(num_node)
        self.num_relation = None if num_relation is None else int(num_relation)
        self._cache = {}
        for name, value, kind in (("node_feature", node_feature, "node"), ("edge_feature", edge_feature, "edge"),
                                  ("graph_feature", graph_feature, "graph")):
            if value is not None:
                self.meta_dict[name] = kind
                object.__setattr__(self, name, value)
        for name, value in kwargs.items():
            object.__setattr__(self, name, value)

    # ------------------------------------------------------------------ attributes
    def __setattr__(self, name, value):
        scope = self.__dict__.get("_scope")
        if scope is not None and not name.startswith("_"):
            self.meta_dict[name] = scope
        object.__setattr__(self, name, value)

    @contextlib.contextmanager
    def _context(self, kind):
        previous = self.__dict__.get("_scope")
        object.__setattr__(self, "_scope", kind)
        try:
            yield
        finally:
            object.__setattr__(self, "_scope", previous)

    def graph(self):
        return self._context("graph")

    def node(self):
        return self._context("node")

    def edge(self):
        return self._context("edge")

    @property
    def data_dict(self):
        return {k: getattr(self, k) for k in self.meta_dict if k in self.__dict__}

    @property
    def edge_list(self):
        return self._edge_list

    @property
    def edge_weight(self):
        return self._edge_weight

    @property
    def num_edge(self):
        return len(self._edge_list)

    @property
    def device(self):
        return self._edge_list.device

    @property
    def requires_grad(self):
        return self.__dict__.get("_force_requires_grad", False) or self._edge_weight.requires_grad

    @requires_grad.setter
    def requires_grad(self, value):
        object.__setattr__(self, "_force_requires_grad", bool(value))

    def requires_grad_(self, mode=True):
        self._edge_weight.requires_grad_(mode)
        return self

    # ------------------------------------------------------------------ derived tensors
    @property
    def adjacency(self):
        """Sparse COO (N, N[, R]) with indices `edge_list.t()` (un-coalesced, no invariant checks)."""
        if "adjacency" not in self._cache or self._edge_weight.requires_grad:
            shape = (self.num_node, self.num_node) + ((self.num_relation,) if self._edge_list.shape[1] == 3 else ())
            adjacency = torch.sparse_coo_tensor(self._edge_list.t(), self._edge_weight, shape,
                                                check_invariants=False)
            if self._edge_weight.requires_grad:
                return adjacency
            _memoise_transpose(adjacency)
            self._cache["adjacency"] = adjacency
        return self._cache["adjacency"]

    @property
    def degree_out(self):
        """Weighted number of edges per destination node (`edge_list[:, 1]`)."""
        if "degree_out" not in self._cache:
            degree = torch.zeros(self.num_node, dtype=self._edge_weight.dtype, device=self.device)
            degree.index_add_(0, self._edge_list[:, 1], self._edge_weight.detach())
            self._cache["degree_out"] = degree
        return self._cache["degree_out"]

    @property
    def degree_in(self):
        if "degree_in" not in self._cache:
            degree = torch.zeros(self.num_node, dtype=self._edge_weight.dtype, device=self.device)
            degree.index_add_(0, self._edge_list[:, 0], self._edge_weight.detach())
            self._cache["degree_in"] = degree
        return self._cache["degree_in"]

    # ------------------------------------------------------------------ transforms
    def _like(self, edge_list, edge_weight, num_relation=None, edge_index=None):
        extra = {}
        for name, kind in self.meta_dict.items():
            if name not in self.__dict__:
                continue
            value = self.__dict__[name]
            if kind == "edge":
                if edge_index is None:
                    continue
                value = value[edge_index]
            extra[name] = value
        meta = {k: v for k, v in self.meta_dict.items() if k in extra}
        return type(self)(edge_list, edge_weight=edge_weight, num_node=self.num_node,
                          num_relation=self.num_relation if num_relation is None else num_relation,
                          meta_dict=meta, **extra)

    def clone(self):
        return self._like(self._edge_list.clone(), self._edge_weight.clone(),
                          edge_index=slice(None))

    def undirected(self, add_inverse=False):
        """Interleave every edge with its flip; with `add_inverse` the flip gets relation r + R.
        The result is memoised per (immutable) graph object: the reference rebuilds it on every forward
        (model.py:166), which would also rebuild the adjacency and make the operator re-identify the edge set."""
        key = ("undirected", bool(add_inverse))
        if key in self._cache and not self._edge_weight.requires_grad and not any(
                kind == "edge" for kind in self.meta_dict.values()):
            return self._cache[key]._fresh_view()
        result = self._undirected(add_inverse)
        if not self._edge_weight.requires_grad:
            self._cache[key] = result
            return result._fresh_view()
        return result

    def _fresh_view(self):
        """A new Graph object sharing this one's tensors and derived-tensor cache (callers attach `query` /
        `boundary` attributes to the graph they get back; those must not leak between calls)."""
        view = type(self)(self._edge_list, edge_weight=self._edge_weight, num_node=self.num_node,
                          num_relation=self.num_relation)
        view._cache = self._cache
        return view

    def _undirected(self, add_inverse=False):
        flipped = self._edge_list[:, [1, 0] + list(range(2, self._edge_list.shape[1]))].clone()
        num_relation = self.num_relation
        if add_inverse:
            if self._edge_list.shape[1] != 3:
                raise ValueError("`add_inverse` needs a relational graph")
            flipped[:, 2] += num_relation
            num_relation = num_relation * 2
        edge_list = torch.stack([self._edge_list, flipped], dim=1).flatten(0, 1)
        edge_weight = torch.stack([self._edge_weight, self._edge_weight], dim=1).flatten()
        return self._like(edge_list, edge_weight, num_relation=num_relation)

    def edge_mask(self, index):
        index = torch.as_tensor(index, device=self.device)
        if index.dtype != torch.bool:
            mask = torch.zeros(self.num_edge, dtype=torch.bool, device=self.device)
            mask[index] = True
            index = mask
        return self._like(self._edge_list[index], self._edge_weight[index], edge_index=index)

    def match(self, pattern):
        """Edges matching each row of `pattern` (-1 = wildcard).  Returns (edge_index, num_match)."""
        pattern = torch.as_tensor(pattern, dtype=torch.long, device=self.device)
        if pattern.dim() == 1:
            pattern = pattern.unsqueeze(0)
        width = self._edge_list.shape[1]
        if pattern.shape[1] != width:
            raise ValueError("pattern width %d != edge_list width %d" % (pattern.shape[1], width))
        num_match = torch.zeros(len(pattern), dtype=torch.long, device=self.device)
        starts = torch.zeros(len(pattern), dtype=torch.long, device=self.device)
        orders = {}
        which = torch.zeros(len(pattern), dtype=torch.long, device=self.device)
        sizes = [self.num_node, self.num_node, max(self.num_relation or 1, 1)][:width]
        wild = pattern < 0
        codes = (wild.long() * (2 ** torch.arange(width, device=self.device))).sum(dim=-1)
        for code in codes.unique().tolist():
            columns = [c for c in range(width) if not (code >> c) & 1]
            rows = (codes == code).nonzero().flatten()
            edge_key = torch.zeros(self.num_edge, dtype=torch.long, device=self.device)
            query_key = torch.zeros(len(rows), dtype=torch.long, device=self.device)
            for c in columns:
                edge_key = edge_key * sizes[c] + self._edge_list[:, c]
                query_key = query_key * sizes[c] + pattern[rows, c]
            edge_key, order = edge_key.sort(stable=True)
            left = torch.searchsorted(edge_key, query_key, right=False)
            right = torch.searchsorted(edge_key, query_key, right=True)
            num_match[rows] = right - left
            starts[rows] = left
            which[rows] = code
            orders[code] = order
        total = int(num_match.sum())
        if total == 0:
            return torch.zeros(0, dtype=torch.long, device=self.device), num_match
        owner = torch.repeat_interleave(torch.arange(len(pattern), device=self.device), num_match)
        offset = torch.arange(total, device=self.device) - (num_match.cumsum(0) - num_match)[owner]
        position = starts[owner] + offset
        edge_index = torch.zeros(total, dtype=torch.long, device=self.device)
        for code, order in orders.items():
            select = which[owner] == code
            edge_index[select] = order[position[select]]
        return edge_index, num_match

    # ------------------------------------------------------------------ device moves
    def to(self, device):
        graph = self._like(self._edge_list.to(device), self._edge_weight.to(device), edge_index=slice(None))
        for name in graph.meta_dict:
            value = graph.__dict__.get(name)
            if isinstance(value, torch.Tensor):
                object.__setattr__(graph, name, value.to(device))
        return graph

    def cuda(self, *args, **kwargs):
        return self.to(torch.device("cuda", *args) if args else "cuda")

    def cpu(self):
        return self.to("cpu")

    @classmethod
    def pack(cls, graphs):
        return PackedGraph(list(graphs))

    def __repr__(self):
        fields = ["num_node=%d" % self.num_node, "num_edge=%d" % self.num_edge]
        if self.num_relation is not None:
            fields.append("num_relation=%d" % self.num_relation)
        return "%s(%s)" % (type(self).__name__, ", ".join(fields))


class PackedGraph(object):
    """Just enough of `Graph.pack` for `rel_graphs[i]` indexing (reference task.py:224,234-239)."""

    def __init__(self, graphs):
        self.graphs = graphs

    def __len__(self):
        return len(self.graphs)

    def __getitem__(self, index):
        return self.graphs[index]

    def to(self, device):
        return PackedGraph([g.to(device) for g in self.graphs])
