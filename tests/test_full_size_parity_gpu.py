"""Element-wise parity with the CPU oracle at the FULL size of every BASELINE.json configuration (`-m gpu`).

For each shape - C2 (FB15k-237), a Zipf-skewed C2, C3 (CoDEx-L, at its per-GPU width 512 and at 4096), C4 (YAGO3-10), a
WN18RR-like low-degree graph, a ragged width (D % 128 != 0), integer-valued operands (exact ties everywhere) and the dense
graph of relations of C2 - all six op combinations run forward + both gradients on the GPU at full width through the
C ABI.  Feature columns are independent, so a column subsample (whole 128-feature slabs spread over the width) is copied
back and compared element by element with the C restatement of the reference's CPU rspmm (oracle/rspmm_cpu_ref.c):

    min / max values and arg-indices : bit-exact
    sums and all gradients           : |error| <= 1e-6 + 1e-5 * sum of |terms|   (oracle accumulated in float64)

Every test also asserts WHICH kernel variant served each pass (`ultra_rspmm_last_pass_info`): L2 eviction hints (KEEP),
grouped task lists, split rows + combine, the gated min/max backward, the rows-in-shared-memory kernel - so every
production template instantiation is checked at the size it is used at (reference call sites: ultra/layer.py:336-369;
the math: ultra/layer.py:52-109).
"""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6

#: name -> (graph, dim, column slabs compared, synthetic options, expectations per pass)
CASES = {
    # C2: in-degree 37 -> plain task list, 7.4 MB slab -> no L2 hints; relation order: 474 long segments -> split
    "c2": dict(graph="fb15k237", dim=4096, slabs=(0, 11, 21, 31),
               expect={"fwd": dict(keep=0, grouped=0, split=False), "gin": dict(keep=0, grouped=0), "grel": dict(keep=0, split=True)}),
    # hub destinations: rows of > 256 edges are split into partial rows + combine
    "c2_zipf": dict(graph="fb15k237", dim=4096, slabs=(0, 17, 31), options=dict(skew=1.0),
                    expect={"fwd": dict(keep=0, split=True), "grel": dict(split=True)}),
    "c2_ragged": dict(graph="fb15k237", dim=1000, slabs=None, expect={"fwd": dict(keep=0, grouped=0)}),
    "c2_ties": dict(graph="fb15k237", dim=1024, slabs=(0, 7), ties=True, expect={}),
    # C3 at the width of one of 8 GPUs (64 / 8 queries): 40 MB slab -> L2 hints; in-degree 14 -> grouped tasks
    "c3": dict(graph="codex_l", dim=512, slabs=None,
               expect={"fwd": dict(keep=1, grouped=1), "gin": dict(keep=1, grouped=1), "grel": dict(keep=1, split=True)}),
    "c3_wide": dict(graph="codex_l", dim=4096, slabs=(3, 30), ops=[("add", "mul"), ("max", "mul")],
                    expect={"fwd": dict(keep=1, grouped=1)}),
    # C4: both gathered slabs of grad_relation together (126 MB) exceed what L2 holds -> destination-blocked pass
    "c4": dict(graph="yago310", dim=4096, slabs=(0, 13, 31), grel_kernel="dst_blocked",
               expect={"fwd": dict(keep=1, grouped=1), "gin": dict(keep=1, grouped=1), "grel": dict(split=True)}),
    # a shape of the configs[4] sweep (E = 4 M, N = E / 32, R' = 474): HBM-resident slabs, 1.3 M (relation, block) runs
    # E = 8.4 M, N = 262,144: the 512-byte slab (134 MB) exceeds L2: HBM-bound generic kernel with hints + blocked grad_relation
    "c5_8m": dict(uniform=(1 << 23, 1 << 18, 474), dim=1024, slabs=(0, 6), grel_kernel="dst_blocked",
                  ops=[("add", "mul"), ("add", "add"), ("min", "mul")], expect={"fwd": dict(keep=1), "grel": dict(split=True)}),
    "c5_4m": dict(uniform=(1 << 22, 1 << 17, 474), dim=1024, slabs=(0, 5), grel_kernel="dst_blocked",
                  ops=[("add", "mul"), ("add", "add"), ("max", "mul")],
                  expect={"fwd": dict(keep=1, grouped=0), "grel": dict(split=True)}),
    # WN18RR-like: in-degree 4.2 -> grouped; 21 MB slab: no hints forward, hints for grad_relation (two gathered operands)
    "wn18rr": dict(graph="wn18rr", dim=4096, slabs=(0, 16, 31),
                   expect={"fwd": dict(keep=0, grouped=1), "gin": dict(keep=0, grouped=1), "grel": dict(keep=1, split=True)}),
}


def _columns(dim, slabs):
    if slabs is None:
        return np.arange(dim)
    return np.concatenate([np.arange(128 * s, min(dim, 128 * (s + 1))) for s in slabs])


class _Case(object):
    """Operands on the GPU at full width + the oracle's CSR of the same coalesced graph."""

    def __init__(self, name, device):
        from oracle import cpu_ref
        from ultra_torchdrug_b200 import functional as F, synthetic
        spec = CASES[name]
        if "uniform" in spec:
            e_raw, n, r = spec["uniform"]
            generator = torch.Generator().manual_seed(1024)
            indices = torch.stack([torch.randint(n, (e_raw,), generator=generator), torch.randint(n, (e_raw,), generator=generator),
                                   torch.randint(r, (e_raw,), generator=generator)])
        else:
            edge_list, n, r = synthetic.named_graph(spec["graph"], **spec.get("options", {}))
            indices = edge_list[:, [1, 0, 2]].t().contiguous()
        self.spec, self.n, self.r, self.dim = spec, n, r, spec["dim"]
        values = torch.ones(indices.shape[1])
        self.index = F.GraphIndex(indices.to(device), values.to(device), (n, n, r))
        self.csr = cpu_ref.CsrOperand(indices.numpy(), values.numpy(), (n, n, r))
        generator = torch.Generator(device=device).manual_seed(1024)
        if spec.get("ties"):
            make = lambda rows: torch.randint(-2, 3, (rows, self.dim), device=device, generator=generator).float()
        else:
            make = lambda rows: torch.randn(rows, self.dim, device=device, generator=generator)
        self.relation, self.input = make(r), make(n)
        self.grad = torch.randn(n, self.dim, device=device, generator=generator)
        self.columns = _columns(self.dim, spec["slabs"])
        self._pick = torch.from_numpy(self.columns).to(device)

    def host(self, tensor):
        return tensor.index_select(1, self._pick).cpu().numpy()


_cases = {}


def _case(name, device):
    if name not in _cases:
        _cases.clear()                      # one full-size case resident at a time
        torch.cuda.empty_cache()
        _cases[name] = _Case(name, device)
    return _cases[name]


def _assert_close(actual, expected64, scale64, what):
    err = np.abs(actual.astype(np.float64) - expected64)
    bound = ATOL + RTOL * scale64
    if not (err <= bound).all():
        worst = np.unravel_index(np.argmax(err - bound), err.shape)
        raise AssertionError("%s: error %g > bound %g at %s" % (what, err[worst], bound[worst], worst))


def _check_info(info, want, what):
    for key, value in want.items():
        if key == "split":
            assert (info["n_split"] > 0) == value, "%s: n_split = %d, expected split=%s" % (what, info["n_split"], value)
        else:
            assert info[key] == value, "%s: %s = %s, expected %s (%s)" % (what, key, info[key], value, info)


def _params():
    out = []
    for name, spec in CASES.items():
        for sum, mul in spec.get("ops", util.OPS):
            out.append(pytest.param(name, sum, mul, id="%s-%s-%s" % (name, sum, mul)))
    return out


@pytest.mark.parametrize("name,sum,mul", _params())
def test_full_size_matches_oracle(cuda, name, sum, mul):
    from oracle import cpu_ref
    from ultra_torchdrug_b200 import _lib
    case = _case(name, cuda)
    index, csr, expect = case.index, case.csr, case.spec["expect"]
    out, arg = index.forward(case.relation, case.input, sum, mul, return_argidx=True)
    forward_info = _lib.pass_info(_lib.PASS_FORWARD)
    g_rel, g_in = index.backward(case.relation, case.input, out, case.grad, sum, mul)
    torch.cuda.synchronize()
    gin_info, grel_info = _lib.pass_info(_lib.PASS_GRAD_INPUT), _lib.pass_info(_lib.PASS_GRAD_RELATION)

    # ---- which kernels ran -----------------------------------------------------------------------------------------
    forward_kernel = case.spec.get("fwd_kernel", "seg_reduce") if sum == "add" else "seg_reduce"
    assert forward_info["kernel_name"] == forward_kernel and forward_info["vec"] == 4 and forward_info["packed"] == 1, forward_info
    if forward_kernel == "subwarp_rows":
        assert forward_info["n_slab"] == case.dim // 64                     # 2 tasks per warp, 64-feature slabs
    want = dict(expect.get("fwd", {}))
    if sum != "add":
        want.pop("grouped", None)           # tasks that carry an arg-index are never grouped
        assert forward_info["grouped"] == 0
    _check_info(forward_info, want, "forward")
    backward_kernel = "seg_reduce" if sum == "add" else "seg_gated"
    assert gin_info["kernel_name"] == (forward_kernel if sum == "add" else backward_kernel), gin_info
    # the destination-blocked pass serves DistMult here; TransE (grad_output is the only gathered operand) keeps the generic
    # kernel until that one slab is far beyond L2 (n_out * 512 B > 150 MB: none of these shapes)
    # min / max: the gated destination-blocked pass wherever the block table exists and the two slabs exceed L2
    gated = "dst_blocked_gated" if "grel_kernel" in case.spec else "seg_gated"
    want_grel = gated if sum != "add" else (case.spec.get("grel_kernel", "seg_reduce") if mul == "mul" else "seg_reduce")
    assert grel_info["kernel_name"] == want_grel, grel_info
    if sum == "add":
        _check_info(gin_info, expect.get("gin", {}), "grad_input")
        want = dict(expect.get("grel", {}))
        if mul == "add":
            want.pop("keep", None)          # TransE: grad_relation gathers grad_output only (half the slab bytes)
            if "grel_kernel" in case.spec:
                want["keep"] = 1            # ... through the generic kernel with L2 hints on these large shapes
        _check_info(grel_info, want, "grad_relation")
    else:
        _check_info(grel_info, {k: v for k, v in expect.get("grel", {}).items() if k == "split"}, "grad_relation")

    # ---- element-wise comparison on the column subsample -----------------------------------------------------------
    relation, input, grad = case.host(case.relation), case.host(case.input), case.host(case.grad)
    got = case.host(out)
    if sum == "add":
        _assert_close(got, cpu_ref.forward_f64(csr, relation, input, mul),
                      cpu_ref.forward_f64(csr, relation, input, mul, absolute=True), "forward %s/%s" % (sum, mul))
        assert arg is None
    else:
        want_out, want_arg = cpu_ref.forward_arg(csr, relation, input, sum, mul)
        assert np.array_equal(got, want_out), "forward %s/%s values are not bit-exact" % (sum, mul)
        assert np.array_equal(case.host(arg).astype(np.int64), want_arg), "arg-index differs from the oracle's"
    want_rel, want_in = cpu_ref.backward_f64(csr, relation, input, got, grad, sum, mul)
    scale_rel, scale_in = cpu_ref.backward_f64(csr, relation, input, got, grad, sum, mul, absolute=True)
    _assert_close(case.host(g_rel), want_rel, scale_rel, "grad_relation %s/%s" % (sum, mul))
    _assert_close(case.host(g_in), want_in, scale_in, "grad_input %s/%s" % (sum, mul))


@pytest.mark.parametrize("name", ["c2", "c4"])
@pytest.mark.parametrize("mul", ["mul", "add"])
def test_full_size_pna_matches_oracle(cuda, name, mul):
    """The fused four-aggregate pass (reference layer.py:343-346) against the oracle - not against the library's own
    separate calls: sum and sum of squared operands vs the float64 evaluation, max / min bit-exact."""
    from oracle import cpu_ref
    from ultra_torchdrug_b200 import _lib
    case = _case(name, cuda)
    total, squares, maximum, minimum = case.index.forward_pna(case.relation, case.input, mul)
    info = _lib.pass_info(_lib.PASS_FORWARD)
    assert info["kernel_name"] == "seg_pna" and info["keep"] == case.spec["expect"]["fwd"]["keep"]
    relation, input = case.host(case.relation), case.host(case.input)
    _assert_close(case.host(total), cpu_ref.forward_f64(case.csr, relation, input, mul),
                  cpu_ref.forward_f64(case.csr, relation, input, mul, absolute=True), "pna sum")
    squared = (relation * relation, input * input)          # formed in float32 first, as `relation ** 2` is in the reference
    _assert_close(case.host(squares), cpu_ref.forward_f64(case.csr, squared[0], squared[1], mul),
                  cpu_ref.forward_f64(case.csr, squared[0], squared[1], mul, absolute=True), "pna sum of squares")
    assert np.array_equal(case.host(maximum), cpu_ref.forward(case.csr, relation, input, "max", mul))
    assert np.array_equal(case.host(minimum), cpu_ref.forward(case.csr, relation, input, "min", mul))


# ---- the graph of relations (rows-in-shared-memory kernel) ------------------------------------------------------------
def _relation_graph_operand(device, name="fb15k237"):
    """The dense graph of relations `construct_relation_graph` (reference rel_model.py:91-147) yields on a uniform
    synthetic graph: every (relation, relation, kind) triple, 4 * R'^2 edges."""
    from ultra_torchdrug_b200 import synthetic
    num_relation = 2 * synthetic.SHAPES[name][1]
    grid = torch.cartesian_prod(torch.arange(num_relation), torch.arange(num_relation), torch.arange(4))
    return grid.t().contiguous(), num_relation


@pytest.mark.parametrize("mul", ["mul", "add"])
@pytest.mark.parametrize("dim", [4096, 1000])
def test_relation_graph_full_size_matches_oracle(cuda, mul, dim):
    """C2's graph of relations (474 nodes, 898,704 edges, 4 edge types) at the inference width.  By default the forward and
    the grad_input pass are served by the pair kernel (4 edges per (node, node) pair); without pair lists by the
    rows-in-shared-memory kernel; without either by the generic kernel.  All three agree with the oracle element-wise."""
    from oracle import cpu_ref
    from ultra_torchdrug_b200 import functional as F, _lib
    lib = _lib.lib()
    indices, n = _relation_graph_operand(cuda)
    values = torch.ones(indices.shape[1])
    csr = cpu_ref.CsrOperand(indices.numpy(), values.numpy(), (n, n, 4))
    generator = torch.Generator(device=cuda).manual_seed(7)
    relation = torch.randn(4, dim, device=cuda, generator=generator)
    input = torch.randn(n, dim, device=cuda, generator=generator)
    grad = torch.randn(n, dim, device=cuda, generator=generator)
    columns = _columns(dim, (0, 15, 31) if dim == 4096 else None)
    pick = torch.from_numpy(columns).to(cuda)
    host = lambda t: t.index_select(1, pick).cpu().numpy()
    relation_h, input_h, grad_h = host(relation), host(input), host(grad)
    want = cpu_ref.forward_f64(csr, relation_h, input_h, mul)
    scale = cpu_ref.forward_f64(csr, relation_h, input_h, mul, absolute=True)
    want_rel, want_in = cpu_ref.backward_f64(csr, relation_h, input_h, None, grad_h, "add", mul)
    scale_rel, scale_in = cpu_ref.backward_f64(csr, relation_h, input_h, None, grad_h, "add", mul, absolute=True)

    def run(expected_kernel):
        index = F.GraphIndex(indices.to(cuda), values.to(cuda), (n, n, 4))
        out = index.forward(relation, input, "add", mul)
        assert _lib.pass_info(_lib.PASS_FORWARD)["kernel_name"] == expected_kernel, _lib.pass_info(_lib.PASS_FORWARD)
        g_rel, g_in = index.backward(relation, input, out, grad, "add", mul)
        torch.cuda.synchronize()
        assert _lib.pass_info(_lib.PASS_GRAD_INPUT)["kernel_name"] == expected_kernel
        _assert_close(host(out), want, scale, "forward (%s)" % expected_kernel)
        _assert_close(host(g_rel), want_rel, scale_rel, "grad_relation")
        _assert_close(host(g_in), want_in, scale_in, "grad_input (%s)" % expected_kernel)
        assert torch.equal(out, index.forward(relation, input, "add", mul)), "two runs differ"

    run("pairs_in_smem")
    try:
        _lib.check(lib.ultra_rspmm_set_extensions(0, 1), "ultra_rspmm_set_extensions")
        run("rows_in_smem")
        _lib.check(lib.ultra_rspmm_set_staged(0), "ultra_rspmm_set_staged")
        run("seg_reduce")
    finally:
        _lib.check(lib.ultra_rspmm_set_extensions(1, 1), "ultra_rspmm_set_extensions")
        _lib.check(lib.ultra_rspmm_set_staged(1), "ultra_rspmm_set_staged")
