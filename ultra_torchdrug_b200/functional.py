"""`generalized_rspmm` - the drop-in operator behind `torchdrug.layers.functional.generalized_rspmm`.

Reference call sites: ultra/layer.py:134-167 (GeneralizedRelationalConvNBF) and :336-369
(GeneralizedRelationalConvNBFMod).  Signature, operand layout, error behaviour and autograd
semantics mirror torchdrug's `layers/functional/spmm.py` (un-vendored):

    generalized_rspmm(sparse, relation, input, sum="add", mul="mul") -> Tensor
        out[i, :] = (sum)_{(i, j, k) in sparse} w_ijk * (relation[k, :] (mul) input[j, :])

Everything below the Python argument checks runs in `libultra_rspmm.so` (hand-written sm_100a kernels
behind the C ABI of `include/ultra_rspmm.h`).  There is no CPU path: CPU tensors raise.

Beyond the drop-in operator, what the layers of `nbf.py` call:

    GraphIndex / graph_index / attach_index      the cached index of one edge set (forward, forward_blocked, forward_pna,
                                                 backward(input_addend=...), derive(new weights))
    rspmm_add_boundary / rspmm_add_one_hot       operator + boundary condition in one differentiable node
    rspmm_pna                                    the four PNA aggregates in one pass (inference)
    layer_norm_relu_residual                     bias + LayerNorm + ReLU + short-cut, fused forward and backward
    combine_linear                               Linear over [input | update] without the cat, tensor cores at fp32 accuracy
    nbf_layer                                    one whole NBFNet layer (sum aggregation) as ONE autograd node (training)
    linear_norm_relu_residual_into / _two        the whole `combine` as one kernel (inference)
    score_head_linear / score_head               the scoring MLP without cat([hidden, query])
"""
import collections
import ctypes
import os

import torch

from . import _lib

__all__ = ["generalized_rspmm", "GraphIndex", "graph_index", "clear_index_cache", "launch_count",
           "layer_norm_relu_residual", "layer_epilogue_supported", "rspmm_add_boundary", "RSPMMAddBoundaryFunction", "rspmm_add_one_hot", "RSPMMAddOneHotFunction", "rspmm_pna", "LayerEpilogueFunction",
           "layer_norm_relu_residual_into", "score_head", "fused_linear_supported",
           "linear_norm_relu_residual_into", "score_head_linear", "attach_index", "combine_linear", "combine_linear_supported", "linear_norm_relu_residual_two", "linear_planes_supported",
           "CombineLinearFunction"]

_SUM_OPS = ("add", "min", "max")
_MUL_OPS = ("mul", "add")
_DTYPE_CODE = {torch.float32: _lib.F32, torch.float64: _lib.F64}

#: LRU of graph indexes found by content fingerprint: at most ULTRA_RSPMM_INDEX_CACHE entries (default 8) and
#: ULTRA_RSPMM_INDEX_CACHE_MB of device memory (default 2048; the newest index is always kept)
INDEX_CACHE_SIZE = int(os.environ.get("ULTRA_RSPMM_INDEX_CACHE", "8"))
INDEX_CACHE_BYTES = int(os.environ.get("ULTRA_RSPMM_INDEX_CACHE_MB", "2048")) << 20
_index_cache = collections.OrderedDict()
#: statistics for tests / benchmarks: how the index of each call was obtained
cache_stats = {"attached": 0, "fingerprint_hit": 0, "built": 0}


def _stream_handle():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(tensor):
    """Raw device address (NULL for None / empty tensors)."""
    if tensor is None or tensor.numel() == 0:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(tensor.data_ptr())


class GraphIndex(object):
    """Device-resident int32 edge orders (CSR / CSC / by-relation) + task lists of one coalesced operand.

    Replaces the per-call `adjacency.transpose(0, 1)` -> `coalesce()` -> `coo2csr3d` of the reference
    (SURVEY.md section 8 row a5): built once per distinct edge set, shared by all layers and by backward.
    """

    def __init__(self, indices, values, shape):
        if indices.dim() != 2 or indices.shape[0] != 3 or indices.dtype != torch.int64:
            raise RuntimeError("Expect `sparse` to have 3 sparse dims with int64 indices, got %s %s"
                               % (tuple(indices.shape), indices.dtype))
        if values.dtype not in _DTYPE_CODE:
            raise RuntimeError("generalized_rspmm supports float32 and float64, got %s" % values.dtype)
        if not indices.is_cuda:
            raise RuntimeError("generalized_rspmm has no CPU implementation: `sparse` must live on a CUDA device")
        lib = _lib.lib()
        self.shape = tuple(int(s) for s in shape)
        self.dtype = values.dtype
        self.device = indices.device
        nnz = indices.shape[1]
        if indices.stride(1) != 1:
            indices = indices.contiguous()
        values = values.contiguous()
        code = _DTYPE_CODE[values.dtype]
        index_bytes, scratch_bytes = ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(lib.ultra_rspmm_index_bytes(nnz, self.shape[0], self.shape[1], self.shape[2], code,
                                               ctypes.byref(index_bytes), ctypes.byref(scratch_bytes)),
                   "ultra_rspmm_index_bytes")
        with torch.cuda.device(self.device):
            self.buffer = torch.empty(max(index_bytes.value, 256), dtype=torch.uint8, device=self.device)
            scratch = torch.empty(max(scratch_bytes.value, 256), dtype=torch.uint8, device=self.device)
            self.c = _lib.Index()
            _lib.check(lib.ultra_rspmm_index_build(
                _ptr(indices), indices.stride(0) if nnz else 0, _ptr(values), nnz, self.shape[0], self.shape[1],
                self.shape[2], code, self.buffer.data_ptr(), self.buffer.numel(), scratch.data_ptr(),
                scratch.numel(), ctypes.byref(self.c), _stream_handle()), "ultra_rspmm_index_build")
            del scratch
            # optional extensions: pair lists (<= 4 relation types, few nodes), destination-block table (huge graphs)
            extend_bytes = ctypes.c_size_t()
            _lib.check(lib.ultra_rspmm_index_extend_bytes(ctypes.byref(self.c), ctypes.byref(extend_bytes)),
                       "ultra_rspmm_index_extend_bytes")
            self.extension = None
            if extend_bytes.value:
                self.extension = torch.empty(extend_bytes.value, dtype=torch.uint8, device=self.device)
                _lib.check(lib.ultra_rspmm_index_extend(ctypes.byref(self.c), self.extension.data_ptr(), self.extension.numel(),
                                                        _stream_handle()), "ultra_rspmm_index_extend")
        self.nnz = int(self.c.nnz)
        self._workspace_bytes = {}

    def derive(self, values):
        """The index of the same edge structure with other edge values (`values`: one per edge of the operand this index
        was built from, in that operand's order).  Shares the structure arrays with `self`, owns new weight arrays and
        task flags; no sort, no host synchronisation (`ultra_rspmm_index_derive`).  Used for the weight-0 form of
        `remove_easy_edges` (reference model.py:57-74) in training steps."""
        if values.shape != (int(self.c.nnz_raw),) or values.dtype != self.dtype or values.device != self.device:
            raise RuntimeError("`values` must hold one %s value per edge of the indexed operand (%d) on %s"
                               % (self.dtype, int(self.c.nnz_raw), self.device))
        derived = object.__new__(GraphIndex)
        derived.shape, derived.dtype, derived.device, derived.nnz = self.shape, self.dtype, self.device, self.nnz
        derived.base = self                       # keeps the structure arrays alive
        derived._workspace_bytes = self._workspace_bytes
        need = ctypes.c_size_t()
        lib = _lib.lib()
        _lib.check(lib.ultra_rspmm_index_derive_bytes(ctypes.byref(self.c), ctypes.byref(need)), "ultra_rspmm_index_derive_bytes")
        with torch.cuda.device(self.device):
            derived.buffer = torch.empty(max(need.value, 256), dtype=torch.uint8, device=self.device)
            derived.c = _lib.Index()
            _lib.check(lib.ultra_rspmm_index_derive(ctypes.byref(self.c), _ptr(values.detach().contiguous()),
                                                    derived.buffer.data_ptr(), derived.buffer.numel(),
                                                    ctypes.byref(derived.c), _stream_handle()), "ultra_rspmm_index_derive")
        return derived

    def workspace_bytes(self, dim):
        if dim not in self._workspace_bytes:
            forward, backward = ctypes.c_size_t(), ctypes.c_size_t()
            _lib.check(_lib.lib().ultra_rspmm_workspace_bytes(ctypes.byref(self.c), dim, _DTYPE_CODE[self.dtype],
                                                              ctypes.byref(forward), ctypes.byref(backward)),
                       "ultra_rspmm_workspace_bytes")
            self._workspace_bytes[dim] = (forward.value, backward.value)
        return self._workspace_bytes[dim]

    # -- raw operator calls (device tensors in, device tensors out) ---------------------------------
    def _check_dense(self, **operands):
        """Shape / dtype / device / layout checks of the dense operands of a raw call: `relation (n_rel, dim)`,
        `input (n_in, dim)`, `output` / `grad_output (n_out, dim)`.  The kernels index these buffers from the graph's
        ids, so a short relation table or a float64 tensor must fail here (the reference raises from TORCH_CHECK),
        never reach the device."""
        rows = {"relation": self.shape[2], "input": self.shape[1], "output": self.shape[0], "grad_output": self.shape[0]}
        dim = None
        for name, tensor in operands.items():
            if tensor is None:
                continue
            if tensor.dim() != 2 or tensor.shape[0] != rows[name]:
                raise RuntimeError("Expect `%s` to be a (%d, dim) matrix, but found %s" % (name, rows[name], tuple(tensor.shape)))
            if dim is None:
                dim = tensor.shape[1]
            if tensor.shape[1] != dim:
                raise RuntimeError("Expect `%s` to have %d features, but found %d" % (name, dim, tensor.shape[1]))
            if tensor.dtype != self.dtype:
                raise RuntimeError("Expect `%s` to be %s like the graph index, but found %s" % (name, self.dtype, tensor.dtype))
            if tensor.device != self.device:
                raise RuntimeError("Expect `%s` on %s like the graph index, but found %s" % (name, self.device, tensor.device))
            if not tensor.is_contiguous():
                raise RuntimeError("Expect `%s` to be contiguous" % name)
        return dim

    def forward(self, relation, input, sum="add", mul="mul", return_argidx=False, addend=None, out=None):
        self._check_dense(relation=relation, input=input)
        if sum not in _SUM_OPS or mul not in _MUL_OPS:
            raise ValueError("No generalized rspmm implementation found for summation `%s` and multiplication `%s`" % (sum, mul))
        dim = input.shape[1]
        if addend is not None and (sum != "add" or addend.shape != (self.shape[0], dim) or addend.dtype != input.dtype
                                   or addend.device != input.device or not addend.is_contiguous()):
            raise RuntimeError("`addend` needs sum='add' and the shape / dtype / device of the output, contiguous")
        if out is not None:
            if out.shape != (self.shape[0], dim) or out.dtype != input.dtype or out.device != input.device or not out.is_contiguous():
                raise RuntimeError("`out` must be a contiguous (%d, %d) tensor of the operands' dtype and device" % (self.shape[0], dim))
            output = out
        else:
            output = torch.empty((self.shape[0], dim), dtype=input.dtype, device=input.device)
        argidx = None
        if return_argidx and sum != "add":
            argidx = torch.empty((self.shape[0], dim), dtype=torch.int32, device=input.device)
        need = self.workspace_bytes(dim)[0]
        with torch.cuda.device(input.device):
            workspace = torch.empty(need, dtype=torch.uint8, device=input.device) if need else None
            _lib.check(_lib.lib().ultra_rspmm_forward(
                ctypes.byref(self.c), _ptr(relation), _ptr(input), _ptr(addend), _ptr(output), _ptr(argidx), dim,
                _DTYPE_CODE[self.dtype], _lib.SUM_CODE[sum], _lib.MUL_CODE[mul], _ptr(workspace), need,
                _stream_handle()), "ultra_rspmm_forward")
        return (output, argidx) if return_argidx else output

    def forward_blocked(self, relation, input_buffer, output_buffer, block, input_offset, output_offset, mul="mul",
                        addend=None):
        """sum-aggregation forward on operands embedded in wider (rows, B, stride) fp32 buffers: the input is
        `input_buffer[..., input_offset:input_offset + block]`, the result (+ addend) is written into
        `output_buffer[..., output_offset:output_offset + block]` (may be the same buffer).  Removes the
        `torch.cat([input, update], -1)` of reference layer.py:387 when the buffer is what the layer's Linear reads."""
        if input_buffer.dim() != 3:
            raise RuntimeError("input buffer must be a (rows, batch, stride) tensor")
        batch = input_buffer.shape[1]
        dim = batch * block
        if mul not in _MUL_OPS:
            raise ValueError("Unknown multiplication `%s`" % mul)
        if block <= 0 or block % 4 or input_offset < 0 or output_offset < 0 or input_offset % 4 or output_offset % 4:
            raise RuntimeError("block and offsets must be non-negative multiples of 4 features")
        if relation.shape != (self.shape[2], dim) or relation.dtype != torch.float32 or relation.device != self.device \
                or not relation.is_contiguous() or self.dtype != torch.float32:
            raise RuntimeError("Expect `relation` to be a contiguous float32 (%d, %d) matrix on %s, but found %s %s on %s"
                               % (self.shape[2], dim, self.device, tuple(relation.shape), relation.dtype, relation.device))
        for name, buffer, offset in (("input", input_buffer, input_offset), ("output", output_buffer, output_offset)):
            if buffer.dim() == 3 and offset + block > buffer.shape[2]:
                raise RuntimeError("%s block [%d, %d) exceeds the buffer's stride %d" % (name, offset, offset + block, buffer.shape[2]))
            if buffer.device != self.device:
                raise RuntimeError("%s buffer must live on %s" % (name, self.device))
        for name, buffer, rows in (("input", input_buffer, self.shape[1]), ("output", output_buffer, self.shape[0])):
            if buffer.dim() != 3 or not buffer.is_contiguous() or buffer.dtype != torch.float32 or buffer.shape[0] != rows \
                    or buffer.shape[1] != batch:
                raise RuntimeError("%s buffer must be a contiguous float32 (%d, %d, stride) tensor" % (name, rows, batch))
        if addend is not None and (addend.shape != (self.shape[0], dim) or addend.dtype != torch.float32
                                   or not addend.is_contiguous() or addend.device != self.device):
            raise RuntimeError("`addend` must be a contiguous float32 (%d, %d) matrix on %s" % (self.shape[0], dim, self.device))
        need = self.workspace_bytes(dim)[0]
        with torch.cuda.device(input_buffer.device):
            workspace = torch.empty(need, dtype=torch.uint8, device=input_buffer.device) if need else None
            _lib.check(_lib.lib().ultra_rspmm_forward_blocked(
                ctypes.byref(self.c), _ptr(relation), ctypes.c_void_p(input_buffer.data_ptr() + 4 * input_offset),
                _ptr(addend), _ptr(output_buffer), dim, _DTYPE_CODE[self.dtype], _lib.MUL_CODE[mul], block,
                input_buffer.shape[2], output_buffer.shape[2], output_offset, _ptr(workspace), need, _stream_handle()),
                "ultra_rspmm_forward_blocked")
        return output_buffer

    def forward_pna(self, relation, input, mul="mul"):
        """(sum, sum of squared operands, max, min) of the messages in one pass - the four operator calls of the
        reference's `pna` aggregation (layer.py:141-144, 343-346) over a single gather per edge.  Forward only."""
        self._check_dense(relation=relation, input=input)
        if mul not in _MUL_OPS:
            raise ValueError("Unknown multiplication `%s`" % mul)
        dim = input.shape[1]
        outputs = [torch.empty((self.shape[0], dim), dtype=input.dtype, device=input.device) for _ in range(4)]
        need = 4 * self.workspace_bytes(dim)[0]
        with torch.cuda.device(input.device):
            workspace = torch.empty(need, dtype=torch.uint8, device=input.device) if need else None
            _lib.check(_lib.lib().ultra_rspmm_forward_pna(
                ctypes.byref(self.c), _ptr(relation), _ptr(input), _ptr(outputs[0]), _ptr(outputs[1]), _ptr(outputs[2]),
                _ptr(outputs[3]), dim, _DTYPE_CODE[self.dtype], _lib.MUL_CODE[mul], _ptr(workspace), need,
                _stream_handle()), "ultra_rspmm_forward_pna")
        return tuple(outputs)

    def backward(self, relation, input, output, grad_output, sum="add", mul="mul", need_relation=True,
                 need_input=True, input_addend=None):
        """(grad_relation, grad_input).  `input_addend` (sum="add" only): an (n_in, dim) tensor added to grad_input in
        the kernel that writes it (`ultra_rspmm_backward_addend`)."""
        self._check_dense(relation=relation, input=input, grad_output=grad_output, output=output if sum != "add" else None)
        if sum not in _SUM_OPS or mul not in _MUL_OPS:
            raise ValueError("No generalized rspmm implementation found for summation `%s` and multiplication `%s`" % (sum, mul))
        if sum != "add" and output is None:
            raise RuntimeError("min / max backward needs the saved `output`")
        if input_addend is not None:
            if sum != "add" or not need_input:
                raise RuntimeError("`input_addend` belongs to grad_input of the sum aggregation")
            if input_addend.shape != input.shape or input_addend.dtype != input.dtype or input_addend.device != input.device \
                    or not input_addend.is_contiguous():
                raise RuntimeError("`input_addend` must be a contiguous tensor with the shape, dtype and device of `input`")
        dim = input.shape[1]
        grad_relation = torch.empty_like(relation) if need_relation else None
        grad_input = torch.empty_like(input) if need_input else None
        need = self.workspace_bytes(dim)[1]
        with torch.cuda.device(input.device):
            workspace = torch.empty(need, dtype=torch.uint8, device=input.device) if need else None
            if input_addend is not None:
                _lib.check(_lib.lib().ultra_rspmm_backward_addend(
                    ctypes.byref(self.c), _ptr(relation), _ptr(input), _ptr(grad_output), _ptr(grad_relation), _ptr(grad_input),
                    _ptr(input_addend), dim, _DTYPE_CODE[self.dtype], _lib.MUL_CODE[mul], _ptr(workspace), need,
                    _stream_handle()), "ultra_rspmm_backward_addend")
            else:
                _lib.check(_lib.lib().ultra_rspmm_backward(
                    ctypes.byref(self.c), _ptr(relation), _ptr(input), _ptr(output), _ptr(grad_output),
                    _ptr(grad_relation), _ptr(grad_input), dim, _DTYPE_CODE[self.dtype], _lib.SUM_CODE[sum],
                    _lib.MUL_CODE[mul], _ptr(workspace), need, _stream_handle()), "ultra_rspmm_backward")
        return grad_relation, grad_input


def clear_index_cache():
    _index_cache.clear()


def launch_count(reset=False):
    """Kernels enqueued by the library so far (`gpu_launches` in bench.py)."""
    lib = _lib.lib()
    count = int(lib.ultra_rspmm_launch_count())
    if reset:
        lib.ultra_rspmm_launch_count_reset()
    return count


def layer_epilogue_supported(x, normalized_dim):
    """The fused epilogue covers fp32 CUDA tensors with 4..128 (power of two) features per row."""
    return (x.is_cuda and x.dtype == torch.float32 and normalized_dim % 4 == 0 and normalized_dim <= 128
            and normalized_dim & (normalized_dim - 1) == 0)


def _epilogue_forward(x, linear_bias, weight, bias, residual, eps, relu):
    dim = x.shape[-1]
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().ultra_layer_norm_relu_residual(
            _ptr(x), _ptr(linear_bias), _ptr(weight), _ptr(bias), _ptr(residual), _ptr(out), x.numel() // max(dim, 1), dim,
            float(eps), int(bool(relu)), _stream_handle()), "ultra_layer_norm_relu_residual")
    return out


class LayerEpilogueFunction(torch.autograd.Function):
    """relu(layer_norm(x + linear_bias) * weight + bias) + residual with a fused, deterministic backward."""

    @staticmethod
    def forward(ctx, x, linear_bias, weight, bias, residual, eps, relu):
        x = x.contiguous()
        tensors = [None if t is None else t.detach().contiguous() for t in (linear_bias, weight, bias, residual)]
        out = _epilogue_forward(x.detach(), tensors[0], tensors[1], tensors[2], tensors[3], eps, relu)
        ctx.save_for_backward(x.detach(), *[t if t is not None else x.new_empty(0) for t in tensors[:3]])
        ctx.present = [t is not None for t in tensors[:3]]
        ctx.has_residual = residual is not None
        ctx.eps, ctx.relu = eps, relu
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, linear_bias, weight, bias = ctx.saved_tensors
        linear_bias, weight, bias = (t if present else None for t, present in zip((linear_bias, weight, bias), ctx.present))
        grad_out = grad_out.contiguous()
        dim = x.shape[-1]
        lib = _lib.lib()
        need = ctypes.c_size_t()
        _lib.check(lib.ultra_layer_norm_relu_residual_backward_bytes(dim, ctypes.byref(need)), "backward_bytes")
        grad_x = torch.empty_like(x)
        sums = torch.empty(3, dim, dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            workspace = torch.empty(need.value, dtype=torch.uint8, device=x.device)
            _lib.check(lib.ultra_layer_norm_relu_residual_backward(
                _ptr(x), _ptr(linear_bias), _ptr(weight), _ptr(bias), _ptr(grad_out), _ptr(grad_x), _ptr(sums[0]),
                _ptr(sums[1]), _ptr(sums[2]), x.numel() // max(dim, 1), dim, float(ctx.eps), int(bool(ctx.relu)),
                _ptr(workspace), need.value, _stream_handle()), "ultra_layer_norm_relu_residual_backward")
        return (grad_x, sums[0] if linear_bias is not None else None, sums[1] if weight is not None else None,
                sums[2] if bias is not None else None, grad_out if ctx.has_residual else None, None, None)


def layer_norm_relu_residual(x, weight=None, bias=None, residual=None, eps=1e-5, relu=True, linear_bias=None):
    """relu(layer_norm(x + linear_bias) * weight + bias) + residual in one pass (reference layer.py:386-392 +
    model.py:126-127), differentiable (fused backward) when any operand requires grad."""
    if not layer_epilogue_supported(x, x.shape[-1]):
        raise RuntimeError("layer_norm_relu_residual needs a float32 CUDA tensor with 4..128 (power of two) features per row")
    if residual is not None and residual.shape != x.shape:
        raise RuntimeError("residual shape %s != input shape %s" % (tuple(residual.shape), tuple(x.shape)))
    operands = [t for t in (x, linear_bias, weight, bias, residual) if t is not None]
    if torch.is_grad_enabled() and any(t.requires_grad for t in operands):
        return LayerEpilogueFunction.apply(x, linear_bias, weight, bias, residual, eps, relu)
    contiguous = [None if t is None else t.contiguous() for t in (linear_bias, weight, bias, residual)]
    return _epilogue_forward(x.contiguous(), contiguous[0], contiguous[1], contiguous[2], contiguous[3], eps, relu)


def _is_row_view(view, shape, device):
    """True for a float32 view of `shape` whose rows (last axis, unit stride) are evenly spaced in memory - what the
    strided kernels address as `base + row * view.stride(-2)`."""
    if view.shape != tuple(shape) or view.dtype != torch.float32 or view.device != device or view.dim() < 2:
        return False
    evenly = all(view.stride(axis) == view.stride(axis + 1) * view.shape[axis + 1] for axis in range(view.dim() - 2))
    return view.stride(-1) == 1 and evenly and view.stride(-2) >= view.shape[-1]


def layer_norm_relu_residual_into(x, out, weight=None, bias=None, residual=None, eps=1e-5, relu=True, linear_bias=None):
    """Strided form of `layer_norm_relu_residual` (inference): `out` and `residual` are (..., dim) views whose rows are
    `stride(-2)` elements apart (e.g. the left halves of (N, B, 2 * dim) layer buffers); x is contiguous."""
    dim = x.shape[-1]
    rows = x.numel() // max(dim, 1)
    if not layer_epilogue_supported(x, dim) or not x.is_contiguous():
        raise RuntimeError("layer_norm_relu_residual_into needs a contiguous float32 CUDA input with 4..128 features per row")
    for name, view in (("out", out), ("residual", residual)):
        if view is not None and not _is_row_view(view, x.shape, x.device):
            raise RuntimeError("`%s` must be a float32 view of shape %s whose rows are evenly spaced" % (name, tuple(x.shape)))
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().ultra_layer_norm_relu_residual_strided(
            _ptr(x), _ptr(linear_bias), _ptr(weight), _ptr(bias),
            ctypes.c_void_p(residual.data_ptr()) if residual is not None else ctypes.c_void_p(0),
            ctypes.c_void_p(out.data_ptr()), rows, dim, residual.stride(-2) if residual is not None else dim,
            out.stride(-2), float(eps), int(bool(relu)), _stream_handle()), "ultra_layer_norm_relu_residual_strided")
    return out


def fused_linear_supported(buffer, out_dim):
    """Whether `linear_norm_relu_residual_into` serves a layer of this width (float32 CUDA, 32 or 64 output features;
    ULTRA_FUSED_LINEAR=0 keeps the cuBLAS Linear + separate epilogue)."""
    return (buffer.is_cuda and buffer.dtype == torch.float32 and out_dim in (32, 64)
            and os.environ.get("ULTRA_FUSED_LINEAR", "1") != "0")


def linear_norm_relu_residual_into(buffer, linear_weight, out, linear_bias=None, weight=None, bias=None, eps=1e-5, relu=True,
                                   shortcut=True):
    """`out = relu(layer_norm(buffer @ linear_weight^T + linear_bias) * weight + bias) + buffer[..., :d]` in one kernel
    (inference): `buffer` is the contiguous (..., 2d) layer buffer [input | update + boundary], `out` a (..., d) view with
    unit feature stride and evenly spaced rows.  3xTF32 split on the tensor cores = fp32 accuracy (csrc/layer_linear.cu)."""
    out_dim = linear_weight.shape[0]
    rows = buffer.numel() // max(buffer.shape[-1], 1)
    if not fused_linear_supported(buffer, out_dim) or not buffer.is_contiguous() or buffer.shape[-1] != 2 * out_dim \
            or linear_weight.shape != (out_dim, 2 * out_dim):
        raise RuntimeError("linear_norm_relu_residual_into needs a contiguous float32 CUDA (..., 2d) buffer, d in {32, 64}")
    if not _is_row_view(out, buffer.shape[:-1] + (out_dim,), buffer.device):
        raise RuntimeError("`out` must be a float32 (..., %d) view whose rows are evenly spaced" % out_dim)
    with torch.cuda.device(buffer.device):
        _lib.check(_lib.lib().ultra_layer_linear_norm_relu_residual(
            _ptr(buffer), buffer.shape[-1], _ptr(linear_weight.contiguous()), _ptr(linear_bias), _ptr(weight), _ptr(bias),
            ctypes.c_void_p(out.data_ptr()), out.stride(-2), rows, out_dim, float(eps), int(bool(relu)),
            int(bool(shortcut)), _stream_handle()), "ultra_layer_linear_norm_relu_residual")
    return out


def combine_linear_supported(input, update, weight):
    """Whether `combine_linear` serves this layer: fp32 CUDA, 64 features per half, a (64, 128) weight
    (ULTRA_FUSED_LINEAR=0 keeps `cat` + the cuBLAS Linear)."""
    return (input.is_cuda and input.dtype == torch.float32 and update.dtype == torch.float32 and input.shape == update.shape
            and input.shape[-1] == 64 and tuple(weight.shape) == (64, 128) and weight.dtype == torch.float32
            and os.environ.get("ULTRA_FUSED_LINEAR", "1") != "0")


class CombineLinearFunction(torch.autograd.Function):
    """`cat([input, update], -1) @ weight^T` of the layer's `combine` (reference layer.py:386-388) without the cat and on the
    tensor cores at fp32 accuracy (3xTF32), forward and backward: `ultra_layer_rows_gemm` (tcgen05 + TMA) for the product
    and for the gradient w.r.t. [input | update], `ultra_layer_rows_gemm_weight` (mma.sync, deterministic fold) for the
    weight gradient.  The bias, LayerNorm, ReLU and the short-cut stay in `layer_norm_relu_residual` (fused forward and
    backward)."""

    @staticmethod
    def forward(ctx, input, update, weight):
        a0 = input.detach().contiguous()
        a1 = update.detach().contiguous()
        w = weight.detach().contiguous()
        rows = a0.numel() // 64
        out = torch.empty(a0.shape[:-1] + (64,), dtype=torch.float32, device=a0.device)
        with torch.cuda.device(a0.device):
            _lib.check(_lib.lib().ultra_layer_rows_gemm(_ptr(a0), 64, _ptr(a1), 64, _ptr(w), _ptr(out), 64, None, 0, None, 0, rows,
                                                        64, 128, _stream_handle()), "ultra_layer_rows_gemm")
        ctx.save_for_backward(a0, a1, w)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        a0, a1, w = ctx.saved_tensors
        dx = grad_output.contiguous()
        rows = dx.numel() // 64
        lib = _lib.lib()
        grad_input = grad_update = grad_weight = None
        with torch.cuda.device(dx.device):
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
                grad_input, grad_update = torch.empty_like(a0), torch.empty_like(a1)
                transposed = w.t().contiguous()                                  # (128, 64): [d input | d update] = dx @ W
                _lib.check(lib.ultra_layer_rows_gemm(_ptr(dx), 64, None, 0, _ptr(transposed), _ptr(grad_input), 64, _ptr(grad_update),
                                                     64, None, 0, rows, 128, 64, _stream_handle()), "ultra_layer_rows_gemm")
            if ctx.needs_input_grad[2]:
                need = ctypes.c_size_t()
                _lib.check(lib.ultra_layer_rows_gemm_weight_bytes(ctypes.byref(need)), "ultra_layer_rows_gemm_weight_bytes")
                workspace = torch.empty(need.value, dtype=torch.uint8, device=dx.device)
                grad_weight = torch.empty_like(w)
                _lib.check(lib.ultra_layer_rows_gemm_weight(_ptr(dx), 64, _ptr(a0), 64, _ptr(a1), 64, rows, _ptr(grad_weight),
                                                            _ptr(workspace), need.value, _stream_handle()),
                           "ultra_layer_rows_gemm_weight")
        return grad_input, grad_update, grad_weight


def combine_linear(input, update, weight):
    """`torch.cat([input, update], dim=-1) @ weight^T` for (..., 64) halves and a (64, 128) weight - the Linear of
    `combine` (reference layer.py:386-388) without its bias - differentiable w.r.t. all three operands."""
    if not combine_linear_supported(input, update, weight):
        raise RuntimeError("combine_linear needs float32 CUDA (..., 64) halves of equal shape and a (64, 128) weight")
    if input.numel() == 0:
        return input.new_zeros(input.shape)
    return CombineLinearFunction.apply(input, update, weight)


def linear_planes_supported(hidden, out_dim):
    """Whether `linear_norm_relu_residual_two` serves this layer: float32 CUDA, 32 or 64 features, the tcgen05 + TMA kernel
    in effect (`ultra_layer_linear_set_kernel(1)` or ULTRA_LAYER_PLANES=0 keep the interleaved (N, B, 2d) buffers)."""
    return (fused_linear_supported(hidden, out_dim) and os.environ.get("ULTRA_LAYER_PLANES", "1") != "0"
            and _lib.lib().ultra_layer_linear_get_kernel() == 2)


def linear_norm_relu_residual_two(input, update, linear_weight, out, linear_bias=None, weight=None, bias=None, eps=1e-5,
                                  relu=True, shortcut=True):
    """`out = relu(layer_norm(cat([input, update], -1) @ linear_weight^T + linear_bias) * weight + bias) + input` with the two
    halves of the Linear's input in two contiguous (..., d) tensors (inference; `ultra_layer_linear_norm_relu_residual_two`):
    no `cat`, no interleaved buffer - the operator reads and writes plain matrices."""
    out_dim = linear_weight.shape[0]
    if not fused_linear_supported(input, out_dim) or input.shape != update.shape or input.shape[-1] != out_dim or \
            not input.is_contiguous() or not update.is_contiguous() or linear_weight.shape != (out_dim, 2 * out_dim) or \
            update.dtype != torch.float32 or update.device != input.device:
        raise RuntimeError("linear_norm_relu_residual_two needs contiguous float32 CUDA (..., d) halves, d in {32, 64}")
    if not _is_row_view(out, input.shape, input.device):
        raise RuntimeError("`out` must be a float32 (..., %d) view whose rows are evenly spaced" % out_dim)
    rows = input.numel() // max(out_dim, 1)
    with torch.cuda.device(input.device):
        _lib.check(_lib.lib().ultra_layer_linear_norm_relu_residual_two(
            _ptr(input), out_dim, _ptr(update), out_dim, _ptr(linear_weight.contiguous()), _ptr(linear_bias), _ptr(weight),
            _ptr(bias), ctypes.c_void_p(out.data_ptr()), out.stride(-2), rows, out_dim, float(eps), int(bool(relu)),
            int(bool(shortcut)), _stream_handle()), "ultra_layer_linear_norm_relu_residual_two")
    return out


def score_head(z, query_bias, weight, bias=None):
    """`relu(z + query_bias[query]) @ weight + bias` for z of shape (N, B, dim) -> (N, B) scores: the second half of the
    2-layer scoring MLP of reference model.py:177-193 in one pass (inference only; see `ultra_score_head`)."""
    num_node, batch, dim = z.shape
    if not layer_epilogue_supported(z, dim) or not z.is_contiguous():
        raise RuntimeError("score_head needs a contiguous float32 CUDA tensor with 4..128 features per row")
    query_bias, weight = query_bias.contiguous(), weight.contiguous().view(-1)
    if query_bias.shape != (batch, dim) or weight.shape != (dim,) or query_bias.dtype != z.dtype or weight.dtype != z.dtype:
        raise RuntimeError("score_head: query_bias must be (%d, %d) and weight (%d,) float32" % (batch, dim, dim))
    score = torch.empty(num_node, batch, dtype=z.dtype, device=z.device)
    with torch.cuda.device(z.device):
        _lib.check(_lib.lib().ultra_score_head(_ptr(z), _ptr(query_bias), _ptr(weight), _ptr(bias), _ptr(score),
                                               num_node * batch, batch, dim, _stream_handle()), "ultra_score_head")
    return score


def score_head_linear(feature, hidden_dim, first_weight, query_bias, out_weight, out_bias=None):
    """Scores (N, B) of every (node, query) row of the (N, B, width) layer buffer `feature`, whose first `hidden_dim`
    columns are the hidden state: `out_weight . relu(first_weight[:, :hidden_dim] hidden + query_bias[query]) + out_bias`
    in one kernel (inference; `ultra_score_head_linear`).  `first_weight` is the MLP's first Linear weight (2d, 2d),
    `query_bias` (B, 2d) its query half applied to the queries plus its bias."""
    num_node, batch, width = feature.shape
    hidden_units = first_weight.shape[0]
    if not fused_linear_supported(feature, hidden_dim) or not feature.is_contiguous() or hidden_units != 2 * hidden_dim \
            or width < hidden_dim or first_weight.stride(-1) != 1 or first_weight.shape[1] < hidden_dim:
        raise RuntimeError("score_head_linear needs a contiguous float32 CUDA buffer and a (2d, >= d) weight, d in {32, 64}")
    query_bias, out_weight = query_bias.contiguous(), out_weight.contiguous().view(-1)
    if query_bias.shape != (batch, hidden_units) or out_weight.shape != (hidden_units,):
        raise RuntimeError("score_head_linear: query_bias must be (%d, %d) and out_weight (%d,)" % (batch, hidden_units, hidden_units))
    score = torch.empty(num_node, batch, dtype=feature.dtype, device=feature.device)
    with torch.cuda.device(feature.device):
        _lib.check(_lib.lib().ultra_score_head_linear(
            _ptr(feature), width, _ptr(first_weight), first_weight.stride(0), _ptr(query_bias), _ptr(out_weight),
            _ptr(out_bias), _ptr(score), num_node * batch, batch, hidden_dim, _stream_handle()), "ultra_score_head_linear")
    return score


def attach_index(sparse, index):
    """Make `graph_index(sparse)` return `index` (e.g. one made by `GraphIndex.derive`) while the operand's indices and
    values are not modified in place."""
    sparse._ultra_rspmm_index = (index, (sparse._indices()._version, sparse._values()._version))
    return sparse


def _fingerprint(indices, values):
    if torch.cuda.is_current_stream_capturing():
        raise RuntimeError("generalized_rspmm: this sparse operand has no graph index yet, and looking one up reads a "
                           "fingerprint back to the host, which cannot be captured into a CUDA graph.  Call the operator "
                           "(or functional.graph_index(sparse)) once on this tensor object before capturing.")
    out = torch.empty(2, dtype=torch.int64, device=indices.device)
    with torch.cuda.device(indices.device):
        _lib.check(_lib.lib().ultra_rspmm_fingerprint(
            _ptr(indices), indices.stride(0) if indices.shape[1] else 0, _ptr(values), indices.shape[1],
            _DTYPE_CODE[values.dtype], out.data_ptr(), _stream_handle()), "ultra_rspmm_fingerprint")
    return tuple(out.tolist())


def graph_index(sparse):
    """The (cached) `GraphIndex` of a sparse COO operand of shape (n_out, n_in, n_rel).

    Lookup order: (1) an index attached to this very tensor object by an earlier call (valid while the
    indices/values version counters are unchanged); (2) a content fingerprint of (indices, values) computed
    on the device, looked up in a small LRU - this is what makes the 6 layers of a forward (each passes a
    fresh `adjacency.transpose(0, 1)`, layer.py:127,328) share one build; (3) build.
    """
    indices, values = sparse._indices(), sparse._values()
    versions = (indices._version, values._version)
    attached = getattr(sparse, "_ultra_rspmm_index", None)
    if attached is not None and attached[1] == versions and attached[0].dtype == values.dtype:
        cache_stats["attached"] += 1
        return attached[0]
    if indices.shape[0] != 3:
        raise RuntimeError("Expect `sparse` to have 3 sparse dims, but found %d" % indices.shape[0])
    if values.dtype not in _DTYPE_CODE:
        raise RuntimeError("generalized_rspmm supports float32 and float64, got %s" % values.dtype)
    if indices.stride(1) != 1:
        indices = indices.contiguous()
    values = values.detach().contiguous()
    key = (_fingerprint(indices, values), tuple(sparse.shape), values.dtype, indices.device, indices.shape[1])
    index = _index_cache.get(key)
    if index is None:
        index = GraphIndex(indices, values, sparse.shape)
        cache_stats["built"] += 1
        _index_cache[key] = index
        while len(_index_cache) > 1 and (len(_index_cache) > max(INDEX_CACHE_SIZE, 1) or
                                         sum(entry.buffer.numel() for entry in _index_cache.values()) > INDEX_CACHE_BYTES):
            _index_cache.popitem(last=False)
    else:
        cache_stats["fingerprint_hit"] += 1
        _index_cache.move_to_end(key)
    try:
        sparse._ultra_rspmm_index = (index, versions)
    except Exception:
        pass
    return index


class RSPMMAddBoundaryFunction(torch.autograd.Function):
    """`generalized_rspmm(sum="add") + boundary` with the addition in the kernel epilogue, differentiable: the gradient
    w.r.t. `boundary` is the upstream gradient itself, the other two are the operator's (mul is a plain attribute)."""

    @staticmethod
    def forward(ctx, sparse, relation, input, boundary, mul):
        index = graph_index(sparse)
        relation, input = relation.contiguous(), input.contiguous()
        ctx.index, ctx.mul = index, mul
        ctx.save_for_backward(relation, input)
        return index.forward(relation, input, "add", mul, addend=boundary.contiguous())

    @staticmethod
    def backward(ctx, output_grad):
        relation, input = ctx.saved_tensors
        relation_grad, input_grad = ctx.index.backward(
            relation, input, None, output_grad.contiguous(), "add", ctx.mul,
            need_relation=ctx.needs_input_grad[1], need_input=ctx.needs_input_grad[2])
        return None, relation_grad, input_grad, output_grad if ctx.needs_input_grad[3] else None, None


class RSPMMAddOneHotFunction(torch.autograd.Function):
    """`generalized_rspmm(sum="add") + boundary` for the one-hot boundary condition of NBFNet (reference model.py:106-109:
    zeros with query[b] at node index[b] of column block b): the boundary is applied as B row updates and its gradient is
    B gathered rows, so neither the (N, D) boundary nor its (N, D) gradient - which autograd would otherwise sum over all
    layers - is ever touched."""

    @staticmethod
    def forward(ctx, sparse, relation, input, node_index, query, mul):
        index = graph_index(sparse)
        relation, input = relation.contiguous(), input.contiguous()
        ctx.index, ctx.mul = index, mul
        ctx.save_for_backward(relation, input, node_index)
        output = index.forward(relation, input, "add", mul)
        batch, width = query.shape
        columns = torch.arange(batch, device=query.device)
        output.view(output.shape[0], batch, width)[node_index, columns] += query
        return output

    @staticmethod
    def backward(ctx, output_grad):
        relation, input, node_index = ctx.saved_tensors
        output_grad = output_grad.contiguous()
        relation_grad, input_grad = ctx.index.backward(
            relation, input, None, output_grad, "add", ctx.mul,
            need_relation=ctx.needs_input_grad[1], need_input=ctx.needs_input_grad[2])
        query_grad = None
        if ctx.needs_input_grad[4]:
            batch = node_index.shape[0]
            columns = torch.arange(batch, device=output_grad.device)
            query_grad = output_grad.view(output_grad.shape[0], batch, -1)[node_index, columns]
        return None, relation_grad, input_grad, None, query_grad, None


def rspmm_add_one_hot(sparse, relation, input, node_index, query, mul="mul"):
    """`generalized_rspmm(sparse, relation, input, sum="add", mul=mul) + boundary` where boundary (N, B * d) is zero except
    `query[b]` (B, d) at row `node_index[b]`, columns [b * d, (b + 1) * d) (reference model.py:106-109 + layer.py:357-358).
    Differentiable w.r.t. relation, input and query."""
    _check_operands(sparse, relation, input)
    if mul not in _MUL_OPS:
        raise ValueError("Unknown multiplication `%s`" % mul)
    if sparse.requires_grad:
        raise RuntimeError("gradient w.r.t. the sparse values is outside the rspmm hot path")
    if query.dim() != 2 or node_index.shape != (query.shape[0],) or query.shape[0] * query.shape[1] != input.shape[1] \
            or query.dtype != input.dtype or sparse.size(0) != sparse.size(1):
        raise RuntimeError("`query` must be (B, d) with B * d == input.size(1), `node_index` (B,), and `sparse` square")
    return RSPMMAddOneHotFunction.apply(sparse, relation, input, node_index, query, mul)


class NBFLayerFunction(torch.autograd.Function):
    """One NBFNet layer with sum aggregation as ONE autograd node (reference layer.py:336-392 + model.py:126-127):

        update = rspmm(adjacency, relation, x) + one_hot_boundary(node_index, query)
        y = relu(layer_norm(cat([x, update]) @ W^T + b) * gamma + beta)  (+ x: the short-cut)

    Forward: the operator, `ultra_layer_rows_gemm` and the fused epilogue, exactly the kernels the three separate nodes
    (`RSPMMAddOneHotFunction`, `CombineLinearFunction`, `LayerEpilogueFunction`) run.  Backward: the layer input x has
    three consumers - the operator, the Linear, the short-cut - and autograd would sum their gradients in two extra passes
    over (N, B, d) tensors per layer (8 % of a C3 fine-tuning step); here the short-cut's gradient enters the `dX` GEMM's
    epilogue and their sum enters the operator's grad_input pass (`ultra_rspmm_backward_addend`)."""

    @staticmethod
    def forward(ctx, sparse, relation, x, node_index, query, weight, linear_bias, gamma, beta, mul, eps, relu, shortcut):
        index = graph_index(sparse)
        x = x.detach().contiguous()
        relation = relation.detach().contiguous()
        num_node, batch, width = x.shape
        update = index.forward(relation, x.view(num_node, batch * width), "add", mul).view(num_node, batch, width)
        columns = torch.arange(batch, device=x.device)
        update[node_index, columns] += query.detach()
        w = weight.detach().contiguous()
        rows = num_node * batch
        lin = torch.empty_like(x)
        small = [None if t is None else t.detach().contiguous() for t in (linear_bias, gamma, beta)]
        lib = _lib.lib()
        if (small[1] is None) == (small[2] is None) and lib.ultra_layer_linear_get_kernel() == 2 and \
                os.environ.get("ULTRA_NBF_FUSED_FORWARD", "1") != "0":
            # Linear + bias + LayerNorm + ReLU + short-cut in one kernel that also returns the Linear's output for the backward
            out = torch.empty_like(x)
            with torch.cuda.device(x.device):
                _lib.check(lib.ultra_layer_linear_norm_relu_residual_two_pre(
                    _ptr(x), 64, _ptr(update), 64, _ptr(w), _ptr(small[0]), _ptr(small[1]), _ptr(small[2]), _ptr(out), 64, _ptr(lin), 64,
                    rows, 64, float(eps), int(bool(relu)), int(bool(shortcut)), _stream_handle()),
                    "ultra_layer_linear_norm_relu_residual_two_pre")
        else:
            with torch.cuda.device(x.device):
                _lib.check(lib.ultra_layer_rows_gemm(_ptr(x), 64, _ptr(update), 64, _ptr(w), _ptr(lin), 64, None, 0, None, 0, rows,
                                                     64, 128, _stream_handle()), "ultra_layer_rows_gemm")
            out = _epilogue_forward(lin, small[0], small[1], small[2], x if shortcut else None, eps, relu)
        ctx.index, ctx.mul, ctx.eps, ctx.relu, ctx.shortcut = index, mul, eps, relu, shortcut
        ctx.present = [t is not None for t in small]
        ctx.save_for_backward(relation, x, update, lin, w, node_index, *[t if t is not None else x.new_empty(0) for t in small])
        return out

    @staticmethod
    def backward(ctx, grad_out):
        relation, x, update, lin, w, node_index, linear_bias, gamma, beta = ctx.saved_tensors
        linear_bias, gamma, beta = (t if present else None for t, present in zip((linear_bias, gamma, beta), ctx.present))
        grad_out = grad_out.contiguous()
        num_node, batch, width = x.shape
        rows = num_node * batch
        lib = _lib.lib()
        need = ctypes.c_size_t()
        grad_lin = torch.empty_like(lin)
        sums = torch.empty(3, width, dtype=x.dtype, device=x.device)
        grad_weight = None
        with torch.cuda.device(x.device):
            _lib.check(lib.ultra_layer_norm_relu_residual_backward_bytes(width, ctypes.byref(need)), "backward_bytes")
            workspace = torch.empty(need.value, dtype=torch.uint8, device=x.device)
            _lib.check(lib.ultra_layer_norm_relu_residual_backward(
                _ptr(lin), _ptr(linear_bias), _ptr(gamma), _ptr(beta), _ptr(grad_out), _ptr(grad_lin), _ptr(sums[0]),
                _ptr(sums[1]), _ptr(sums[2]), rows, width, float(ctx.eps), int(bool(ctx.relu)), _ptr(workspace), need.value,
                _stream_handle()), "ultra_layer_norm_relu_residual_backward")
            # [d x (Linear) + d x (short-cut) | d update] = grad_lin @ W  (+ grad_out on the first half)
            grad_x_partial, grad_update = torch.empty_like(x), torch.empty_like(update)
            transposed = w.t().contiguous()
            _lib.check(lib.ultra_layer_rows_gemm(_ptr(grad_lin), 64, None, 0, _ptr(transposed), _ptr(grad_x_partial), 64,
                                                 _ptr(grad_update), 64, _ptr(grad_out) if ctx.shortcut else None, 64, rows, 128, 64,
                                                 _stream_handle()), "ultra_layer_rows_gemm")
            if ctx.needs_input_grad[5]:
                _lib.check(lib.ultra_layer_rows_gemm_weight_bytes(ctypes.byref(need)), "ultra_layer_rows_gemm_weight_bytes")
                workspace = torch.empty(need.value, dtype=torch.uint8, device=x.device)
                grad_weight = torch.empty_like(w)
                _lib.check(lib.ultra_layer_rows_gemm_weight(_ptr(grad_lin), 64, _ptr(x), 64, _ptr(update), 64, rows, _ptr(grad_weight),
                                                            _ptr(workspace), need.value, _stream_handle()),
                           "ultra_layer_rows_gemm_weight")
        flat = (num_node, batch * width)
        if ctx.needs_input_grad[2]:
            grad_relation, grad_x = ctx.index.backward(relation, x.view(flat), None, grad_update.view(flat), "add", ctx.mul,
                                                       need_relation=ctx.needs_input_grad[1], input_addend=grad_x_partial.view(flat))
            grad_x = grad_x.view(x.shape)
        else:
            grad_relation, _ = ctx.index.backward(relation, x.view(flat), None, grad_update.view(flat), "add", ctx.mul,
                                                  need_relation=ctx.needs_input_grad[1], need_input=False)
            grad_x = None
        grad_query = None
        if ctx.needs_input_grad[4]:
            grad_query = grad_update[node_index, torch.arange(batch, device=x.device)]
        return (None, grad_relation, grad_x, None, grad_query, grad_weight, sums[0] if linear_bias is not None else None,
                sums[1] if gamma is not None else None, sums[2] if beta is not None else None, None, None, None, None)


def nbf_layer_supported(relation, x, query, weight):
    """Whether `nbf_layer` serves this layer: float32 CUDA, (N, B, 64) hidden state, (64, 128) Linear."""
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and x.shape[-1] == 64 and tuple(weight.shape) == (64, 128)
            and weight.dtype == torch.float32 and weight.device == x.device and relation.dtype == torch.float32
            and relation.device == x.device and query.dim() == 2 and query.shape == (x.shape[1], 64) and x.numel() > 0)


def nbf_layer(sparse, relation, x, node_index, query, weight, linear_bias=None, gamma=None, beta=None, mul="mul", eps=1e-5,
              relu=True, shortcut=True):
    """relu(layer_norm(cat([x, rspmm(sparse, relation, x) + boundary]) @ weight^T + linear_bias) * gamma + beta) (+ x), boundary
    = `query[b]` at node `node_index[b]` - one NBFNet layer with sum aggregation (reference layer.py:336-392,
    model.py:106-109, 126-127) as a single differentiable node; x: (N, B, 64), relation: (R, B * 64)."""
    if not nbf_layer_supported(relation, x, query, weight):
        raise RuntimeError("nbf_layer needs float32 CUDA operands: x (N, B, 64), query (B, 64), weight (64, 128)")
    if mul not in _MUL_OPS:
        raise ValueError("Unknown multiplication `%s`" % mul)
    if sparse.requires_grad:
        raise RuntimeError("gradient w.r.t. the sparse values is outside the rspmm hot path")
    _check_operands(sparse, relation, x.view(x.shape[0], -1))
    if node_index.shape != (query.shape[0],) or sparse.size(0) != sparse.size(1):
        raise RuntimeError("`node_index` must be (B,) and `sparse` square")
    return NBFLayerFunction.apply(sparse, relation, x, node_index, query, weight, linear_bias, gamma, beta, mul, eps, relu, shortcut)


def rspmm_add_boundary(sparse, relation, input, boundary, mul="mul"):
    """`generalized_rspmm(sparse, relation, input, sum="add", mul=mul) + boundary` with the addition done in the
    kernel epilogue (reference layer.py:155-156, 357-358): one pass fewer over an (N, D) tensor per layer, under
    autograd as well."""
    _check_operands(sparse, relation, input)
    if mul not in _MUL_OPS:
        raise ValueError("Unknown multiplication `%s`" % mul)
    if sparse.requires_grad:
        raise RuntimeError("gradient w.r.t. the sparse values is outside the rspmm hot path")
    if boundary.shape != (sparse.size(0), input.size(1)) or boundary.dtype != input.dtype or boundary.device != input.device:
        raise RuntimeError("`boundary` must have the shape, dtype and device of the output")
    if torch.is_grad_enabled() and (relation.requires_grad or input.requires_grad or boundary.requires_grad):
        return RSPMMAddBoundaryFunction.apply(sparse, relation, input, boundary, mul)
    index = graph_index(sparse)
    return index.forward(relation.contiguous(), input.contiguous(), "add", mul, addend=boundary.contiguous())


def rspmm_pna(sparse, relation, input, mul="mul"):
    """The four aggregates of the `pna` layer - generalized_rspmm(sum="add"), generalized_rspmm(relation ** 2,
    input ** 2, sum="add"), sum="max", sum="min" (reference layer.py:141-144, 343-346) - in one kernel pass.
    Inference only; under autograd call generalized_rspmm four times."""
    _check_operands(sparse, relation, input)
    if mul not in _MUL_OPS:
        raise ValueError("Unknown multiplication `%s`" % mul)
    if torch.is_grad_enabled() and (relation.requires_grad or input.requires_grad):
        raise RuntimeError("rspmm_pna is an inference-only entry point")
    return graph_index(sparse).forward_pna(relation.contiguous(), input.contiguous(), mul)


def _check_operands(sparse, relation, input):
    """Argument checks of torchdrug's `rspmm_forward_check` (TORCH_CHECK -> RuntimeError)."""
    if not isinstance(sparse, torch.Tensor) or sparse.layout != torch.sparse_coo:
        raise RuntimeError("Expect `sparse` to be a sparse COO tensor")
    if sparse.sparse_dim() != 3 or sparse.dense_dim() != 0:
        raise RuntimeError("Expect `sparse` to be a 3D sparse tensor, but found %dD sparse / %dD dense"
                           % (sparse.sparse_dim(), sparse.dense_dim()))
    if relation.dim() != 2:
        raise RuntimeError("Expect `relation` to be a 2D tensor, but found %dD" % relation.dim())
    if input.dim() != 2:
        raise RuntimeError("Expect `input` to be a 2D tensor, but found %dD" % input.dim())
    if not (sparse.dtype == relation.dtype == input.dtype):
        raise RuntimeError("Expect all tensors to have the same dtype, but found %s, %s and %s"
                           % (sparse.dtype, relation.dtype, input.dtype))
    if not (sparse.device == relation.device == input.device):
        raise RuntimeError("Expect all tensors to be on the same device, but found %s, %s and %s"
                           % (sparse.device, relation.device, input.device))
    if not input.is_cuda:
        raise RuntimeError("generalized_rspmm (B200 build) has no CPU implementation; tensors must be on a CUDA device")
    if sparse.size(1) != input.size(0):
        raise RuntimeError("Expect sparse.size(1) == input.size(0), but found %d and %d" % (sparse.size(1), input.size(0)))
    if sparse.size(2) != relation.size(0):
        raise RuntimeError("Expect sparse.size(2) == relation.size(0), but found %d and %d"
                           % (sparse.size(2), relation.size(0)))
    if relation.size(1) != input.size(1):
        raise RuntimeError("Expect relation.size(1) == input.size(1), but found %d and %d"
                           % (relation.size(1), input.size(1)))


def _forward(ctx, sum, mul, sparse, relation, input):
    _check_operands(sparse, relation, input)
    if sparse.requires_grad:
        raise RuntimeError("gradient w.r.t. the sparse values is outside the rspmm hot path "
                           "(reference layer.py:112,299 use message()+aggregate() when graph.requires_grad)")
    index = graph_index(sparse)
    relation = relation.contiguous()
    input = input.contiguous()
    output = index.forward(relation, input, sum, mul)
    ctx.index = index
    if sum == "add":
        ctx.save_for_backward(relation, input)
    else:
        ctx.save_for_backward(relation, input, output)
    return output


def _backward(ctx, sum, mul, output_grad):
    saved = ctx.saved_tensors
    relation, input = saved[0], saved[1]
    output = saved[2] if len(saved) > 2 else None
    relation_grad, input_grad = ctx.index.backward(
        relation, input, output, output_grad.contiguous(), sum, mul,
        need_relation=ctx.needs_input_grad[1], need_input=ctx.needs_input_grad[2])
    return None, relation_grad, input_grad


def _make_function(sum, mul):
    """RSPMM{Add,Min,Max}{Mul,Add}Function (torchdrug `spmm.py` naming).  forward saves
    (relation, input[, output]); backward returns (None, relation_grad, input_grad) - the gradient w.r.t.
    the sparse values is only produced by the reference when `sparse.requires_grad`, which the hot path
    never requests (layer.py:112,299 route that case to message()+aggregate())."""
    name = "RSPMM%s%sFunction" % (sum.capitalize(), mul.capitalize())

    def forward(ctx, sparse, relation, input):
        return _forward(ctx, sum, mul, sparse, relation, input)

    def backward(ctx, output_grad):
        return _backward(ctx, sum, mul, output_grad)

    function = type(name, (torch.autograd.Function,), {
        "forward": staticmethod(forward), "backward": staticmethod(backward), "sum": sum, "mul": mul,
        "__doc__": _make_function.__doc__})
    return name, function


for _sum in _SUM_OPS:
    for _mul in _MUL_OPS:
        _name, _function = _make_function(_sum, _mul)
        globals()[_name] = _function
        __all__.append(_name)


def generalized_rspmm(sparse, relation, input, sum="add", mul="mul"):
    r"""Generalized relational sparse-dense matrix multiplication (torchdrug semantics).

    .. math:: output_{i,l} = \bigoplus_{(i,j,k) \in sparse} sparse_{i,j,k} \cdot (relation_{k,l} \otimes input_{j,l})

    Parameters:
        sparse (SparseTensor): 3D sparse COO tensor (n_out, n_in, n_rel); need not be coalesced
        relation (Tensor): (n_rel, dim)
        input (Tensor): (n_in, dim)
        sum (str): "add", "min" or "max"
        mul (str): "mul" (DistMult) or "add" (TransE)
    """
    name = "RSPMM%s%sFunction" % (str(sum).capitalize(), str(mul).capitalize())
    if sum not in _SUM_OPS or mul not in _MUL_OPS:
        raise ValueError("No generalized rspmm implementation found for summation `%s` and multiplication `%s`"
                         % (sum, mul))
    return globals()[name].apply(sparse, relation, input)
