"""Kernel-time breakdown of one ULTRA fine-tuning step (development tool)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultra_torchdrug_b200 import nbf, synthetic
from ultra_torchdrug_b200.compat.torchdrug import data
name = sys.argv[1] if len(sys.argv) > 1 else "codex_l"
batch_size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
device = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = False
num_node, num_relation, num_triple = synthetic.SHAPES[name]
triples = synthetic.triples(num_node, num_relation, num_triple)
graph = data.Graph(triples, num_node=num_node, num_relation=num_relation).to(device)
model, rel_model = nbf.ultra_models(num_relation)
model, rel_model = model.to(device).train(), rel_model.to(device).train()
rel_graph = nbf.construct_relation_graph(graph)
params = list(model.parameters()) + list(rel_model.parameters())
opt = torch.optim.AdamW(params, lr=5e-4)
def step():
    batch = triples[torch.randint(num_triple, (batch_size,))].to(device)
    neg = torch.randint(num_node, (batch_size, 128), device=device)
    h, t, r = batch.t()
    hi, ti, ri = (x.unsqueeze(-1).repeat(1, 129) for x in (h, t, r))
    ti[:batch_size // 2, 1:] = neg[:batch_size // 2]
    hi[batch_size // 2:, 1:] = neg[batch_size // 2:]
    rel_input = rel_model(rel_graph, r)
    pred = model(graph, [rel_input], hi, ti, ri, remove_easy_edges=True)
    target = torch.zeros_like(pred); target[:, 0] = 1
    loss = torch.nn.functional.binary_cross_entropy_with_logits(pred, target)
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
by_op = os.environ.get("ULTRA_PROFILE_OPS") == "1"      # aten-level view (which autograd ops the elementwise kernels are)
activities = [ProfilerActivity.CUDA, ProfilerActivity.CPU] if by_op else [ProfilerActivity.CUDA]
with profile(activities=activities, record_shapes=by_op) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=by_op).table(sort_by="cuda_time_total", row_limit=40 if by_op else 22,
                                                         max_name_column_width=60 if by_op else 80))
