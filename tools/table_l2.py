"""Isolates the relation table's share of the forward kernel's L2 traffic (BASELINE.json north_star: "L2 hit rate for
the relation table").  ncu counts sectors per kernel, not per array, so the table is isolated by difference: the same
forward runs twice under ncu, (1) on the graph as it is and (2) on the same (destination, source) structure with every
edge's relation set to 0 - the kernel then keeps the one relation row in registers and issues (almost) no table loads.

    ncu --metrics lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,\\
l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,\\
dram__bytes_read.sum,gpu__time_duration.sum -k regex:seg_reduce -c 2 --csv --log-file table_l2.csv \\
        python tools/table_l2.py --graph fb15k237
    python tools/table_l2.py --read table_l2.csv          # prints the table's L1 / L2 hit rates

table L2 sectors = sectors(1) - sectors(2), table L2 hits = hits(1) - hits(2).
"""
import argparse
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def read(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    header = rows[0]
    name, value, ident = header.index("Metric Name"), header.index("Metric Value"), header.index("ID")
    launches = {}
    for row in rows[1:]:
        launches.setdefault(row[ident], {})[row[name]] = float(row[value].replace(",", ""))
    first, second = (launches[k] for k in sorted(launches, key=int)[:2])
    l2 = "lts__t_sectors_srcunit_tex_op_read.sum"
    l2_hit = "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum"
    l1 = "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"
    l1_hit = "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum"
    table_l1 = first[l1] - second[l1]
    table_l1_hit = first[l1_hit] - second[l1_hit]
    table_l2 = first[l2] - second[l2]
    table_l2_hit = first[l2_hit] - second[l2_hit]
    print("forward with the relation table : L1 sectors %.4g (hit %.1f%%), L2 sectors %.4g (hit %.1f%%)"
          % (first[l1], 100 * first[l1_hit] / first[l1], first[l2], 100 * first[l2_hit] / first[l2]))
    print("forward, one relation row only  : L1 sectors %.4g (hit %.1f%%), L2 sectors %.4g (hit %.1f%%)"
          % (second[l1], 100 * second[l1_hit] / second[l1], second[l2], 100 * second[l2_hit] / second[l2]))
    print("relation table (difference)     : L1 sectors %.4g, hit rate %.1f%%;  L2 sectors %.4g (%.2f GB), hit rate %.1f%%"
          % (table_l1, 100 * table_l1_hit / max(table_l1, 1), table_l2, table_l2 * 32 / 1e9, 100 * table_l2_hit / max(table_l2, 1)))


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--graph", default="fb15k237")
    parser.add_argument("--batch", type=int, default=64)
    parser.add_argument("--read", default=None)
    args = parser.parse_args()
    if args.read:
        return read(args.read)
    import torch
    from ultra_torchdrug_b200 import functional as F, synthetic
    device = torch.device("cuda", 0)
    edge_list, n, r = synthetic.named_graph(args.graph)
    d = args.batch * 64
    generator = torch.Generator(device=device).manual_seed(1)
    relation = torch.randn(r, d, device=device, generator=generator)
    input = torch.randn(n, d, device=device, generator=generator)
    indices = edge_list[:, [1, 0, 2]].t().contiguous().to(device)
    for one_row in (False, True):
        if one_row:
            indices = indices.clone()
            indices[2] = 0
        index = F.GraphIndex(indices, torch.ones(indices.shape[1], device=device), (n, n, r))
        out = index.forward(relation, input)
        torch.cuda.synchronize()
    print("done", float(out[0, 0]))


if __name__ == "__main__":
    main()
