"""Development: time the host-buffer C-ABI call (bench.py's e2e leg) alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ultra_torchdrug_b200 import _lib, synthetic
edge_list, n, r = synthetic.named_graph(bench.GRAPH)
ms, h2d, d2h = bench.e2e_host_buffers(_lib.lib(), edge_list, n, r, bench.BATCH * bench.HIDDEN, 0, 5, 0)
print("chunk", os.environ.get("ULTRA_RSPMM_CHUNK_COLS", "default"), "e2e %.2f ms  (%.1f GB/s each way)" % (ms, h2d / ms / 1e6))
