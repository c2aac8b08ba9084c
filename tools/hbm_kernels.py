#!/usr/bin/env python
"""The HBM-bound kernels of the path (Adam, context aggregation, adaptive max-pool + flatten, loss) alone, at sizes
larger than L2 -- the command `ncu --metrics dram__...` wraps for profiles/ncu_hbm_kernels_rN.txt; prints the same
CUDA-event numbers bench.py reports under `hbm_kernels`."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200")]
import torch
import bench
print(json.dumps(bench.hbm_bound_kernels(torch.device("cuda:0"), 20, 15, 21, 3195956), indent=1))
