#!/usr/bin/env python
"""One iNeRF refinement step (bench.inerf_step_ms) and one coarse matching call, for ncu captures of the backward-pass and
matcher kernels (gemm3_tc_kernel, composite_backward_kernel, encode_backward_kernel, row_stats_kernel ...)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
print(json.dumps(bench.inerf_step_ms(dev, os.environ.get("NMR_PRECISION", "bf16"), steps=1)))
