import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, lsf_b200
from lsf_b200 import synthetic
canonical, live = synthetic.circle_line_pair_2d(128, shift=(5.0, -3.0), line_shift=-4.0)
opt = lsf_b200.HierarchicalOptimizer2d(tikhonov_term_enabled=True, tikhonov_strength=0.2, gradient_kernel_enabled=True,
                                       kernel=synthetic.sobolev_kernel_1d(), maximum_chunk_size=8, rate=0.2,
                                       maximum_iteration_count=100, maximum_warp_update_threshold=0.0)
for _ in range(3):
    opt.optimize(canonical, live)
torch.cuda.synchronize()
