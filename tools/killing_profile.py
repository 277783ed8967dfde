"""Two iterations of 3D KillingFusion (Killing + level set + 7-tap filter) on cuda:0 -- the command profiled with ncu
(profiles/r2_killing_*.md). Usage: python tools/killing_profile.py [size]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, lsf_b200
from lsf_b200 import synthetic
size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
canonical, live = synthetic.sphere_plane_pair_3d(size, xp=torch, device="cuda")
opt = lsf_b200.SlavchevaOptimizer3d(smoothing_term_method=lsf_b200.SmoothingTermMethod.KILLING, level_set_term_enabled=True,
    max_iterations=2, min_iterations=2, maximum_warp_length_lower_threshold=0.0, sobolev_kernel=synthetic.sobolev_kernel_1d())
out = opt.optimize(live.clone(), canonical)
torch.cuda.synchronize()
print("ok", size, opt.get_iteration_count())
