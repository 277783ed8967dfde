import sys, os
sys.path.insert(0, "/root/repo")
import torch, lsf_b200
from lsf_b200 import synthetic
size = int(sys.argv[1])
canonical, live = synthetic.sphere_plane_pair_3d(size, xp=torch, device="cuda")
opt = lsf_b200.SlavchevaOptimizer3d(smoothing_term_method=lsf_b200.SmoothingTermMethod.KILLING, level_set_term_enabled=True,
    max_iterations=2, min_iterations=2, maximum_warp_length_lower_threshold=0.0, sobolev_kernel=synthetic.sobolev_kernel_1d())
out = opt.optimize(live.clone(), canonical)
torch.cuda.synchronize()
print("ok", size, opt.get_iteration_count())
