import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lsf_b200, bench
from lsf_b200 import synthetic
torch.cuda.set_device(0)
c, l = synthetic.sphere_plane_pair_3d(256)
opt = lsf_b200.HierarchicalOptimizer3d(**bench.optimizer_kwargs())
def t(f, n=4):
    f(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): r=f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
print("pinned alloc 201MB: %.2f ms" % t(lambda: torch.empty((256,256,256,3), dtype=torch.float32, pin_memory=True)))
print("np.empty+touch: %.2f ms" % t(lambda: np.empty((256,256,256,3), np.float32).fill(0)))
cd, ld = torch.from_numpy(c).cuda(), torch.from_numpy(l).cuda()
print("device optimize: %.2f ms" % t(lambda: opt.optimize(cd, ld)))
print("pageable e2e (pinned results): %.2f ms" % t(lambda: opt.optimize(c, l)))
os.environ["LSF_PINNED_RESULTS"]="0"
print("pageable e2e (np.empty results): %.2f ms" % t(lambda: opt.optimize(c, l)))
out = np.empty((256,256,256,3), np.float32)
print("pageable e2e (out= reused pageable): %.2f ms" % t(lambda: opt.optimize(c, l, out=out)))
pc = torch.from_numpy(c).pin_memory().numpy(); pl = torch.from_numpy(l).pin_memory().numpy(); po = torch.empty((256,256,256,3), dtype=torch.float32, pin_memory=True).numpy()
print("pinned e2e: %.2f ms" % t(lambda: opt.optimize(pc, pl, out=po)))
print("pinned in, pageable out=: %.2f ms" % t(lambda: opt.optimize(pc, pl, out=out)))
print("pageable in, pinned out=: %.2f ms" % t(lambda: opt.optimize(c, l, out=po)))
