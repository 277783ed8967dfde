# 8-GPU run of the slab decomposition: bit-check at 512^3, 1024^3 with the peer-memory exchange and with the NCCL exchange
mkdir -p gpurun_out/r2j
N=${1:-8}
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/slab_multigpu_check.py "$@" 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$\|NCCL version" | tail -4; }
run --size 512 --check --iterations 20 --exchange peer | tee -a gpurun_out/r2j/slab_n$N.log
run --size 1024 --iterations 20 --repeat 3 --exchange peer | tee -a gpurun_out/r2j/slab_n$N.log
run --size 1024 --iterations 20 --repeat 2 --exchange dist | tee -a gpurun_out/r2j/slab_n$N.log
