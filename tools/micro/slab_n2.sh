# 2-GPU check of the slab decomposition: peer-memory exchange against the NCCL exchange, bit-check against the whole volume
mkdir -p gpurun_out/r2j
N=${1:-2}
run() { timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/slab_multigpu_check.py "$@" 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$\|NCCL version" | tail -4; }
run --size 256 --check --iterations 20 --exchange peer | tee -a gpurun_out/r2j/slab_n$N.log
run --size 512 --check --iterations 20 --repeat 3 --exchange peer | tee -a gpurun_out/r2j/slab_n$N.log
run --size 512 --iterations 20 --repeat 3 --exchange dist | tee -a gpurun_out/r2j/slab_n$N.log
