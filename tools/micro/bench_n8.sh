# 8-GPU box: bench.py at N=8 (as the driver launches it)
mkdir -p gpurun_out/r2m
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2m/bench_n8_numa.json 2> gpurun_out/r2m/bench_n8_numa.err
echo "bench rc=$?"; grep -o '"e2e": {[^}]*}[^}]*}' gpurun_out/r2m/bench_n8_numa.json | head -3; nvidia-smi topo -m 2>/dev/null | head -14; lscpu | grep -i "numa\|socket\|^CPU(s)" | head -8
