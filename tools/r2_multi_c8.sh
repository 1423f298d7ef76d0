set -x
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c8m16 --steps 20 --warmup 5 > gpurun_out/r2_scale_c8m16_n$N.json 2> gpurun_out/r2_scale_c8m16_n$N.err
tail -n 3 gpurun_out/r2_scale_c8m16_n$N.err | cut -c1-300
tail -n 1 gpurun_out/r2_scale_c8m16_n$N.json | cut -c1-900
