set -x
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_scale_c4m4_n$N.json 2> gpurun_out/r2_scale_c4m4_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c8m16 --steps 20 --warmup 5 > gpurun_out/r2_scale_c8m16_n$N.json 2> gpurun_out/r2_scale_c8m16_n$N.err
tail -n 3 gpurun_out/r2_scale_c4m4_n$N.err gpurun_out/r2_scale_c8m16_n$N.err | cut -c1-300
tail -n 1 gpurun_out/r2_scale_c4m4_n$N.json gpurun_out/r2_scale_c8m16_n$N.json | cut -c1-700
