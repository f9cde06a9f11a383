# bench.py at N GPUs of one box, the way the driver launches it (run under `gpurun --gpus N`): bash tools/scale_run.sh N
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_r2_n$N.json 2> gpurun_out/scale_r2_n$N.err
echo "rc=$?"; head -c 300 gpurun_out/scale_r2_n$N.json; tail -3 gpurun_out/scale_r2_n$N.err
