# weak-scaling sweep of bench.py plus the full config 5 on all GPUs (run under `gpurun --gpus 8`)
#   SCALE_NS="1 2 4 8" (default) picks the GPU counts, SCALE_C5=0 skips the config-5 run
for N in ${SCALE_NS:-1 2 4 8}; do
  if [ $N = 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --gather --no-cpu > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; fi
  tail -c 300 gpurun_out/scale_n$N.json | head -c 300; echo
done
if [ "${SCALE_C5:-1}" = 1 ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/config5_report.py --slabs 256 --steps 2 > gpurun_out/config5_n8.json 2> gpurun_out/config5_n8.err
cat gpurun_out/config5_n8.json | cut -c1-700
fi
