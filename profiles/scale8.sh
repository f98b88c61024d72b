mkdir -p gpurun_out/r2
nvidia-smi -L | head -8 > gpurun_out/r2/scale8_gpus.txt
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2/scale8_n1.json 2> gpurun_out/r2/scale8_n1.err
for N in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2/scale8_n$N.json 2> gpurun_out/r2/scale8_n$N.err
done
timeout 300 python -m pytest tests/test_multi.py -m gpu -q > gpurun_out/r2/scale8_multi_test.log 2>&1
tail -3 gpurun_out/r2/scale8_multi_test.log
tail -c 600 gpurun_out/r2/scale8_n8.json
