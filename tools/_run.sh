set -u
OUT=gpurun_out/s3e; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench2.json 2> $OUT/bench2.err
echo "bench2 rc=$?"; tail -3 $OUT/bench2.err; cut -c1-300 $OUT/bench2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $OUT/ref2.json 2> $OUT/ref2.err
echo "ref2 rc=$?"; cut -c1-200 $OUT/ref2.json
timeout 600 python bench.py --gpus 2 --single-process --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench2sp.json 2> $OUT/bench2sp.err
echo "bench2sp rc=$?"; tail -2 $OUT/bench2sp.err; cut -c1-200 $OUT/bench2sp.json
