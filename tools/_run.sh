set -u
OUT=gpurun_out/s3f; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; nproc > $OUT/nproc.txt; free -g > $OUT/mem.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 5 --warmup 3 > $OUT/bench4.json 2> $OUT/bench4.err
echo "bench4 rc=$?"; tail -3 $OUT/bench4.err; cut -c1-200 $OUT/bench4.json
