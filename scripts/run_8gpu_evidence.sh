set -u
OUT=gpurun_out
P="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $P --master-port 29511 scripts/run_split.py > $OUT/r02f_split_8gpu.json 2> $OUT/r02f_split_8gpu.err; echo "split rc=$? $(tail -1 $OUT/r02f_split_8gpu.json | cut -c1-700)"
timeout 400 $P --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > $OUT/r02f_bench_8gpu.json 2> $OUT/r02f_bench_8gpu.err; echo "bench8 rc=$? $(tail -1 $OUT/r02f_bench_8gpu.json | cut -c1-300)"
mkdir -p /tmp/cfg4 && timeout 600 python scripts/run_config4.py --dir /tmp/cfg4 --gpus all --procs 24 --check 6 > $OUT/r02f_config4_8gpu.json 2> $OUT/r02f_config4_8gpu.err; echo "cfg4 rc=$? $(cut -c1-600 $OUT/r02f_config4_8gpu.json)"; tail -2 $OUT/r02f_config4_8gpu.err
