set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/v_bench_n2.json 2> gpurun_out/v_bench_n2.log
head -c 900 gpurun_out/v_bench_n2.json; echo; tail -n 3 gpurun_out/v_bench_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | head -c 300; echo
python scripts/fs_bench.py 200000 2000 2>&1 | tail -3 > gpurun_out/u_fs.log; cat gpurun_out/u_fs.log
