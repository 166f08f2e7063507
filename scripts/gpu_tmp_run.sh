mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2v_bench_2gpu.log 2> gpurun_out/r2v_bench_2gpu.err; echo "exit $?"
tail -c 1500 gpurun_out/r2v_bench_2gpu.log | head -c 700; echo; tail -5 gpurun_out/r2v_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2v_ref_2gpu.log 2> gpurun_out/r2v_ref_2gpu.err; echo "exit $?"; tail -c 400 gpurun_out/r2v_ref_2gpu.log
