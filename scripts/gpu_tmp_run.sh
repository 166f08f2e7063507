mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2q_bench.log 2> gpurun_out/r2q_bench.err; echo "exit $?" >> gpurun_out/r2q_bench.log
tail -c 6000 gpurun_out/r2q_bench.log; tail -5 gpurun_out/r2q_bench.err
