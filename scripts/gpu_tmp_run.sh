mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r02_bench.log 2> gpurun_out/r02_bench.err; echo "exit $?"
grep '^{"metric' gpurun_out/r02_bench.log | tail -1 > gpurun_out/r02_bench_line.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.log 2>&1; tail -c 600 gpurun_out/r02_bench_ref.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_driver.log 2>&1; grep '^{"metric' gpurun_out/r02_bench_driver.log | tail -1 > gpurun_out/r02_bench_driver_line.json
tail -3 gpurun_out/r02_bench.err
