set -x
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_C2.json 2> gpurun_out/r02_bench_C2.err; tail -c 200 gpurun_out/r02_bench_C2.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_C2_b.json 2> gpurun_out/r02_bench_C2.err
for c in C1 C3 C4 C5; do timeout 400 python bench.py --config $c --steps 20 --warmup 5 --cpu-steps 1 > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err; tail -c 200 gpurun_out/r02_bench_$c.err; done
