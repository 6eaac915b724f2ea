#!/bin/bash
# How the round's evidence under profiles/ is produced (run on a B200 box through gpurun; outputs land in gpurun_out/ and are then
# copied to profiles/ as r02_*):   gpurun --timeout 3000 -- 'bash tools/final_profile.sh'
set -x
mkdir -p gpurun_out
# 1. the whole GPU suite
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "^FAILED|passed|failed|^E  " | cut -c1-300 | head -20
# 2. bench lines: default (C2) and the other BASELINE.json configs
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_C2.json 2> gpurun_out/r02_bench_C2.err; tail -c 200 gpurun_out/r02_bench_C2.err
for c in C1 C3 C4 C5; do
  timeout 400 python bench.py --config $c --steps 20 --warmup 5 --cpu-steps 1 > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err; tail -c 200 gpurun_out/r02_bench_$c.err
done
# 3. kernel timeline of a replayed step (CUPTI through torch.profiler)
timeout 200 python tools/step_timeline.py > gpurun_out/r02_timeline.json 2> gpurun_out/timeline.err; tail -c 200 gpurun_out/timeline.err
if [ "$1" = "--ncu" ]; then
  # 4. launch list and full captures of one eagerly launched step; transform-sized K1, row kernel vs hot-rows kernel
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-fit-api > gpurun_out/ncu_launches.log 2>&1
  timeout 500 ncu --set full --clock-control none --import-source on \
      -k regex:"gemm_bf16x3|triplet_batch_all|encode_fwd|encode_bwd|optimizer_kernel|batch_prepare|step_finalize" -s 80 -c 24 -o gpurun_out/r02_step \
      python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-fit-api > gpurun_out/ncu_step.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"encode_fwd" -c 4 -o gpurun_out/r02_transform \
      python tools/bench_transform.py 100000 1 --configs C2 --variants row,hot_g8 --warmup 1 > gpurun_out/ncu_transform.log 2>&1
  timeout 300 python tools/bench_transform.py 100000 10 > gpurun_out/r02_transform_variants.json 2>/dev/null
fi
