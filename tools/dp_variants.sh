run() { echo "== $*"; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*\|"samples": [0-9]*\|"grad_exchange": "[a-z_]*"' | tr '\n' ' '; echo; }
run DAE_ALLREDUCE=nccl_graph DAE_BENCH_CLOCKS=0
run DAE_ALLREDUCE=multimem DAE_BENCH_CLOCKS=0
run DAE_ALLREDUCE=nccl_graph DAE_BENCH_CLOCKS=0 DAE_STAGE=0
run DAE_ALLREDUCE=nccl DAE_BENCH_CLOCKS=0
run DAE_ALLREDUCE=multimem DAE_BENCH_CLOCKS=1 DAE_STAGE=0
