#!/bin/bash
# One parameterised GPU-box script (run under gpurun):  tools/gpu.sh <what> [args...]
#   tests [pytest args]     pytest -m gpu                       -> gpurun_out/pytest_gpu.log
#   bench [bench args]      python bench.py ...                 -> gpurun_out/bench.log (+ bench.json)
#   benchN N [bench args]   torchrun N ranks bench.py --gpus N  -> gpurun_out/bench_nN.log
#   launches [bench args]   ncu launch list of bench.py         -> gpurun_out/launches.csv
#   ncu KERNEL_REGEX [bench args]  ncu --set full of one kernel -> gpurun_out/prof.ncu-rep
#   sanitizer TOOL script.py       compute-sanitizer --tool TOOL -> gpurun_out/sanitizer_TOOL.log
set -u
mkdir -p gpurun_out
what=${1:-tests}; shift || true
case "$what" in
  tests)  python -m pytest tests -q -m gpu -x -s "$@" 2>&1 | tail -120 | tee gpurun_out/pytest_gpu.log ;;
  bench)  python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err ;;
  benchN) n=$1; shift
          python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
            bench.py --gpus $n "$@" > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
          tail -c 6000 gpurun_out/bench_n$n.json; tail -20 gpurun_out/bench_n$n.err ;;
  launches) # the first ~1000 launches are torch's weight generation + the warm-up steps
          ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1000 -c 400 --csv \
            --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-others --no-cpu-baseline "$@" \
            > gpurun_out/ncu_launches.log 2>&1; tail -3 gpurun_out/ncu_launches.log ;;
  ncu)    k=$1; shift   # e.g. "rowpair_matvec|qkv_attn": skips the first two steps of matching kernels
          ncu --set full --clock-control none --import-source on -k "regex:$k" --launch-skip 260 -c 10 -o gpurun_out/prof -f \
            python bench.py --steps 4 --warmup 3 --no-others --no-cpu-baseline "$@" > gpurun_out/ncu_full.log 2>&1
          tail -3 gpurun_out/ncu_full.log ;;
  sanitizer) tool=$1; shift
          timeout 600 compute-sanitizer --tool $tool --print-limit 2000 --error-exitcode 0 python "$@" > gpurun_out/sanitizer_$tool.log 2>&1
          echo "exit $?" >> gpurun_out/sanitizer_$tool.log
          grep -c "=========" gpurun_out/sanitizer_$tool.log; tail -12 gpurun_out/sanitizer_$tool.log ;;
  *) echo "unknown: $what"; exit 2 ;;
esac
