#!/bin/bash
# Reduced evidence set of a build (no ncu passes): test log, bench line + per-launch table, image-kernel micro-benchmarks,
# sanitizer logs.  Outputs under gpurun_out/ with the prefix $1.  (tools/evidence_run.sh adds the ncu launch list and full pass.)
P=${1:-r02}
mkdir -p gpurun_out
(timeout 700 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|skipped" | tail -60) > gpurun_out/${P}_pytest_gpu_full.log
timeout 500 python bench.py --steps 50 --warmup 5 --profile-out gpurun_out/${P}_gemm_table.json > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err
python tools/bench_img.py --reps 9 --out gpurun_out/${P}_bench_img_kernels.json > gpurun_out/${P}_bench_img.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/${P}_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/${P}_sanitizer_racecheck.log
tail -3 gpurun_out/${P}_pytest_gpu_full.log; head -c 300 gpurun_out/${P}_bench.json; echo; tail -2 gpurun_out/${P}_sanitizer_memcheck.log; tail -4 gpurun_out/${P}_sanitizer_racecheck.log; tail -14 gpurun_out/${P}_bench_img.log
