#!/bin/bash
# Evidence set of a build (run on the GPU box through gpurun): test log, bench line + per-launch table, ncu launch list by
# layer (duration, tensor-pipe %, DRAM bytes), sanitizer logs.  Outputs under gpurun_out/ with the prefix $1.
P=${1:-r02}
mkdir -p gpurun_out
(timeout 700 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|skipped" | tail -60) > gpurun_out/${P}_pytest_gpu_full.log
timeout 500 python bench.py --steps 50 --warmup 5 --profile-out gpurun_out/${P}_gemm_table.json > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err
T2I_NVTX=1 timeout 600 ncu --nvtx --print-nvtx-rename kernel --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__grid_size --clock-control none -s 600 -c 460 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 2 --warmup 3 --only-resident --no-graphs > gpurun_out/${P}_ncu_bench.log 2>&1
# one `--set full` capture of GEMM-type launches of the second step (D run), named by layer; summarised to text on the box
T2I_NVTX=1 timeout 900 ncu --set full --nvtx --print-nvtx-rename kernel -k regex:"conv_gemm|wgrad_gemm|deconv_img" --clock-control none -s 130 -c 48 -f -o /tmp/${P}_full python bench.py --steps 1 --warmup 1 --only-resident --no-graphs > gpurun_out/${P}_ncu_full.log 2>&1
python tools/summarize_ncu.py full /tmp/${P}_full.ncu-rep > gpurun_out/${P}_ncu_full_by_layer.txt 2>> gpurun_out/${P}_ncu_full.log
python tools/bench_img.py --reps 9 --out gpurun_out/${P}_bench_img_kernels.json > gpurun_out/${P}_bench_img.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/${P}_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/${P}_sanitizer_racecheck.log
tail -3 gpurun_out/${P}_pytest_gpu_full.log; head -c 600 gpurun_out/${P}_bench.json; echo; tail -2 gpurun_out/${P}_sanitizer_memcheck.log; tail -4 gpurun_out/${P}_sanitizer_racecheck.log; wc -l gpurun_out/${P}_launches.csv
