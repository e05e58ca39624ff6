P=r02h
mkdir -p gpurun_out
(timeout 700 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|skipped" | tail -60) > gpurun_out/${P}_pytest_gpu_full.log
timeout 500 python bench.py --steps 50 --warmup 5 --profile-out gpurun_out/${P}_gemm_table.json > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err
T2I_NVTX=1 timeout 600 ncu --nvtx --print-nvtx-rename kernel --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__grid_size --clock-control none -s 600 -c 470 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 2 --warmup 3 --only-resident --no-graphs > gpurun_out/${P}_ncu_bench.log 2>&1
tail -2 gpurun_out/${P}_pytest_gpu_full.log; head -c 250 gpurun_out/${P}_bench.json; echo; wc -l gpurun_out/${P}_launches.csv
