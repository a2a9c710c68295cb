set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r01_final_pytest_gpu.txt
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r01_final_bench_gn.json 2> gpurun_out/r01_final_bench_gn.err
timeout 300 python bench.py --steps 5 --warmup 3 --solver subgrad --no-cpu-baseline > gpurun_out/r01_final_bench_subgrad.json 2>> gpurun_out/r01_final_bench_gn.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_final_bench_reference.json 2>> gpurun_out/r01_final_bench_gn.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_final_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r01_final_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o gpurun_out/r01_final_solve_full -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/r01_final_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none -c 14 -o gpurun_out/r01_final_all_full -f python bench.py --pairs 296 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/r01_final_ncu2.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:solve_kernel -c 1 --csv --log-file gpurun_out/r01_final_solve_dram_subgrad.csv python bench.py --steps 1 --warmup 0 --solver subgrad --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 300 python tools/bench_photometric.py 1024 6 > gpurun_out/r01_final_config3.json 2>/dev/null
timeout 300 python tools/bench_sequence.py 64 16 > gpurun_out/r01_final_config4.json 2>/dev/null
timeout 300 python tools/bench_rgbd.py 1024 > gpurun_out/r01_final_rgbd.json 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r01_final_smoke.txt 2>&1
cat gpurun_out/r01_final_pytest_gpu.txt; tail -1 gpurun_out/r01_final_smoke.txt; head -c 300 gpurun_out/r01_final_bench_gn.json; echo; head -c 300 gpurun_out/r01_final_bench_subgrad.json; echo; head -c 400 gpurun_out/r01_final_bench_reference.json
