# End-of-round measurement pass (round 2): tests, bench lines, ncu launch list and one `--set full` capture per hot kernel.
# Run on the GPU box from the repo root: gpurun --timeout 2400 -- 'bash tools/final_pass_r02.sh'
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/r02_final_pytest_gpu.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r02_final_bench_all_configs.json 2> $O/r02_final_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_final_bench_reference.json 2>> $O/r02_final_bench.err
timeout 300 python bench.py --steps 5 --warmup 3 --solver subgrad --no-cpu-baseline --no-other-configs > $O/r02_final_bench_subgrad.json 2>> $O/r02_final_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_final_launches_1024pairs.csv python bench.py --steps 2 --warmup 1 --no-other-configs --no-e2e --no-cpu-baseline > $O/r02_final_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -o $O/r02_final_solve_full -f python tools/one_solve.py 1024 gn 10 2 > $O/r02_final_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sobel_nms -c 1 -o $O/r02_final_nms_full -f python tools/one_prepare.py 1024 > $O/r02_final_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:canny_kernel -c 1 -o $O/r02_final_canny_full -f python tools/one_prepare.py 296 > $O/r02_final_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edt_pack -c 1 -o $O/r02_final_edt_full -f python tools/one_prepare.py 296 > $O/r02_final_ncu4.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:solve_kernel -s 1 -c 1 --csv --log-file $O/r02_final_solve_kernel_dram_subgrad50_1024pairs.csv python tools/one_solve.py 1024 subgrad 50 2 > /dev/null 2>&1
for k in solve nms canny edt; do ncu -i $O/r02_final_${k}_full.ncu-rep --page raw --csv > $O/r02_final_${k}_raw.csv 2>/dev/null; python profiles/ncu_extract.py $O/r02_final_${k}_raw.csv > $O/r02_final_ncu_full_${k}.txt; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02_final_smoke.txt 2>&1
cat $O/r02_final_pytest_gpu.txt; tail -1 $O/r02_final_smoke.txt; head -c 400 $O/r02_final_bench_all_configs.json; echo; head -c 300 $O/r02_final_bench_subgrad.json; echo; head -c 400 $O/r02_final_bench_reference.json
