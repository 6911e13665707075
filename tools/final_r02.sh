python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_gpu_tests_final.txt
ncu --set full --clock-control none --import-source on -k regex:raster_spans -s 2 -c 1 -o gpurun_out/r02_prof_raster_spans -f python tools/quickbench.py tests/golden/c2_4k.rvct.xz > /dev/null 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_final_quick.json 2>/dev/null
cat gpurun_out/r02_gpu_tests_final.txt; cut -c1-300 gpurun_out/r02_bench_final_quick.json
