set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_gpu_tests.txt
python bench.py > gpurun_out/r02_bench_c2_4k.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_c2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_bench_under_ncu.log 2>&1
for k in raster_spans setup_patches sort_tiles scatter_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r02_prof_$k -f python tools/quickbench.py tests/golden/c2_4k.rvct.xz > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
cat gpurun_out/r02_gpu_tests.txt; cat gpurun_out/r02_bench_c2_4k.json
