#!/bin/bash
# Round-2 compute-sanitizer pass over the kernels added this round (span rasteriser, device front end
# with clockwise fills / gradients / clip paths / images, exact rasteriser), through the scene player.
export RIVECUDA_LIB=rive-runtime_b200/_build/librivecuda.so
P=rive-runtime_b200/_build/rive_cuda_player
OUT=gpurun_out/r02_sanitizer.txt
: > $OUT
for scene in f1w f1g f1p f1i; do
  echo "== memcheck $scene --gpu-front-end (1280x720, 1200 paths)" >> $OUT
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 $P --scene $scene --gpu-front-end --budget-ms 0 --width 1280 --height 720 --paths 1200 2>&1 | tail -4 >> $OUT
done
echo "== memcheck c2 1920x1080 3000 paths (span rasteriser)" >> $OUT
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 $P --scene c2 --width 1920 --height 1080 --paths 3000 2>&1 | tail -4 >> $OUT
echo "== racecheck c2 1280x720 1000 paths (span rasteriser: shared-memory delta planes)" >> $OUT
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 $P --scene c2 --width 1280 --height 720 --paths 1000 2>&1 | tail -4 >> $OUT
echo "== racecheck f1p --gpu-front-end 640x360 400 paths" >> $OUT
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 $P --scene f1p --gpu-front-end --budget-ms 0 --width 640 --height 360 --paths 400 2>&1 | tail -4 >> $OUT
echo "== synccheck c2 1280x720 1000 paths" >> $OUT
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 3 $P --scene c2 --width 1280 --height 720 --paths 1000 2>&1 | tail -4 >> $OUT
cat $OUT
