export RIVECUDA_LIB=rive-runtime_b200/_build/librivecuda.so
P=rive-runtime_b200/_build/rive_cuda_player
for a in off_road_car bullet_man; do
  for mode in "" "--budget-ms 0" "--gpu-front-end"; do
    echo "== $a [$mode]"; $P --scene riv:tests/_riv_assets/$a.riv --frames 600 $mode 2>&1 | tail -1
  done
  echo "== $a null backend"; $P --scene riv:tests/_riv_assets/$a.riv --frames 600 --null-backend 2>&1 | tail -1
done
