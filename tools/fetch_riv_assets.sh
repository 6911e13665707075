#!/bin/bash
# Copies the .riv test assets the GPU end-to-end tests use from the reference tree into
# tests/_riv_assets/ (git-ignored: reference assets are not committed to this repo).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
mkdir -p "$ROOT/tests/_riv_assets"
# off_road_car, bullet_man: the end-to-end tests; the rest: vector-only files the device front end
# can draw whole (tests/test_front_end_gpu.py compares it with the reference front end on them;
# tape and image_fit_alignment_2 hold image meshes, juice gradients and clip paths).
for n in off_road_car bullet_man shapetest fix_rectangle follow_path_solos trim_path_linear magic_alley_db_reduced_export \
         nested_artboard_opacity lock_icon_demo follow_path_shapes solos_collapse_tests group_effect tape image_fit_alignment_2 juice; do
  cp "/root/reference/tests/unit_tests/assets/$n.riv" "$ROOT/tests/_riv_assets/"
done
# feathers: drawn by delegating the feather draws to the reference front end (several flushes per frame)
mkdir -p "$ROOT/tests/_riv_assets/feathers"
for n in coin path_effect_with_feathers ai_assitant bankcard rewards_demo hunter_x_demo; do
  cp "/root/reference/tests/unit_tests/assets/$n.riv" "$ROOT/tests/_riv_assets/feathers/"
done
ls -la "$ROOT/tests/_riv_assets"
