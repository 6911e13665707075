#!/bin/bash
# Copies the .riv test assets the GPU end-to-end tests use from the reference tree into
# tests/_riv_assets/ (git-ignored: reference assets are not committed to this repo).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
mkdir -p "$ROOT/tests/_riv_assets"
for n in off_road_car bullet_man; do cp "/root/reference/tests/unit_tests/assets/$n.riv" "$ROOT/tests/_riv_assets/"; done
ls -la "$ROOT/tests/_riv_assets"
