#!/bin/bash
# Records a flush trace of every .sriv silver of the reference (tests/unit_tests/silvers) at
# 1920x1080 with the reference's own front end + the ABI recorder, xz-compressed, into <outdir>.
# Needs /root/reference and the built player (python -c 'import __graft_entry__ as g; g.build()').
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$1"; mkdir -p "$OUT"
for f in /root/reference/tests/unit_tests/silvers/*.sriv; do
  n=$(basename "$f" .sriv)
  RIVECUDA_LIB="$ROOT/rive-runtime_b200/_build/librivecuda_trace.so" RIVECUDA_TRACE_OUT="$OUT/$n.rvct" \
    timeout 120 "$ROOT/rive-runtime_b200/_build/rive_cuda_player" --scene "sriv:$f" --frames 0 > /dev/null
done
xz -T0 -1 -f "$OUT"/*.rvct
ls "$OUT" | wc -l
