#!/bin/bash
# Regenerates the tests/golden/*.paths.xz dumps (the RawPaths + matrices + paints of a frame's plain
# fill and stroke draws: the input of rivecuda_front_end_paths) together with the flush trace of the
# same frame, and checks that the trace equals the committed one. Needs the player built against
# /root/reference (python -c 'import __graft_entry__ as g; g.build()').
# usage: tools/record_path_dumps.sh [outdir]     (default: tests/golden)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="${1:-$ROOT/tests/golden}"
TMP="$(mktemp -d)"
SCENES="f1 s1 gm:trickycubicstrokes gm:trickycubicstrokes_roundcaps gm:emptystroke gm:strokes3 gm:labyrinth_round
gm:labyrinth_square gm:zero_control_stroke gm:zerolinestroke gm:OverStroke gm:bevel180strokes gm:roundjoinstrokes
gm:widebuttcaps gm:lots_of_tess_spans_stroke gm:beziers gm:CubicStroke gm:inner_join_geometry gm:teenyStrokes gm:quadcap gm:strokefill gm:zeroPath"
for SCENE in $SCENES; do
  NAME="${SCENE#gm:}"
  RIVECUDA_LIB="$ROOT/rive-runtime_b200/_build/librivecuda_trace.so" RIVECUDA_TRACE_OUT="$TMP/$NAME.rvct" \
    "$ROOT/rive-runtime_b200/_build/rive_cuda_player" --scene "$SCENE" --budget-ms 0 --dump-paths "$TMP/$NAME.paths" > /dev/null
  if [ -f "$ROOT/tests/golden/$NAME.rvct.xz" ] && ! xz -dc "$ROOT/tests/golden/$NAME.rvct.xz" | cmp -s - "$TMP/$NAME.rvct"; then
    echo "WARNING: $NAME: the recorded trace differs from tests/golden/$NAME.rvct.xz"
  fi
  xz -9 -c "$TMP/$NAME.paths" > "$OUT/$NAME.paths.xz"
  [ -f "$OUT/$NAME.rvct.xz" ] || xz -9 -c "$TMP/$NAME.rvct" > "$OUT/$NAME.rvct.xz"
  echo "$NAME: $(stat -c %s "$OUT/$NAME.paths.xz") bytes"
done
# c2 at 4K (BASELINE.json configs[1])
RIVECUDA_LIB="$ROOT/rive-runtime_b200/_build/librivecuda_trace.so" RIVECUDA_TRACE_OUT="$TMP/c2_4k.rvct" \
  "$ROOT/rive-runtime_b200/_build/rive_cuda_player" --scene c2 --budget-ms 0 --dump-paths "$TMP/c2_4k.paths" > /dev/null
xz -dc "$ROOT/tests/golden/c2_4k.rvct.xz" | cmp -s - "$TMP/c2_4k.rvct" || echo "WARNING: c2_4k trace differs"
xz -9 -c "$TMP/c2_4k.paths" > "$OUT/c2_4k.paths.xz"
rm -rf "$TMP"
