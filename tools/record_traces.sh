#!/bin/bash
# Records flush traces with the reference's own front end (built in place under
# oracle/_ref) + RenderContextCUDAImpl + the ABI call recorder. Needs
# /root/reference to have been built (python -c 'import __graft_entry__ as g; g.build()').
# usage: tools/record_traces.sh <outdir> <scene> [player args...]   (scene: gm:NAME | c1 | c2 | c3 | c5 | sriv:PATH)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$1"; SCENE="$2"; shift 2
NAME="${SCENE#gm:}"; NAME="${NAME#sriv:}"; NAME="$(basename "$NAME" .sriv)"
[ -n "$TRACE_NAME" ] && NAME="$TRACE_NAME"
mkdir -p "$OUT"
RIVECUDA_LIB="$ROOT/rive-runtime_b200/_build/librivecuda_trace.so" RIVECUDA_TRACE_OUT="$OUT/$NAME.rvct" \
  "$ROOT/rive-runtime_b200/_build/rive_cuda_player" --scene "$SCENE" "$@"
xz -f -9 -T0 "$OUT/$NAME.rvct"
ls -la "$OUT/$NAME.rvct.xz"
