#!/usr/bin/env bash
# TEST INFRASTRUCTURE. Builds the UNMODIFIED reference (headers under /root/reference/gato, compiled where
# they lie) into oracle/_ref/libgref_<plant>_N<N>_<fast|ieee>.so through oracle/ref_harness.cu.
#   fast = the reference's own flags (CMakeLists.txt:20-22: -O3 -use_fast_math -DNDEBUG), arch swapped to sm_100
#   ieee = same without -use_fast_math (isolates fast-math noise from algorithmic parity)
# Usage: oracle/build_ref.sh                 the libraries the GPU box uses: bench.py's (iiwa14 N=32, fast) and the live parity tests' (iiwa14 / indy7 N=32, ieee)
#        oracle/build_ref.sh all             every library oracle/gen_golden.py needs, in parallel
#        oracle/build_ref.sh pin             the IEEE builds the tools/pin_*.py comparisons use
#        oracle/build_ref.sh plant N batches mode [...]
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${GATO_REFERENCE_DIR:-/root/reference}
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -d "$REF/gato" ]; then echo "reference not present at $REF; keeping prebuilt files in $OUT"; exit 0; fi

build_one() {
  local plant=$1 N=$2 batches=$3 mode=$4
  local def; if [ "$plant" = iiwa14 ]; then def=-DPLANT_IIWA14=1; else def=-DPLANT_INDY7=1; fi
  local fm=""; if [ "$mode" = fast ]; then fm="-use_fast_math -DGREF_FAST_MATH=1"; fi
  local out="$OUT/libgref_${plant}_N${N}_${mode}.so"
  if [ -f "$out" ] && [ "$out" -nt "$HERE/ref_harness.cu" ] && [ -z "${GREF_FORCE:-}" ]; then echo "up to date: $out"; return 0; fi
  # nvcc splits -D values at commas, so the batch list goes through a generated header
  local cfg="$OUT/cfg_${plant}_N${N}_${mode}.h"
  echo "#define GREF_BATCHES $batches" > "$cfg"
  nvcc -std=c++17 -O3 $fm -DNDEBUG -Xcompiler -fPIC -shared -w \
       -gencode arch=compute_100,code=sm_100 \
       -I"$REF/gato" -DKNOT_POINTS=$N $def -include "$cfg" \
       "$HERE/ref_harness.cu" -o "$out.tmp"
  mv "$out.tmp" "$out"
  echo "built $out"
}

if [ $# -ge 4 ]; then
  while [ $# -ge 4 ]; do build_one "$1" "$2" "$3" "$4"; shift 4; done
  exit 0
fi

# no arguments: the libraries the GPU box needs at run time (bench.py's reference_gpu row: iiwa14, N=32, B=512, the reference's flags; and
# the IEEE build tests/test_gpu_reference_live.py compares the CUDA path with).
# "all": the whole matrix oracle/gen_golden.py uses to mint tests/golden/ (BASELINE.json configs + the bench shape; ~15 min on 6 jobs).
# "pin": the IEEE builds tools/pin_cost_gradient.py and tools/pin_whole_solve.py compare the oracle with (15 horizons / plants).
if [ "${1:-}" = pin ]; then
  JOBS=${GREF_JOBS:-6}
  { for n in 3 4 6 8 9 10 11 12 16 32 64; do echo "iiwa14 $n 1,16 ieee"; done; echo "iiwa14 128 8 ieee"; for n in 8 16 32; do echo "indy7 $n 16 ieee"; done; } | xargs -P "$JOBS" -L 1 "$0"
  exit 0
fi
if [ "${1:-}" != all ]; then
  # bench.py's reference_gpu row (the reference's own flags) and the live parity test's IEEE build, side by side
  build_one iiwa14 32 16,128,512 fast &
  p1=$!
  build_one iiwa14 32 16,128 ieee &
  p2=$!
  # indy7: the live stage-chain parity test (tests/test_gpu_reference_live.py) and the stage fixture (gen_golden.py --stages-only)
  build_one indy7 32 16 ieee &
  p3=$!
  rc=0
  wait $p1 || rc=1
  wait $p2 || rc=1
  wait $p3 || rc=1
  exit $rc
fi
JOBS=${GREF_JOBS:-6}
cat <<LIST | xargs -P "$JOBS" -L 1 "$0"
iiwa14 32 16,128,512 fast
iiwa14 128 8,1024 fast
iiwa14 32 16,128 ieee
indy7 32 16,512 fast
iiwa14 8 1,16 fast
iiwa14 8 1,16 ieee
indy7 32 16 ieee
indy7 16 16 fast
LIST
