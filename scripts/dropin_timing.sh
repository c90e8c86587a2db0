#!/bin/bash
# Stage timings of the segment_transfer drop-in executable next to the pure-CPU reference build on the same files
# (integration/_build/*, built where /root/reference exists).  Usage: scripts/dropin_timing.sh [folder]
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
D="${1:-/tmp/rsgpu_dropin_timing}"
rm -rf "$D"; mkdir -p "$D"
python "$ROOT/integration/make_dropin_case.py" "$D" > /dev/null
"$ROOT/integration/_build/pose_proposal_rsgpu" "$D/scan0.rsdb" "$D/scan1.ply" "$D/scan1_pp.rsdb" -v > "$D/pp.log"
cp "$ROOT/tests/golden/dropin_pp.bin" "$D/scan1_pp/scan1_pp.bin"
mkdir -p "$D/out_gpu" "$D/out_cpu"
for arm in gpu cpu; do
  exe="$ROOT/integration/_build/segment_transfer_rsgpu"; [ $arm = cpu ] && exe="$ROOT/integration/_build/segment_transfer_ref"
  s=$(date +%s%N)
  "$exe" "$D/scan1_pp.rsdb" -o "$D/out_$arm/scan1_st.rsdb" > "$D/st_$arm.log"
  e=$(date +%s%N)
  echo "== segment_transfer ($arm): $(( (e - s) / 1000000 )) ms wall"
  grep -E "finished in|done in|Done in|took|\(GPU\)" "$D/st_$arm.log" | grep -v "SIMULATED_ANNEALING: Iter\|Loading .ply\|Computing levels\|GREEDY STEP" | head -40
done
