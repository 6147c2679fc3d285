#!/usr/bin/env bash
# Developer tool: ONE gpurun call that checks run-to-run reproducibility of the product build (and of any variant libraries
# built with tools/build_variant.py into tools/bin/lib_<name>.so):
#   gpurun --timeout 900 -- 'bash tools/repro_campaign.sh'
# Writes gpurun_out/repro_<variant>.txt (determinism probe, 100 repetitions, every other one perturbed) and
# gpurun_out/bisect_default.txt (first differing per-layer tensor, if any).  History: round 2 used this script with the
# -DPDSC_STRICT_TMEM_WAR / -DPDSC_ATTN_MMA_WAITS_OPERANDS diagnostic builds of round 1 (profiles/r02_determinism_campaign.txt);
# those macros are gone — the waits they guarded are unconditional now and the actual bug was a counted mbarrier (tc_common.cuh).
set -u
mkdir -p gpurun_out
run_variant() {   # $1 = name, $2 = library path ("" = product build)
  if [ -n "$2" ]; then export POINTDSC_B200_LIB="$2"; else unset POINTDSC_B200_LIB; fi
  PDSC_TAPS=features PDSC_REPS=100 timeout 300 python tools/determinism_probe.py > "gpurun_out/repro_$1.txt" 2>&1
  echo "$1: $(tail -1 "gpurun_out/repro_$1.txt")"
}
run_variant default ""
for lib in tools/bin/lib_*.so; do
  [ -f "$lib" ] || continue
  name=$(basename "$lib" .so); name=${name#lib_}
  run_variant "$name" "$PWD/$lib"
done
unset POINTDSC_B200_LIB
PDSC_REPS=120 timeout 300 python tools/race_bisect.py > gpurun_out/bisect_default.txt 2>&1
tail -3 gpurun_out/bisect_default.txt
