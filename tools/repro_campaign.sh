#!/usr/bin/env bash
# Developer tool: ONE gpurun call that decides the open reproducibility item ("Run-to-run reproducibility").
#   (here, on the CPU container)   python tools/build_variant.py strict -DPDSC_STRICT_TMEM_WAR=1
#                                  python tools/build_variant.py opwait -DPDSC_ATTN_MMA_WAITS_OPERANDS=1
#                                  python tools/build_variant.py both -DPDSC_STRICT_TMEM_WAR=1 -DPDSC_ATTN_MMA_WAITS_OPERANDS=1
#   gpurun --timeout 900 -- 'bash tools/repro_campaign.sh'
# Writes gpurun_out/repro_<variant>.txt (determinism probe, 40 repetitions, features tap only), gpurun_out/bisect_default.txt
# (first differing per-layer tensor of the product build) and gpurun_out/bench_<variant>.json (same box, back to back).
set -u
mkdir -p gpurun_out
run_variant() {   # $1 = name, $2 = library path ("" = product build)
  if [ -n "$2" ]; then export POINTDSC_B200_LIB="$2"; else unset POINTDSC_B200_LIB; fi
  PDSC_TAPS=features PDSC_REPS=40 timeout 120 python tools/determinism_probe.py > "gpurun_out/repro_$1.txt" 2>&1
  tail -2 "gpurun_out/repro_$1.txt"
  timeout 200 python bench.py --steps 10 --warmup 3 > "gpurun_out/bench_$1.json" 2> "gpurun_out/bench_$1.err"
  python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
    print(sys.argv[1], round(d["value"]), "sets/s", d["ms_per_step"], "ms/step, attention", d["roofline"]["launch_ms"], "ms, chain+rest", d["stages"]["linear"]["ms_per_step"])
except Exception as e:
    print(sys.argv[1], "bench failed:", e)
PY
}
run_variant default ""
for v in strict opwait both; do
  [ -f "tools/bin/lib_$v.so" ] && run_variant "$v" "$PWD/tools/bin/lib_$v.so"
done
unset POINTDSC_B200_LIB
PDSC_REPS=240 timeout 300 python tools/race_bisect.py > gpurun_out/bisect_default.txt 2>&1
tail -15 gpurun_out/bisect_default.txt
