#!/bin/bash
# A/B of library builds on ONE B200 (run through gpurun from the repository root):
#   gpurun --timeout 900 -- 'bash tools/ab_bench.sh tag base uni all base'
# Every name N is ab/libufo_N.so (built here with UFO_LIB_PATH / UFO_NVCC_EXTRA, see uforecon_b200/build.py); "all" is the in-tree
# default library.  Prints ms per depth map and per-kernel ms for each; JSON lines land in gpurun_out/ab_<tag>_<i>_<name>.json.
set -u
T=$1; shift
O=gpurun_out
mkdir -p $O
QUICK="--no-cpu-baseline --no-costvolume --no-extras --no-ref-cuda --no-accuracy --e2e-steps 1"
EXTRA=${AB_ARGS:-}
i=0
for n in "$@"; do
  i=$((i+1))
  if [ "$n" = "all" ]; then unset UFO_LIB_PATH; else export UFO_LIB_PATH=$PWD/ab/libufo_$n.so; fi
  f=$O/ab_${T}_${i}_$n.json
  python bench.py --steps ${AB_STEPS:-5} --warmup 3 $QUICK $EXTRA > $f 2> $O/ab_${T}_${i}_$n.err || { echo "$n FAILED"; tail -5 $O/ab_${T}_${i}_$n.err; continue; }
  python - "$f" "$n" <<'EOF'
import json, sys
j = json.load(open(sys.argv[1]))
steps = j["steps"]
rf = "  ".join(f"{k['name']} {k['ms'] / steps:.1f}" for k in j["kernels"][:4])
print(f"{sys.argv[2]:>10}: {j['ms_per_step']:.1f} ms/map  {j['value']:.4g} rays/s  clk {j['clocks'].get('sm_mhz')}  | ms/map: {rf}")
EOF
done
