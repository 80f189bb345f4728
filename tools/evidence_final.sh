#!/bin/bash
# End-of-round evidence on ONE B200, most important first (through gpurun from the repository root):  bash tools/evidence_final.sh <tag>
set -u
T=${1:-vX}
O=gpurun_out
mkdir -p $O
QUICK="--no-cpu-baseline --no-costvolume --no-extras --no-ref-cuda"
t0=$(date +%s)
python -m pytest tests -q -m gpu -s 2>&1 | grep -v '^$' | tail -70 > $O/r02_gpu_tests_$T.log
tail -2 $O/r02_gpu_tests_$T.log; echo "[$(( $(date +%s) - t0 )) s] tests"
python bench.py --steps 20 --warmup 5 > $O/r02_bench_${T}_1gpu.json 2> $O/bench_$T.err
python tools/bench_summary.py $O/r02_bench_${T}_1gpu.json | head -8; echo "[$(( $(date +%s) - t0 )) s] bench"
bash tools/ncu_tc.sh $T > /dev/null 2>&1; echo "[$(( $(date +%s) - t0 )) s] ncu"
python bench.py --nv 10 --steps 3 --warmup 3 $QUICK > $O/r02_bench_${T}_nv10.json 2>> $O/bench_$T.err
python bench.py --nv 5 --steps 5 --warmup 3 $QUICK > $O/r02_bench_${T}_nv5.json 2>> $O/bench_$T.err
python bench.py --views favorable --steps 5 --warmup 3 $QUICK > $O/r02_bench_${T}_fav.json 2>> $O/bench_$T.err
for f in nv10 nv5 fav; do echo "== $f"; python tools/bench_summary.py $O/r02_bench_${T}_$f.json 2>/dev/null | head -1; done; echo "[$(( $(date +%s) - t0 )) s] variants"
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_${T}_reference_arm.json 2>> $O/bench_$T.err
tail -c 400 $O/r02_bench_${T}_reference_arm.json; echo; echo "[$(( $(date +%s) - t0 )) s] reference arm"
