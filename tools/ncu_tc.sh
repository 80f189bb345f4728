#!/bin/bash
# ncu evidence of the three dominant kernels on ONE B200 (through gpurun, from the repository root):  bash tools/ncu_tc.sh <tag>
#   launch list of one short bench run, then --set full (with source counters) of gather / view / ray of the second chunk, coarse + fine pass
set -u
T=${1:-vX}
O=gpurun_out
mkdir -p $O
QUICK="--no-cpu-baseline --no-costvolume --no-extras --no-ref-cuda --no-accuracy"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_ncu_launches_$T.csv \
    python bench.py --steps 1 --warmup 1 --e2e-steps 1 $QUICK > $O/ncu_l_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_gather_tc|k_view_tc2|k_ray_tc2' -s 6 -c 6 -f -o $O/r02_tc_$T \
    python bench.py --steps 1 --warmup 0 --e2e-steps 1 --rays 189440 $QUICK > $O/ncu_tc_$T.log 2>&1
ncu -i $O/r02_tc_$T.ncu-rep --page raw --csv > $O/r02_ncu_tc_${T}_raw.csv 2>/dev/null
ncu -i $O/r02_tc_$T.ncu-rep --page source --csv > $O/r02_ncu_tc_${T}_source.csv 2>/dev/null
ls -la $O | grep $T
rm -f $O/r02_tc_$T.ncu-rep
