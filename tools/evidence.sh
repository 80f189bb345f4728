#!/bin/bash
# Round evidence on ONE B200 (run through gpurun from the repository root): full GPU test suite, the driver-style bench line
# with all extras, the reference arm, the NV / view-set variants, the ncu launch list and the ncu --set full captures that
# profiles/ summarises.  Everything lands in gpurun_out/ with the tag given as $1.
#   gpurun --timeout 2400 -- 'bash tools/evidence.sh v26'
set -u
T=${1:-vX}
O=gpurun_out
mkdir -p $O
QUICK="--no-cpu-baseline --no-costvolume --no-extras --no-ref-cuda"

python -m pytest tests -q -m gpu -s 2>&1 | grep -v '^$' | tail -60 > $O/r02_gpu_tests_$T.log
tail -3 $O/r02_gpu_tests_$T.log

python bench.py --steps 20 --warmup 5 > $O/r02_bench_${T}_1gpu.json 2> $O/bench_$T.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_${T}_reference_arm.json 2>> $O/bench_$T.err
python bench.py --nv 5 --steps 5 --warmup 3 $QUICK > $O/r02_bench_${T}_nv5.json 2>> $O/bench_$T.err
python bench.py --nv 10 --steps 3 --warmup 3 $QUICK > $O/r02_bench_${T}_nv10.json 2>> $O/bench_$T.err
python bench.py --views favorable --steps 5 --warmup 3 $QUICK > $O/r02_bench_${T}_fav.json 2>> $O/bench_$T.err
for f in 1gpu nv5 nv10 fav; do echo "== $f"; python tools/bench_summary.py $O/r02_bench_${T}_$f.json 2>/dev/null | head -5; done

# launch list of one short bench run (per-launch times are cold-cache and serialised: the SHARES are what must agree)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_ncu_launches_$T.csv \
    python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-accuracy $QUICK > $O/ncu_l_$T.log 2>&1
# full capture of the three dominant kernels at the default chunk: third chunk of the map, coarse + fine pass
ncu --set full --clock-control none --import-source on -k 'regex:k_gather_tc|k_view_tc2|k_ray_tc2' -s 12 -c 6 -f -o $O/r02_tc_$T \
    python bench.py --steps 1 --warmup 0 --e2e-steps 1 --rays 189440 --no-accuracy $QUICK > $O/ncu_tc_$T.log 2>&1
ncu -i $O/r02_tc_$T.ncu-rep --page raw --csv > $O/r02_ncu_tc_${T}_raw.csv 2>/dev/null
ncu -i $O/r02_tc_$T.ncu-rep --page source --csv > $O/r02_ncu_tc_${T}_source.csv 2>/dev/null
rm -f $O/r02_tc_$T.ncu-rep
# small kernels of the render path and kernel 1
ncu --set full --clock-control none -k 'regex:k_coarse_z|k_render|k_importance|k_ray_setup' -s 10 -c 10 -f -o $O/r02_small_$T \
    python bench.py --steps 1 --warmup 0 --e2e-steps 1 --rays 189440 --no-accuracy $QUICK > $O/ncu_small_$T.log 2>&1
ncu -i $O/r02_small_$T.ncu-rep --page raw --csv > $O/r02_ncu_small_${T}_raw.csv 2>/dev/null
rm -f $O/r02_small_$T.ncu-rep
ncu --set full --clock-control none -k regex:k_costvol -s 1 -c 5 -f -o $O/r02_cv_$T python tools/costvol_profile.py --iters 1 > $O/ncu_cv_$T.log 2>&1
ncu -i $O/r02_cv_$T.ncu-rep --page raw --csv > $O/r02_ncu_costvol_${T}_raw.csv 2>/dev/null
rm -f $O/r02_cv_$T.ncu-rep
python tools/costvol_profile.py > $O/r02_costvol_${T}_nv3.json 2>> $O/bench_$T.err
python tools/costvol_profile.py --nv 5 > $O/r02_costvol_${T}_nv5.json 2>> $O/bench_$T.err
python tools/costvol_profile.py --nv 10 --iters 1 > $O/r02_costvol_${T}_nv10.json 2>> $O/bench_$T.err
python tools/encoder_bench.py > $O/r02_encoder_seconds_$T.json 2>> $O/bench_$T.err
ls -la $O | grep $T
