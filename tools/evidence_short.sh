set -u
T=v38; O=gpurun_out; mkdir -p $O
for t in memcheck racecheck; do timeout 120 compute-sanitizer --tool $t python tools/sanitize_case.py 3 > $O/san_${T}_${t}_nv3.log 2>&1; echo "$t nv3: $(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $O/san_${T}_${t}_nv3.log | tail -1)"; done
python bench.py --steps 20 --warmup 5 > $O/r02_bench_${T}_1gpu.json 2> $O/bench_$T.err; python tools/bench_summary.py $O/r02_bench_${T}_1gpu.json | head -1 | cut -c1-80
bash tools/ncu_tc.sh $T > /dev/null 2>&1; ls $O | grep $T | head
python -m pytest tests -q -m gpu -s 2>&1 | grep -v '^$' | tail -70 > $O/r02_gpu_tests_$T.log; tail -1 $O/r02_gpu_tests_$T.log
