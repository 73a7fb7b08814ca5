set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
python tools/prof_driver.py knn 3 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 3 --warmup 3 --extras 0 > gpurun_out/ncu_l.log 2>&1
for op in knn ball three_nn knn_direct nn1; do
  ncu --set full --clock-control none --import-source on -k regex:search_kernel --launch-skip 2 -c 1 -o gpurun_out/r01f_$op -f python tools/prof_driver.py $op 3 > gpurun_out/ncu_$op.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -8
