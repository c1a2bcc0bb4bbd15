python bench.py > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01e_ref.json 2> gpurun_out/bench_r01e_ref.err; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/bench_launches_e.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --render-steps 1 > gpurun_out/ncu_e1.log 2>&1; echo ncu1 rc=$?
