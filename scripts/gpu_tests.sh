set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err
echo "bench rc=$?"; cat gpurun_out/bench_first.json; tail -5 gpurun_out/bench_first.err
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
