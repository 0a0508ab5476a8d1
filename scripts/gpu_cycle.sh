# usage: bash scripts/gpu_cycle.sh [tests] [bench] [launches] [prof_sweep] [prof_mstep]
mkdir -p gpurun_out
for what in "$@"; do
case $what in
tests) python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ;;
bench) python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err ;;
launches) python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "launches rc=$?" ;;
prof_sweep) python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sbm_vertex_kernel -s 4 -c 1 -o gpurun_out/prof_sweep -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_sweep.log 2>&1; echo "prof_sweep rc=$?" ;;
prof_mstep) python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:sbm_mstep_kernel -s 3 -c 1 -o gpurun_out/prof_mstep -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mstep.log 2>&1; echo "prof_mstep rc=$?" ;;
esac
done
