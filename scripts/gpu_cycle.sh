# usage (on the GPU box, through gpurun): bash scripts/gpu_cycle.sh [tests] [bench] [fullbench] [refbench] [launches] [prof_sweep] [part] [partrun] [benchN] [smoke]
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline"
BP="$B --no-convergence-leg --no-batch-leg --no-labeled-leg"   # profiler runs: the headline path only
for what in "$@"; do
case $what in
tests) python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ;;
bench) $B --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python scripts/show_bench.py gpurun_out/bench.json; tail -3 gpurun_out/bench.err ;;
launches) $BP --steps 4 --warmup 3 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches.csv $BP --steps 4 --warmup 3 > gpurun_out/ncu.log 2>&1; echo "launches rc=$?" ;;
prof_sweep) $BP --steps 3 --warmup 3 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'sbm_vertex_kernel' -s 4 -c 1 -o gpurun_out/prof_sweep -f $BP --steps 3 --warmup 3 > gpurun_out/ncu_sweep.log 2>&1; echo "prof_sweep rc=$?" ;;
fullbench) python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "fullbench rc=$?"; python scripts/show_bench.py gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err ;;
refbench) python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "refbench rc=$?"; cut -c1-400 gpurun_out/bench_ref.json ;;
part) python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29517 tests/dist_check_part.py > gpurun_out/part.log 2>&1; echo "part rc=$?"; grep -v 'OMP_NUM_THREADS\|^\*\*\*\*\|^frame\|^$' gpurun_out/part.log | tail -45 ;;
benchN) python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus ${NGPU:-2} --steps ${STEPS:-50} --warmup 5 ${BARGS:-} > gpurun_out/bench_n${NGPU:-2}${TAG:-}.json 2> gpurun_out/bench_n${NGPU:-2}${TAG:-}.err; echo "benchN rc=$?"; python scripts/show_bench.py gpurun_out/bench_n${NGPU:-2}${TAG:-}.json; grep -v 'OMP_NUM_THREADS\|^\*\*\*\*\|^$' gpurun_out/bench_n${NGPU:-2}${TAG:-}.err | tail -15 ;;
partrun) python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29521 scripts/part_run.py --warmup 5 --steps ${STEPS:-50} ${BARGS:-} > gpurun_out/partrun${TAG:-}.json 2> gpurun_out/partrun${TAG:-}.err; echo "partrun rc=$?"; cut -c1-900 gpurun_out/partrun${TAG:-}.json; grep -v 'OMP_NUM_THREADS\|^\*\*\*\*\|^$' gpurun_out/partrun${TAG:-}.err | grep -i "error\|rank" | head -12 | cut -c1-300 ;;
smoke) python -c "import __graft_entry__ as g; g.smoke()" ;;
esac
done
