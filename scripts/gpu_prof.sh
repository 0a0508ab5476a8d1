set -x
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sbm_vertex_kernel -s 4 -c 1 -o gpurun_out/prof_sweep -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_sweep.log 2>&1
echo "ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sbm_mstep_kernel -s 3 -c 1 -o gpurun_out/prof_mstep -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mstep.log 2>&1
echo "ncu rc=$?"
