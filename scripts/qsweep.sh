# sweep-kernel bandwidth across q (sbm.jl model), N = 1M, degree 16
for q in 2 3 4 5 6 8 10 12 16; do
  python bench.py --q $q --steps 20 --warmup 3 --no-cpu-baseline --no-convergence-leg --no-batch-leg --no-labeled-leg > gpurun_out/bq_$q.json 2> gpurun_out/bq_$q.err || tail -3 gpurun_out/bq_$q.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bq_$q.json').read().strip().splitlines()[-1])
r=d['roofline']
print(f"q=$q  EM iteration {d['ms_per_step']:.3f} ms  {d['value']/1e9:.2f} G upd/s | sweep {r['ms_per_launch']*1e3:.0f} us  {r['achieved']:.0f} GB/s  frac {r['frac']:.3f}  bytes/update {r['algorithmic_bytes']/d['config']['directed_edges_per_gpu']:.1f}")
PY
done
