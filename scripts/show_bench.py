import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline", {})
print(f"value={d['value']/1e9:.3f} G upd/s  ms/step={d['ms_per_step']:.4f}  e2e={d['e2e']['value']/1e9:.3f} G upd/s ({d['e2e'].get('seconds',0):.2f}s)  "
      f"sweep={r.get('ms_per_launch',0)*1e3:.1f}us  {r.get('achieved',0):.0f} GB/s  frac={r.get('frac',0):.3f}  launches={d.get('gpu_launches')}  clocks={d.get('clocks')}")
p = d.get("partitioned")
if p and 'error' in p:
    print('partitioned leg failed:', p['error'])
elif p:
    print(f"partitioned: value={p['value']/1e9:.3f} G upd/s  ms/step={p['ms_per_step']:.4f}  N={p['n_vertices']}  K={p['directed_edges']}  "
          f"cut={p['cut_fraction']:.3f} ({p['vertex_numbering']})  gen={p['generate_s']:.1f}s build={p['build_s']:.1f}s resid={p['residual_after']:.3g}")
c = d.get("em_convergence")
if c and "error" in c:
    print("convergence leg failed:", c["error"])
elif c:
    print(f"em_convergence: total {c['seconds_total']:.2f}s :", " ".join(f"q{r['q']}:{r['iterations']}it/{r['seconds']*1e3:.0f}ms{'' if r['converged'] else '(!)'}" for r in c["runs"]))
b = d.get("batched_graphs")
if b and "error" in b:
    print("batch leg failed:", b["error"])
elif b:
    print(f"batched_graphs: {b['graphs']} graphs in {b['seconds']:.2f}s = {b['graphs_per_sec']:.1f} graphs/s  converged {b['converged_fraction']:.2f}  mean it {b['mean_iterations']:.0f}  overlap {b['overlap_by_epsilon']}")
l = d.get("labeled_graphs")
if l and "error" in l:
    print("labeled leg failed:", l["error"])
elif l:
    lg = l.get("large_graph", {})
    print(f"labeled_graphs: {l['graphs']} graphs in {l['seconds']:.2f}s = {l['graphs_per_sec']:.1f} graphs/s (second batch {l.get('graphs_per_sec_second_batch', 0):.1f})  converged {l['converged_fraction']:.2f}  mean it {l['mean_iterations']:.0f}  overlap {l['overlap_by_epsilon']}  large: {lg.get('ms_per_em_iteration', 0):.2f} ms/iter = {lg.get('updates_per_sec', 0)/1e9:.2f} G upd/s")
