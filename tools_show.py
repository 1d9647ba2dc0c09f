import json, sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.3fM  ms/step %.4f  wall %.4f  e2e %.3fM (%.3f ms)  launches/step %s  clocks %s" % (l['value']/1e6, l['ms_per_step'], l.get('wall_ms_per_step',0), l['e2e']['value']/1e6, l['e2e']['ms_per_step'], l.get('launches_per_step'), l['clocks']))
for k,v in l.items():
    if k.startswith('act_select') or k in ('cpu_baseline','roofline'): print(k, v)
for k in l.get('hbm_kernels',[]): print("  HBM", k)
for k in l['kernels']:
    try: print("  %-32s x%-4s %8.2f us/launch %8.2f us/step share %.3f  %s GB/s" % (k['kernel'], k['launches_per_step'], k['us_per_launch'], k['us_per_step'], k['share'], k['algo_gbs']))
    except BrokenPipeError: break
