import json, sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.3fM  ms/step %.4f  e2e %.3fM (%.3f ms)  launches/step %s  clocks %s" % (l['value']/1e6, l['ms_per_step'], l['e2e']['value']/1e6, l['e2e']['ms_per_step'], l.get('launches_per_step'), l['clocks']))
for k in ('act_select_agent_steps_per_s_bs1','act_select_agent_steps_per_s_bs32','replay_sample','cpu_baseline','roofline'):
    if k in l: print(k, l[k])
for k in l['kernels']: print("  %-32s %8.2f us  share %.3f  %s GB/s" % (k['kernel'], k['us'], k['share'], k['gbs']))
