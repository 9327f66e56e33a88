import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e=d.get('e2e') or {}
print(sys.argv[1],'count %.4e (%.2f ms)  ids %.4e (%.2f ms) e2e %s (%s ms) frac %.3f'%(d['value'], d['ms_per_step'], d['ids_mode']['value'], d['ids_mode']['ms_per_step'], e.get('value'), e.get('ms_per_step'), d['roofline']['frac']))
