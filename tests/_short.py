import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], 'value %.4g'%d['value'], 'ms/step %.2f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'peak %.1f'%d['roofline']['peak'], d['clocks'])
