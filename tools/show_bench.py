import json, sys
for line in open(sys.argv[1]):
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line)
        c = d['config']
        print(c['workload'][:30], 'parts', c.get('pipeline_parts'), 'value', round(d['value']), 'ms', round(d['ms_per_step'], 2),
              'e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
        print('  kernels', {k: round(v['avg_launch_us'], 1) for k, v in d['kernels'].items()}, 'conv TF exec',
              round(d['roofline'].get('executed_tflops', 0)), 'frac', round(d['roofline']['frac'], 3))
    elif 'passed' in line or 'failed' in line or 'gpurun]' in line or 'rror' in line:
        print(line[:300])
