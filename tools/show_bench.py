import json, sys
for line in open(sys.argv[1]):
    line = line.strip()
    if line.startswith('{'):
        d = json.loads(line)
        c = d['config']
        print(c['workload'][:30], 'parts', c.get('pipeline_parts'), 'value', round(d['value']), 'ms', round(d['ms_per_step'], 2),
              'e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
        def flat(kd):
            out = {}
            for k, v in kd.items():
                if 'avg_launch_us' in v:
                    out[k] = round(v['avg_launch_us'], 1)
                elif isinstance(v, dict):
                    out.update({k + '/' + kk: vv for kk, vv in flat(v).items()})
            return out
        print('  kernels', flat(d.get('kernels', {})), 'conv TF exec',
              round(d['roofline'].get('executed_tflops', 0)), 'frac', round(d['roofline']['frac'], 3))
        for name, c in d.get('configs', {}).items():
            if 'value' in c:
                print('  config', name, 'value', round(c['value']), 'e2e', round(c['e2e']['value']), 'steps', c['steps'],
                      'clk', c['clocks'].get('sm_mhz'), c['clocks'].get('samples'), flat(c.get('kernels', {})))
            else:
                print('  config', name, c)
        if 'train_step' in d:
            print('  train_step', {k: v for k, v in d['train_step'].items() if k != 'impl'})
    elif 'passed' in line or 'failed' in line or 'gpurun]' in line or 'rror' in line:
        print(line[:300])
