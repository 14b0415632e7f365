import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('%s: value %.4g pts/s, ms/step %.3f, serial %.3f, e2e %.3g pts/s (%.2f ms), d2h %d, build alone %.3f' % (
    sys.argv[2] if len(sys.argv) > 2 else '', d['value'], d['ms_per_step'], d['serial_ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'],
    d['e2e']['d2h_bytes_per_step'], d['stage_ms']['total']))
