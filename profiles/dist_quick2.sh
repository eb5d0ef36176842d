# dist-only timing of the configs[2] shape for a given library (FPMASH_B200_LIB), no tests
timeout 300 python bench.py --genomes 20 --steps 3 --no-cpu 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); dd=d['dist']
print('dist value %.3f Gpairs/s  step %.1f ms  smem-frac %.3f  e2e %.3f Gpairs/s' % (dd['value']/1e9, dd['ms_per_step'], dd['roofline_smem']['frac'], dd['e2e']['value']/1e9))"
