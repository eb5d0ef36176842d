# dist-only timing of the configs[2] shape for a given library (FPMASH_B200_LIB), no tests
timeout 300 python bench.py --genomes 20 --steps 3 --no-cpu 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); dd=d['dist']; m=dd['merge_all_pairs']
print('dist value %.3f Gpairs/s  step %.1f ms  kernels %s  e2e %.3f  e2e_filtered %.3f Gpairs/s (%s)' % (dd['value']/1e9, dd['ms_per_step'], dd['kernel_ms'], dd['e2e']['value']/1e9, dd['e2e_filtered']['value']/1e9, dd['e2e_filtered'].get('kernel_ms')))
print('merge-all-pairs %.3f Gpairs/s  step %.1f ms  smem-frac %.3f' % (m['value']/1e9, m['ms_per_step'], m['roofline_smem']['frac']))"
