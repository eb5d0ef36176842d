# quick dist-only iteration: correctness (dist tests) then the configs[2] shape; prints tile-kernel and step times
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed" | head -12
timeout 300 python bench.py --genomes 20 --steps 3 --no-cpu 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); dd=d['dist']
print('dist value %.3f Gpairs/s  step %.1f ms  e2e %.3f Gpairs/s  launches %d  kernels %s  steps/pair %.1f' % (dd['value']/1e9, dd['ms_per_step'], dd['e2e']['value']/1e9, dd['gpu_launches'], dd['kernel_ms'], dd['merge_steps_per_pair']))"
