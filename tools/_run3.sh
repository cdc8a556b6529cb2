timeout 900 python -m pytest tests/test_rowpass_gpu.py tests/test_aread_gpu.py tests/test_step_graph_gpu.py tests/test_fullsize_gpu.py -q --timeout 600 > gpurun_out/pytest_t5.log 2>&1; tail -3 gpurun_out/pytest_t5.log
timeout 300 python bench.py --no-extra --no-cpu-baseline --steps 50 --warmup 5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms_per_step', j['ms_per_step'], 'value', j['value'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'])
"
