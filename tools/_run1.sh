timeout 900 python -m pytest tests/test_bn_gpu.py tests/test_hei_gpu.py tests/test_aread_gpu.py tests/test_tower_gpu.py tests/test_fullsize_gpu.py tests/test_graph_gpu.py -q --timeout 600 > gpurun_out/pytest_bn1.log 2>&1; tail -4 gpurun_out/pytest_bn1.log
timeout 300 python bench.py --no-extra --no-cpu-baseline --steps 50 --warmup 5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('ms_per_step', j['ms_per_step'], 'value', j['value'], 'e2e', j['e2e']['value'])
"
HEI_SHAPES='3,64,64' HEI_ITERS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:hei_tc_fwd -c 8 -o gpurun_out/prof_r2_hei_tc_fwd python tools/bench_hei.py > gpurun_out/ncu_hei_f.log 2>&1
HEI_SHAPES='3,64,64' HEI_ITERS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:hei_tc_bwd -c 8 -o gpurun_out/prof_r2_hei_tc_bwd python tools/bench_hei.py > gpurun_out/ncu_hei_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
