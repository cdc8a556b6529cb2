timeout 600 python -m pytest tests/test_hei_gpu.py tests/test_abi.py -q --timeout 300 > gpurun_out/pytest_hei2.log 2>&1; tail -5 gpurun_out/pytest_hei2.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2p.log 2> gpurun_out/bench_r2p.err; tail -c 600 gpurun_out/bench_r2p.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/bench_r2p.log') if l.startswith('{')][-1])
print(j['ms_per_step'], j['value'], j['e2e']['value'])
for r in j['extra']['hei_layers']['layers']:
    print(r['towers'], r['k'], r['n'], r['input_is_preactivation'], {k: (round(v['fwd_us'],1), round(v['bwd_us'],1), round(v['fwd_frac_of_hbm'],2), round(v['bwd_frac_of_hbm'],2)) for k,v in r.items() if isinstance(v, dict)})
PY
