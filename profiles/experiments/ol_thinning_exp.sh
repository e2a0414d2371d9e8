# one-loop grid thinning experiment: parity of the one-loop tables (dense fixtures) and the kernel times
python -m pytest tests/test_gpu_dense.py tests/test_gpu_pt.py tests/test_gpu_galaxy.py -m gpu -q 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_parity.json'))
for name in ('golden','bench0','bench1','benchmax','benchmin'):
    print(name, {k:'%.2e'%v[0] for k,v in d[name].items() if k.startswith('oneloop')})
PY
python bench.py --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
sh=d['roofline']['kernel_share_of_step']
print('   step', d['ms_per_step'], {k:round(v*d['ms_per_step'],3) for k,v in sh.items()}, d['config'].get('quadrature_nodes_per_cosmology'))
"
