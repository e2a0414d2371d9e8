# one-loop grid thinning experiment: parity of the tables (dense fixtures) and the kernel times for several (low, high) steps
for st in ${OL_STEPS:-12,6}; do  # needs the PRSD_EXPERIMENT_OL_STEPS hook of commit "J00, J01, J11 ... subset" re-applied in plan.cu
export PRSD_EXPERIMENT_OL_STEPS=$st
echo "=== steps $st"
python -m pytest tests/test_gpu_dense.py tests/test_gpu_pt.py tests/test_gpu_galaxy.py -m gpu -q 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_parity.json'))
for name in ('golden','bench0','bench1','benchmax','benchmin'):
    v=d[name]
    print(name, {k:'%.2e'%x[0] for k,x in v.items() if k.startswith('oneloop')}, 'K max %.2e'%max(x[0] for k,x in v.items() if k.startswith('K[')),
          'I max %.2e'%max(x[0] for k,x in v.items() if len(k)==3 and k[0]=='I'), 'ML max %.2e'%max(x[0] for k,x in v.items() if k.startswith('ML')))
PY
python bench.py --steps 6 --warmup 3 --no-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
sh=d['roofline']['kernel_share_of_step']
print('   step', d['ms_per_step'], {k:round(v*d['ms_per_step'],3) for k,v in sh.items()}, d['config'].get('quadrature_nodes_per_cosmology'))
"
done
