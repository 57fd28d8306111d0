set -x
O=gpurun_out/r02; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $O/final_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -3 | tee $O/final_smoke.txt
timeout 600 python bench.py 2>$O/final_bench.err | tail -1 > $O/final_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>$O/final_ref.err | tail -1 > $O/final_ref.json
python -c "
import json
d=json.load(open('$O/final_bench.json')); r=json.load(open('$O/final_ref.json'))
print('OURS', d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])
print('SEC', {k:(v['ms_per_step'], v['value']) for k,v in d.get('secondary',{}).items()})
print('REF', r['value'], r.get('cpu_baseline',{}).get('kind'), r['config']==d['config'])"
