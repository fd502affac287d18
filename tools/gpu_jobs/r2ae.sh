# round-2 GPU job ae: final build -- full GPU suite, smoke(), the driver's two commands at N = 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r2ae_pytest.log 2>&1; tail -3 gpurun_out/r2ae_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ae_reference_n1.json 2> gpurun_out/r2ae_reference_n1.err; tail -c 600 gpurun_out/r2ae_reference_n1.json; echo
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ae_n1.json 2> gpurun_out/r2ae_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2ae_n1.json"))
print(f"value {d['value']/1e6:.1f} e2e {d['e2e']['value']/1e6:.1f} Mbase/s hbm {d['config']['hbm_used_gb']} GB parity {d.get('parity_checked')} cpu {d['cpu_baseline']['value']/1e6:.1f} launches {d['gpu_launches']}")
print(d['roofline']['frac'], d['roofline']['block_cells_per_s']/1e9, d['clocks'])
PY
