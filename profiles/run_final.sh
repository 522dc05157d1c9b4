# Round-end measurement set on one GPU: full parity suite, smoke, default bench line (with CPU
# baseline), reference arm, float32 / 2^27-particle projection, voxel grid and its hmin variant.
mkdir -p gpurun_out
SPHB_FUZZ_IMAGES=60 SPHB_FUZZ_GRIDS=20 timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t.txt 2>&1; echo tests rc=$?; tail -2 gpurun_out/t.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/bench_default.json 2>gpurun_out/bench_default.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err; echo ref rc=$?
python bench.py --no-cpu --dtype f32 > gpurun_out/bench_proj_f32.json 2>>gpurun_out/b_err.txt
python bench.py --no-cpu --particles 134217728 --steps 3 > gpurun_out/bench_proj_128m.json 2>>gpurun_out/b_err.txt
python bench.py --no-cpu --workload grid > gpurun_out/bench_grid.json 2>>gpurun_out/b_err.txt
python bench.py --no-cpu --workload grid --dtype f32 > gpurun_out/bench_grid_f32.json 2>>gpurun_out/b_err.txt
python bench.py --no-cpu --workload grid_hmin > gpurun_out/bench_grid_hmin.json 2>>gpurun_out/b_err.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f.split('/')[-1], 'ms', round(d.get('ms_per_step', 0), 2), 'value %.3g' % d.get('value', 0), 'e2e ms', round((d.get('e2e') or {}).get('ms_per_step', 0), 1),
          'cpu', (d.get('cpu_baseline') or {}).get('value'), 'clk', (d.get('clocks') or {}).get('sm_mhz'), (d.get('clocks') or {}).get('reasons'))
PY
# ncu evidence for the same command (after it exited 0 above): launch list and one full capture
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain.json 2>gpurun_out/plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_proj.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_list.log 2>&1; echo ncu list rc=$?
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/proj_v9 -f python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo ncu full rc=$?
