# GPU-box check used during development: parity suite, then the headline benches.
mkdir -p gpurun_out
SPHB_FUZZ_IMAGES=40 SPHB_FUZZ_GRIDS=12 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t.txt 2>&1; echo tests rc=$?; tail -3 gpurun_out/t.txt
for c in ${LUT_COPIES_SWEEP:-16}; do
  for dt in f64 f32; do
    SPHB_LUT_COPIES=$c python bench.py --steps 5 --warmup 3 --no-cpu --dtype $dt > gpurun_out/b_proj_${dt}_c$c.json 2>>gpurun_out/b_err.txt
    echo "copies=$c $dt: $(python -c "import json,sys; d=json.loads(open('gpurun_out/b_proj_${dt}_c$c.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('phases'))")"
  done
done
python bench.py --steps 5 --warmup 3 --no-cpu --workload grid > gpurun_out/b_grid64.json 2>>gpurun_out/b_err.txt; tail -1 gpurun_out/b_grid64.json | cut -c1-200
