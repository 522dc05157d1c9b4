mkdir -p gpurun_out
SPHB_FUZZ_IMAGES=40 SPHB_FUZZ_GRIDS=12 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t.txt 2>&1; echo tests rc=$?; tail -3 gpurun_out/t.txt
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/b_proj64.json 2>gpurun_out/b_err.txt; tail -1 gpurun_out/b_proj64.json | cut -c1-400
python bench.py --steps 5 --warmup 3 --no-cpu --dtype f32 > gpurun_out/b_proj32.json 2>>gpurun_out/b_err.txt; tail -1 gpurun_out/b_proj32.json | cut -c1-300
python bench.py --steps 5 --warmup 3 --no-cpu --workload grid > gpurun_out/b_grid64.json 2>>gpurun_out/b_err.txt; tail -1 gpurun_out/b_grid64.json | cut -c1-300
