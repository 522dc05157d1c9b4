# Development helper: bench (and parity-check) experimental library variants built with
#   python -m sarracen_b200.build --suffix _x -DNAME=1
# usage: VARIANTS="_s1 _s2" bash profiles/run_variants.sh
mkdir -p gpurun_out
for v in "" $VARIANTS; do
  lib=/root/repo/sarracen_b200/lib/libsphb200$v.so
  SPHB_LIB_OVERRIDE=$lib timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1
  for dt in f64 f32; do
    SPHB_LIB_OVERRIDE=$lib python bench.py --steps 5 --warmup 3 --no-cpu --dtype $dt ${BENCH_ARGS:-} > gpurun_out/b_var${v}_$dt.json 2>>gpurun_out/b_err.txt
    echo "variant '$v' $dt: $(python -c "import json; d=json.loads(open('gpurun_out/b_var${v}_$dt.json').read().strip().splitlines()[-1]); print(d['ms_per_step'])")"
  done
done
