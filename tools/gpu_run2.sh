#!/bin/bash
# kernel iteration: parity tests, then the sweep cells of the windows given in $1 (default r1)
mkdir -p gpurun_out
WINS=${1:-r1}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc $?"; tail -8 gpurun_out/pytest.log
python bench.py --steps 3 --warmup 3 --workload sweep --sweep-windows $WINS > gpurun_out/sweep.json 2> gpurun_out/sweep.err; echo "sweep rc $?"; tail -c 400 gpurun_out/sweep.err
python - <<PY
import json
d=json.loads(open("gpurun_out/sweep.json").read().strip().splitlines()[-1])
for c in d["cells"]: print("bins %3d %-3s %9.4g pairs/s  kern %8.1f us  hbm %.3f  fp64 %.3f" % (c["bins"],c["window"],c["pairs_per_s"],c["kernel_us"],c["roofline_frac"],c["fp64_frac"]))
PY
python bench.py --steps 50 --warmup 5 --workload chr2_25kb --no-cpu-baseline --no-traffic --window r1 2>gpurun_out/chr2.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('chr2 r1: %.4g pairs/s  frac %.3f  %s' % (d['value'], d['roofline']['frac'], d['roofline']['kernel']))"
