#!/bin/bash
# A/B runs of environment-switched kernel variants on the -r 1 sweep cells: each argument is "ENV=val[,ENV=val]"
mkdir -p gpurun_out
for spec in "$@"; do
  envs=$(echo "$spec" | tr ',' ' ')
  [ "$spec" = "default" ] && envs=""
  echo "== $spec"
  env $envs python bench.py --steps 3 --warmup 3 --workload sweep --sweep-windows ${AB_WINDOWS:-r1} > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -5 gpurun_out/ab.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab.json").read().strip().splitlines()[-1])
for c in d["cells"]: print("bins %3d %-3s %9.4g pairs/s  kern %8.1f us  hbm %.3f" % (c["bins"],c["window"],c["pairs_per_s"],c["kernel_us"],c["roofline_frac"]))
PY
  env $envs python bench.py --steps 50 --warmup 5 --workload chr2_25kb --no-cpu-baseline --no-traffic --window r1 2>gpurun_out/chr2.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('chr2 r1: %.4g pairs/s  frac %.3f  e2e %.4g  %s' % (d['value'], d['roofline']['frac'], d['e2e']['value'], d['roofline']['kernel']))"
done
