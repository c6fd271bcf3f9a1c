python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline --no-extras --no-traffic --steps 20 --warmup 5 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; tail -c 300 gpurun_out/bench_tmp.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_tmp.json").read().strip().splitlines()[-1])
print("value %.4g pairs/s  ms/step %.4f  frac %.4f  e2e %.4g  %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["roofline"]["kernel"]))
PY
