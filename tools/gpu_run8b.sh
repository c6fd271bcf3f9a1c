#!/bin/bash
# 8-GPU box, short form: multi-GPU parity tests, the whole command (GPUs planned by the work), the strong-scaling line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "multi_gpu or ranged or background_regions" > gpurun_out/r02_pytest_n8.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02_pytest_n8.log
python tools/cli_e2e.py --workload hg38_10kb --runs 3 2>&1 | grep -v "INFO\|^chr\|fanc is not" > gpurun_out/r02_cli_e2e_hg38_n8.txt; grep -A9 "^run" gpurun_out/r02_cli_e2e_hg38_n8.txt | cut -c1-120
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench n8 rc $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n8.json").read().strip().splitlines()[-1])
print("  N=%d value %.4g pairs/s ms/step %.3f  e2e %.4g (%.3f ms)  frac %.3f  %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step", 0), d["roofline"]["frac"], {k:v for k,v in d["run"].items() if k in ("sharding_identical","band_bytes_this_rank","pairs_ok")}))
PY
