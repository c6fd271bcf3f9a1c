#!/bin/bash
# 8-GPU box: the multi-GPU parity tests, the whole command on all GPUs, and the strong-scaling bench lines.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "multi_gpu or ranged or background_regions" > gpurun_out/r02_pytest_n8.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02_pytest_n8.log
python tools/cli_e2e.py --workload hg38_10kb --runs 3 2>&1 | grep -v "INFO\|^chr\|fanc is not" > gpurun_out/r02_cli_e2e_hg38_n8.txt; grep -A9 "^run" gpurun_out/r02_cli_e2e_hg38_n8.txt | cut -c1-120
run_bench() {  # n, tag, extra args
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $1 --steps 20 --warmup 5 $3 > gpurun_out/r02_bench_$2.json 2> gpurun_out/r02_bench_$2.err
  echo "bench $2 rc $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_$2.json").read().strip().splitlines()[-1])
    print("  N=%d value %.4g pairs/s ms/step %.3f  e2e %.4g (%.3f ms)  frac %.3f  %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step", 0), d["roofline"]["frac"], {k:v for k,v in d["run"].items() if k in ("sharding_identical","band_bytes_this_rank","pairs_ok","device_call_matches_engine")}))
except Exception as e: print("  parse failed", e)
PY
}
run_bench 8 n8 ""
run_bench 4 n4 ""
run_bench 2 n2 ""
run_bench 8 n8_a7 "--window a7"
run_bench 8 n8_dlbcl "--workload dlbcl_5kb_bg1000 --steps 5"
