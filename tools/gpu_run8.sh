#!/bin/bash
# 8-GPU box: the multi-GPU parity test, the whole command on all GPUs, and the strong-scaling bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "multi_gpu or ranged" > gpurun_out/pytest8.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest8.log
python tools/cli_e2e.py --workload hg38_10kb --runs 2 2>&1 | grep -v "INFO\|^chr\|fanc is not" > gpurun_out/r02_cli_e2e_hg38_n8.txt; tail -22 gpurun_out/r02_cli_e2e_hg38_n8.txt | cut -c1-400
for n in 8 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err
  echo "bench n=$n rc $?"; tail -c 300 gpurun_out/r02_bench_n$n.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_n$n.json").read().strip().splitlines()[-1])
    print("N=%d value %.4g pairs/s ms/step %.3f  e2e %.4g (%.3f ms)  frac %.3f  run %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], {k:v for k,v in d["run"].items() if k!="e2e_cold"}))
    print("  cold", d["run"].get("e2e_cold"))
except Exception as e: print("parse failed", e)
PY
done
