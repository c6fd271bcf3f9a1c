#!/bin/bash
# 2-GPU box: multi-GPU parity tests, then the strong-scaling line, the window sweep and configs[3] under torchrun
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "multi_gpu or ranged or background_regions" > gpurun_out/r02_pytest_n2.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02_pytest_n2.log
run_bench() {  # n, tag, extra args
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $1 $3 > gpurun_out/r02_bench_$2.json 2> gpurun_out/r02_bench_$2.err
  echo "bench $2 rc $?"; tail -c 300 gpurun_out/r02_bench_$2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_$2.json").read().strip().splitlines()[-1])
    print("  N=%d value %.4g %s ms/step %.3f  e2e %.4g (%.3f ms)  frac %.3f  %s" % (d["n_gpus"], d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step", 0), d["roofline"]["frac"], {k:v for k,v in d.get("run", {}).items() if k in ("sharding_identical","band_bytes_this_rank","pairs_ok","device_call_matches_engine")}))
    for c in d.get("cells", []): print("   bins %3d %-3s %9.4g pairs/s  kern %8.1f us  hbm %.3f" % (c["bins"],c["window"],c["pairs_per_s"],c["kernel_us"],c["roofline_frac"]))
except Exception as e: print("  parse failed", e)
PY
}
run_bench 2 n2 "--steps 20 --warmup 5"
run_bench 2 n2_sweep "--steps 3 --warmup 3 --workload sweep --sweep-windows r1,a7"
run_bench 2 n2_dlbcl "--workload dlbcl_5kb_bg1000 --steps 3 --warmup 3"
