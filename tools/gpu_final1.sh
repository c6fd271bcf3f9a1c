#!/bin/bash
# round-2 evidence on one GPU: parity tests, the default bench line (with extras), the reference arm, dlbcl, the
# sweep, the whole command, launch list, and --set full captures of the final kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/r02_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc $?"
python bench.py --steps 20 --warmup 5 --window a7 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_n1_a7.json 2> gpurun_out/r02_bench_n1_a7.err; echo "bench a7 rc $?"
python bench.py --steps 5 --warmup 3 --workload dlbcl_5kb_bg1000 > gpurun_out/r02_bench_dlbcl.json 2>/dev/null; echo "dlbcl rc $?"
python bench.py --steps 3 --warmup 3 --workload dlbcl_5kb_bg1000 --window-bp 3000000 > gpurun_out/r02_bench_dlbcl_601.json 2>/dev/null; echo "dlbcl601 rc $?"
python bench.py --steps 3 --warmup 3 --workload sweep > gpurun_out/r02_bench_sweep.json 2>/dev/null; echo "sweep rc $?"
python tools/cli_e2e.py --workload hg38_10kb --runs 3 --devices 0 2>&1 | grep -v "INFO\|^chr\|fanc is not" > gpurun_out/r02_cli_e2e_hg38_n1.txt; tail -12 gpurun_out/r02_cli_e2e_hg38_n1.txt | cut -c1-200
python tools/cli_e2e.py --workload chr2_25kb --runs 3 --devices 0 --background 2>&1 | grep -v "INFO\|^chr\|fanc is not" > gpurun_out/r02_cli_e2e_chr2_bg_n1.txt; tail -12 gpurun_out/r02_cli_e2e_chr2_bg_n1.txt | cut -c1-200
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-traffic > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_default.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-traffic > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc $?"
