#!/bin/bash
# round 2, GPU call 1: parity tests, the default bench line, launch list, and --set full captures of the
# two kernels that carry the north-star numbers (hg38 @ 10 kb, -r 1 and -a 7).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r02_pytest1.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc $?"; tail -c 600 gpurun_out/r02_bench_n1.err
python bench.py --steps 5 --warmup 3 --workload dlbcl_5kb_bg1000 > gpurun_out/r02_bench_dlbcl.json 2> gpurun_out/r02_bench_dlbcl.err; echo "dlbcl rc $?"; tail -c 600 gpurun_out/r02_bench_dlbcl.err
python bench.py --steps 3 --warmup 3 --workload sweep > gpurun_out/r02_bench_sweep.json 2> gpurun_out/r02_bench_sweep.err; echo "sweep rc $?"; tail -c 600 gpurun_out/r02_bench_sweep.err
LOCAL="smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum,smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_lsu.sum"
for win in r1 a7; do
  python bench.py --probe-kernel --window $win > gpurun_out/probe_$win.log 2>&1 &&
  ncu --set full --metrics $LOCAL --clock-control none --import-source on -k 'regex:^compare_(r1|win)' -s 2 -c 1 \
      -o gpurun_out/r02_hg38_$win -f python bench.py --probe-kernel --window $win > gpurun_out/ncu_$win.log 2>&1
  echo "ncu $win rc $?"; tail -3 gpurun_out/ncu_$win.log
done
ls -la gpurun_out | tail -20
