#!/bin/bash
# parity tests, then --set full captures (with local-memory and pipe counters) of the kernels named in $@ as
# window:tag[:workload[:bin_size:window_bp]]   (default workload hg38_10kb)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/pytest.log
LOCAL="smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum,smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_lsu.sum"
for spec in "$@"; do
  IFS=: read win tag wl bs wbp <<< "$spec"
  wl=${wl:-hg38_10kb}
  extra=""
  [ -n "$bs" ] && extra="--bin-size $bs --window-bp $wbp"
  python bench.py --probe-kernel --workload $wl --window $win $extra > gpurun_out/probe_$win.log 2>&1 &&
  ncu --set full --metrics $LOCAL --clock-control none --import-source on -k 'regex:^compare_(r1|win)' -s 2 -c 1 \
      -o gpurun_out/r02_${wl}_${win}_$tag -f python bench.py --probe-kernel --workload $wl --window $win $extra > gpurun_out/ncu_$win.log 2>&1
  echo "ncu $wl $win rc $?"; tail -2 gpurun_out/ncu_$win.log
done
