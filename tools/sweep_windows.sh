#!/bin/bash
# BASELINE configs[4]: SSIM window sweep (-a 3..11 and the default -r 1) x submatrix size 41..401 bins on chr2 @ 25 kb,
# one bench.py run per cell; prints pairs/s, roofline fraction and the compare kernel time.
#   bash tools/sweep_windows.sh > gpurun_out/sweep_windows.txt
for wbp in 1000000 2000000 3000000 5000000 10000000; do
  for win in r1 a3 a5 a7 a9 a11; do
    python bench.py --window-bp $wbp --window $win --steps 20 --warmup 3 --no-cpu-baseline --no-genome-wide \
      --no-ingest --no-background --no-synthetic 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('bins %4d window %-3s: %10.4g pairs/s  roofline %.3f  %s' % (d['config']['bins_per_pair'], d['config']['window'], d['value'], d['roofline']['frac'], d['roofline']['kernel'].split(', ',1)[1]))"
  done
done
