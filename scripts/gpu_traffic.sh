#!/bin/bash
# ncu --set full capture of ONE 256-channel 3x3 k_conv_tma launch of the default bench step (15 frames x 4 agents) -> the numbers for
# profiles/traffic.json.  k_conv_tma launches per step: 4 level-1, deblock 1, 6 level-2, deblock 2, level-3 stride-2, 8 level-3, ...
# = 29; after 3 warm-up steps the launch with index 3 * 29 + 14 is a stride-1 256-channel layer.
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none -k regex:k_conv_tma -s 101 -c 1 -o $OUT/conv_tma_traffic -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu -i $OUT/conv_tma_traffic.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for k in ('Kernel Name','Grid Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum'):
    i=h.index(k); print(k, '|', rows[1][i], '|', rows[2][i][:120])
"
rm -f $OUT/conv_tma_traffic.ncu-rep
