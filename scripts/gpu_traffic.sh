#!/bin/bash
# ncu --set full capture of one 256-channel 3x3 k_conv_tma launch of the detector step -> profiles/traffic.json entry
OUT=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tma -s 40 -c 40 -o $OUT/conv_tma_r02as -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu -i $OUT/conv_tma_r02as.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin)); h=rows[0]
ki=h.index('Kernel Name'); ti=h.index('gpu__time_duration.sum'); ri=h.index('dram__bytes_read.sum'); wi=h.index('dram__bytes_write.sum')
units=rows[1]
print('units', units[ti], units[ri], units[wi])
for r in rows[2:]:
    print(r[ki][:60], r[ti], r[ri], r[wi])
"
