#!/bin/bash
# ncu --set full capture of one kernel (regex $1, launch-skip $2) inside a UNet forward; report lands in gpurun_out/$3.ncu-rep
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$1" -s ${2:-4} -c 1 -f -o gpurun_out/${3:-k} python tools/time_unet.py 16 1 ${4:-} > gpurun_out/${3:-k}_ncu.log 2>&1
tail -2 gpurun_out/${3:-k}_ncu.log
