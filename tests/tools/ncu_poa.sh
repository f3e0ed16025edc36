#!/bin/bash
# Development aid (GPU box): one ncu --set full capture of the POA window kernel on a quick_poa workload.
# usage: ncu_poa.sh <name> <kernel-regex> <quick_poa args...>   -> gpurun_out/ncu/<name>.ncu-rep + <name>_raw.csv
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/ncu
NAME=$1; shift
KREGEX=$1; shift
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:$KREGEX" -c 1 -f -o gpurun_out/ncu/$NAME \
    python tests/tools/quick_poa.py "$@" > gpurun_out/ncu/${NAME}_app.log 2>&1
ncu -i gpurun_out/ncu/$NAME.ncu-rep --page raw --csv > gpurun_out/ncu/${NAME}_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu/$NAME.ncu-rep --page details --csv > gpurun_out/ncu/${NAME}_details.csv 2>/dev/null
ls -la gpurun_out/ncu/
