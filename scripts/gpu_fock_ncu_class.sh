#!/bin/bash
# light ncu (no source) over every Fock-mode launch of one class inside a full-size build
# Usage: CLASSES="1010" bash scripts/gpu_fock_ncu_class.sh [tag] [nx,ny,nz] [count]
TAG=${1:-fc}
W=${2:-4,4,4}
N=${3:-40}
O=gpurun_out/$TAG
mkdir -p $O
for c in ${CLASSES:-1010}; do
  re=$(echo $c | sed -E 's/(.)(.)(.)(.)/\\(int\\)\1, \\(int\\)\2, \\(int\\)\3, \\(int\\)\4/')
  timeout 1200 ncu --section LaunchStats --section Occupancy --section SpeedOfLight --section WarpStateStats \
    --section SchedulerStats --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section InstructionStats \
    --clock-control none --kernel-name-base demangled -k "regex:eri_rowreg.*kernel<$re" -c $N -f -o /tmp/fc_$c \
    python scripts/fock_once.py def2-tzvp $W > $O/fc_$c.log 2>&1
  python scripts/ncu_summary.py /tmp/fc_$c.ncu-rep > $O/fc_${c}_summary.txt 2>&1
done
ls -la $O
