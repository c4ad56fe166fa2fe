#!/bin/bash
# Full ncu captures (one launch each) of the convolution kernel variants of one guided score evaluation.
#   tools/ncu_capture.sh <tag>      -> gpurun_out/<tag>_<name>.ncu-rep
# name | PLANES, NB, LN template arguments | launches of that instance to skip
tag=${1:-cap}
while IFS='|' read -r name targs skip; do
  rx="patch_kernel<\\(int\\)${targs// /, \\(int\\)}"
  ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$rx" \
      --launch-skip "$skip" -c 1 -f -o "gpurun_out/${tag}_${name}" python tools/one_eval.py > "gpurun_out/${tag}_ncu_${name}.log" 2>&1
  tail -1 "gpurun_out/${tag}_ncu_${name}.log"
done <<'SPECS'
ln1_c96|2 1 1|1
ln2_c96|2 1 2|1
plain_c96|2 1 0|0
ln1_c384|2 2 1|1
ln2_c384|2 2 2|1
SPECS
ls -la gpurun_out/*.ncu-rep
