#!/bin/bash
# A/B timing WITHOUT the parity tests (sensitivity probes that compute wrong results on purpose):
#   bash profiles/ab_notest.sh "<flags A>" "<flags B>" ...      ("-" = default build)
mkdir -p gpurun_out
for v in "$@"; do
  f="$v"; [ "$v" = "-" ] && f=""
  tag=$(echo "$v" | tr -c 'A-Za-z0-9=\n' '_')
  NVCC_EXTRA="$f" python mxdetection_b200/build.py --force > /dev/null || { echo "build failed: $v"; continue; }
  echo "== variant [$v]"
  NVCC_EXTRA="$f" timeout 300 python profiles/roi_bench.py 30 | tee gpurun_out/abn_$tag.json
done
