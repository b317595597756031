#!/bin/bash
# A/B timing of build variants on the GPU box:  bash profiles/ab.sh "<flags A>" "<flags B>" ...   ("-" = default build)
# Each variant: rebuild the library with NVCC_EXTRA, run the RoIAlign parity tests, then profiles/roi_bench.py.
mkdir -p gpurun_out
for v in "$@"; do
  f="$v"; [ "$v" = "-" ] && f=""
  tag=$(echo "$v" | tr -c 'A-Za-z0-9=\n' '_')
  NVCC_EXTRA="$f" python mxdetection_b200/build.py --force > /dev/null || { echo "build failed: $v"; continue; }
  echo "== variant [$v]"
  NVCC_EXTRA="$f" timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "roi or fpn or mask" 2>&1 | tail -2
  NVCC_EXTRA="$f" timeout 300 python profiles/roi_bench.py 30 | tee gpurun_out/ab_$tag.json
done
