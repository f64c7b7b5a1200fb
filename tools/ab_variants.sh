#!/bin/bash
# A/B of build variants on the 8-view step: bash tools/ab_variants.sh v1 v2 ...   ("-" = the shipped library)
for v in "$@"; do
  if [ "$v" = "-" ]; then r=$(python tools/stage_times.py --views 8); else r=$(GHR_TOOL_VARIANT=$v python tools/stage_times.py --views 8); fi
  echo "$v $(echo "$r" | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fwd', d['blend_forward'], 'bwd', d['blend_backward'], 'eager', round(d['eager_ms'],4), 'graph', round(d['graph_ms'],4), 'ov2', round(d.get('graph_ms_overlap2',0),4))")"
done
