for d in 8 4 2 1 0; do echo "GHR_DIRECT=$d"; GHR_DIRECT=$d python tools/stage_times.py --views 8 --steps 20 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print(d['V'], 'bwd', d['blend_backward'], 'fwd', d['blend_forward'], 'graph_ms', round(d['graph_ms'],4))
    except Exception as e: print(l[:200])
"; done
for i in 1 3; do echo "GHR_ILPB=$i"; GHR_ILPB=$i python tools/stage_times.py --views 8 --steps 20 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print(d['V'], 'bwd', d['blend_backward'], 'graph_ms', round(d['graph_ms'],4))
    except Exception as e: print(l[:200])
"; done
