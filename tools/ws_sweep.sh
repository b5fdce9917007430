for o in 3 4 5 6 7; do
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 --opt ws_scatter_warps=$o 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$o', round(d['value']), round(d['ms_per_step'],4), round(d['stage_ms_per_step']['warp'],4))"
done
