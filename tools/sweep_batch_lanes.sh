python -m pytest tests/test_bayes_gpu.py -x -q -m gpu 2>&1 | tail -2
for cfg in "1 2" "2 1" "2 2" "4 1" "4 2" "3 2"; do set -- $cfg; python bench.py --steps 24 --warmup 6 --no-reference-gpu --no-cpu-baseline --job 0 --batch $1 --lanes $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('batch $1 lanes $2:', round(d['value'],1), 'img/s  e2e', round(d['e2e']['value'],1))"; done
