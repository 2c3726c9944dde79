for s in 2 3 4 5; do
python bench.py --streams $s --no-cpu-baseline --no-train --e2e-repeats 1 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams',$s, round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))"
done
