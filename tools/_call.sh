mkdir -p gpurun_out
for c in 16 32 64; do
CB_DESC_CHUNK=$c timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        b=json.loads(l);print('chunk $c', b['value'],b['e2e']['value'])"
done
