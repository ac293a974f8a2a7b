for A in "" "--capacity 672" "--capacity 640" "--replicas 256" "--replicas 384 --capacity 672"; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-configs --no-literal $A 2>/dev/null | grep '^{' | tail -1 > /tmp/b.json; python -c "
import json; d=json.load(open('/tmp/b.json')); print('args [$A]', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], round(d['roofline']['frac'],4), d['arm']['capacity'], d['arm']['chains_per_gpu'])"
done
