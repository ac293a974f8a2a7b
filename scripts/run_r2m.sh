for A in "--config c3" "--config c3 --threads 512" "--config c3 --threads 128" "--config c5" "--config c5 --speculation 0" "--config c5 --speculation 0 --threads 64" "--config c5 --speculation 4"; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu $A 2>/dev/null | grep '^{' | tail -1 > /tmp/b.json; python -c "
import json; d=json.load(open('/tmp/b.json')); print('[$A]', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['roofline']['kernel'], round(d['roofline']['frac'],4))"
done
