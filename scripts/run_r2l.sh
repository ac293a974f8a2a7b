for L in mc-porousmaterials_b200/libmcpm.so mc-porousmaterials_b200/libmcpm_kc2.so; do
 for R in 148 296; do
  MCPM_LIB=$L timeout 300 python bench.py --config c4 --replicas $R --steps 3 --warmup 3 --no-cpu 2>/dev/null | grep '^{' | tail -1 > /tmp/b.json; python -c "
import json; d=json.load(open('/tmp/b.json')); print('$L', $R, '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['roofline']['kernel'], round(d['roofline']['frac'],4))"
 done
done
timeout 600 python -m pytest tests -m gpu -q -k "cell_list or c4_" 2>&1 | tail -3
