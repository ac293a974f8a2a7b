# bench the default library and every variant build libmcpm_<name>.so present (kernel experiments; see scripts/variants.md)
mkdir -p gpurun_out/var
for L in mc-porousmaterials_b200/libmcpm.so mc-porousmaterials_b200/libmcpm_*.so; do
  [ -f "$L" ] || continue
  name=$(basename $L .so)
  for C in c2 c1; do
    MCPM_LIB=$L timeout 300 python bench.py --config $C --steps 5 --warmup 3 --no-cpu --no-configs --no-literal $VARARGS 2>/dev/null | grep '^{' | tail -1 > gpurun_out/var/${name}_$C.json
    python -c "
import json; d=json.load(open('gpurun_out/var/${name}_$C.json')); print('$name', '$C', '%.4g' % d['value'], d['roofline']['kernel'], round(d['roofline']['frac'],4))"
  done
done
