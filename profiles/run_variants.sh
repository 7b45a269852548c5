for v in v1 v2 v3 v4 v5 v6 v7 v8; do
  SMCB_LIB=$PWD/particles_b200/variants/libsmcb_$v.so timeout 120 python bench.py --no-cpu --steps 400 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'ms/step', round(d['ms_per_step'],4), 'move_us', round(d['roofline']['avg_launch_us'],1), 'frac', round(d['roofline']['frac'],3))"
done
