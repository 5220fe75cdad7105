for v in u1 u2 u4 u1; do cp conan-fgw_b200/lib/variants/$v.so conan-fgw_b200/lib/libconanmp.so; python bench.py --lean --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); a=d['roofline']['all_timed']
print('$v', round(d['ms_per_step'],4), 'fwd', round(a['cmp_cfconv_dense_fwd']['avg_launch_us'],2), 'bwd', round(a['cmp_cfconv_dense_bwd_weights']['avg_launch_us'],2))"; done
