for cfg in "SSD3D_STEM_TZ=1 SSD3D_RESERVE_SMS=0" "SSD3D_STEM_TZ=0 SSD3D_RESERVE_SMS=0" "SSD3D_STEM_TZ=1 SSD3D_RESERVE_SMS=8" "SSD3D_STEM_TZ=1 SSD3D_RESERVE_SMS=16" "SSD3D_STEM_TZ=0 SSD3D_RESERVE_SMS=16" "SSD3D_STEM_TZ=1 SSD3D_FUSE_DWPW_PIPELINE=1" "SSD3D_STEM_TZ=1 SSD3D_TAIL_FROM=4" "SSD3D_STEM_TZ=1 SSD3D_TAIL_FROM=2"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-train --no-extras 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), round(d['ms_per_step']*1000,1))
"
done
