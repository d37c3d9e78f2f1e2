#!/bin/bash
# A/B runs of tuning variants of the library on the GPU box: tools/ab.sh <variant> [<variant> ...]   ("default" = libslacken_gpu.so)
for v in "$@"; do
  so=libslacken_gpu.so; [ "$v" != default ] && so=libslacken_gpu_$v.so
  SLK_SO=$so timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.log
  python - "$v" <<'P'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/ab_{v}.json").read().strip().splitlines()[-1])
    print(f"{v:12s} value {d['value']/1e6:8.1f} M reads/s  {d['ms_per_step']:7.3f} ms/step  probes/s {d['probes_per_s']/1e9:6.2f} G  ascii {d['value_ascii_input']['value']/1e6:7.1f} M  e2e {d['e2e']['value']/1e6:7.1f} M  e2e_ascii {d['e2e_ascii_input']['value']/1e6:7.1f} M  build {d['build']['seconds']:.2f}s")
except Exception as e:
    print(v, "FAILED", e); print(open(f"gpurun_out/ab_{v}.log").read()[-2000:])
P
done
