#!/bin/bash
# tools/ab_env.sh NAME=VALUE ... : one bench run per setting of an environment variable (default library)
for kv in "$@"; do
  env "$kv" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/abenv.json 2> gpurun_out/abenv.log
  python - "$kv" <<'P'
import json,sys
try:
    d=json.loads(open("gpurun_out/abenv.json").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:34s} value {d['value']/1e6:8.1f} M reads/s  {d['ms_per_step']:7.3f} ms/step  probes/s {d['probes_per_s']/1e9:6.2f} G  e2e {d['e2e']['value']/1e6:7.1f} M  build {d['build']['seconds']:.2f}s")
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open("gpurun_out/abenv.log").read()[-1500:])
P
done
