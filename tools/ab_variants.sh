#!/bin/bash
# A/B of kernel variants on one GPU box: tools/ab_variants.sh <out.jsonl> <variant> [<variant> ...]   ("default" = libslacken_gpu.so)
out=$1; shift
for v in "$@"; do
  if [ "$v" = default ]; then so=libslacken_gpu.so; else so=libslacken_gpu_$v.so; fi
  SLK_SO=$so python bench.py --quick --steps 10 --warmup 3 2>/dev/null | tail -1 >> "$out"
done
cat "$out"
