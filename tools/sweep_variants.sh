#!/bin/bash
# runs the default bench (device-resident part only) once per variant library under mass_b200/csrc/variants/
for so in mass_b200/csrc/variants/*.so; do
  echo "== $(basename $so .so)"
  MASSB200_LIB=$so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-class-ids --no-c4 2>gpurun_out/sweep.err | python tools/stage_line.py
done
