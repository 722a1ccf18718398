for i in 1 2 3; do
  echo "coop"; python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-class-ids --no-c4 2>/dev/null | python tools/stage_line.py
  echo "nocoop"; MASSB200_NO_COOP=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-class-ids --no-c4 2>/dev/null | python tools/stage_line.py
done
