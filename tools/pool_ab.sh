for cfg in "SVIT_POOL_PERSIST=-1" "SVIT_POOL_PERSIST=0" "SVIT_POOL_PERSIST=1"; do
  echo "== $cfg"; env $cfg timeout 120 python tools/pool_bench.py 2>&1 | grep packed | sed 's/| contiguous.*//'
done
