# A/B of draw-kernel builds on the GPU box (alternating, one box): float4 groups per thread x CTA size x register cap
for cfg in "2 128 8" "4 128 4" "4 64 8" "2 128 8" "4 128 4" "3 128 5"; do
  set -- $cfg
  BDL_NVCC_EXTRA="-DBDL_DRAW_U=$1 -DBDL_DRAW_THREADS=$2 -DBDL_DRAW_MINBLOCKS=$3" python -m bayesdll_b200.build bdl_draw.cu > /dev/null
  echo "--- groups/thread $1 threads $2 minblocks $3"; python tools/ab_draw.py
done
python -m bayesdll_b200.build bdl_draw.cu > /dev/null   # back to the default build
