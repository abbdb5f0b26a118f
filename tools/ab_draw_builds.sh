# A/B of draw-kernel builds on the GPU box: min resident CTAs per SM (register cap) x CTA size
for cfg in "8 128" "12 128" "16 128" "8 256" "4 256"; do
  set -- $cfg
  BDL_NVCC_EXTRA="-DBDL_DRAW_MINBLOCKS=$1 -DBDL_DRAW_THREADS=$2" python -m bayesdll_b200.build bdl_draw.cu > /dev/null
  echo "--- minblocks $1 threads $2"; python tools/ab_draw.py
done
python -m bayesdll_b200.build bdl_draw.cu > /dev/null   # back to the default build
