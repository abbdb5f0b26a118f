# A/B of the MC-Dropout mix kernel on one box: float4 groups per thread
for u in 1 2 1 2; do
  BDL_NVCC_EXTRA="-DBDL_MIX_U=$u" python -m bayesdll_b200.build bdl_draw.cu > /dev/null
  echo "--- groups/thread $u"; python tools/ab_draw.py --mix
done
python -m bayesdll_b200.build bdl_draw.cu > /dev/null   # back to the default build
