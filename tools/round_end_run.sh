# The round-end measurement pass (run on the GPU box through gpurun): GPU test suite, default bench, reference arm,
# launch list of the profiled command and one ncu --set full capture of the draws / gradient-pointer step.
# Outputs land in gpurun_out/; the summaries judged are copied to profiles/.
mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "batch/s" | tail -4
( time python bench.py ) > gpurun_out/bench_final6.json 2> gpurun_out/bench_final6.err; tail -4 gpurun_out/bench_final6.err
python bench.py --impl reference > gpurun_out/bench_ref6.json 2> gpurun_out/bench_ref6.err
python tools/show_variants.py gpurun_out/bench_final6.json
PROF="python bench.py --steps 20 --warmup 3 --no-train-step --no-ensemble --no-cpu-baseline --no-e2e --no-variants --no-eager-gpu --no-sample-store"
$PROF > gpurun_out/plain7.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01j_launches.csv $PROF > gpurun_out/ncu_l7.log 2>&1
python tools/run_draws.py > gpurun_out/plain_draws.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'dropout_mix|draw_kernel|step_kernel' -s 8 -c 8 -o gpurun_out/r01j_draws_full python tools/run_draws.py > gpurun_out/ncu_d7.log 2>&1
tail -2 gpurun_out/ncu_d7.log
