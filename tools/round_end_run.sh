# The round-end measurement pass (run on the GPU box through gpurun): GPU test suite, default bench, reference arm,
# launch list of the profiled command and ncu --set full captures of the step builds and the draws.
# Outputs land in gpurun_out/; the summaries judged are copied to profiles/ (and profiles/traffic.json is regenerated
# from the capture with tools/traffic_from_ncu.py, which ties it to the build).
R=${R:-r02}
mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "batch/s" | tail -4
( time python bench.py ) > gpurun_out/${R}_bench_final.json 2> gpurun_out/${R}_bench_final.err; tail -4 gpurun_out/${R}_bench_final.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${R}_bench_ref_final.json 2> gpurun_out/${R}_bench_ref_final.err
python tools/show_variants.py gpurun_out/${R}_bench_final.json
PROF="python bench.py --steps 20 --warmup 3 --no-train-step --no-ensemble --no-cpu-baseline --no-e2e --no-variants --no-eager-gpu --no-sample-store --no-cfg1"
$PROF > gpurun_out/${R}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv $PROF > gpurun_out/${R}_ncu_launches.log 2>&1
# the .ncu-rep files exceed what gpurun copies back (64 MiB): export the raw pages here, keep the reports in /tmp
python tools/run_steps.py --generic > gpurun_out/${R}_plain_steps.log 2>&1 && ncu --set full --clock-control none -k regex:'step_' -s 12 -c 16 -o /tmp/${R}_steps_full python tools/run_steps.py --generic > gpurun_out/${R}_ncu_steps.log 2>&1
tail -2 gpurun_out/${R}_ncu_steps.log
ncu -i /tmp/${R}_steps_full.ncu-rep --page raw --csv > gpurun_out/${R}_steps_full_raw.csv
python -c "from bayesdll_b200 import build; print(build.source_hash())" > gpurun_out/${R}_steps_build_hash.txt
python tools/run_draws.py > gpurun_out/${R}_plain_draws.log 2>&1 && ncu --set full --clock-control none -k regex:'dropout_mix|draw_kernel' -s 6 -c 6 -o /tmp/${R}_draws_full python tools/run_draws.py > gpurun_out/${R}_ncu_draws.log 2>&1
tail -2 gpurun_out/${R}_ncu_draws.log
ncu -i /tmp/${R}_draws_full.ncu-rep --page raw --csv > gpurun_out/${R}_draws_full_raw.csv
python tools/run_probe.py > gpurun_out/${R}_plain_probe.log 2>&1 && ncu --set full --clock-control none -k regex:probe_stream -c 9 -o /tmp/${R}_probe python tools/run_probe.py > gpurun_out/${R}_ncu_probe.log 2>&1
ncu -i /tmp/${R}_probe.ncu-rep --page raw --csv > gpurun_out/${R}_probe_stream_raw.csv
# then, back in the build container:  cp gpurun_out/${R}_steps_full_raw.csv profiles/ && python tools/traffic_from_ncu.py profiles/${R}_steps_full_raw.csv --build-hash $(cat gpurun_out/${R}_steps_build_hash.txt)
