#!/bin/sh
# Single-GPU evidence on the GPU box (run from the repository root): the GPU parity suite, the bench lines of BASELINE's
# three single-GPU configs, the reference arm, the ncu launch list of the bench command and one `--set full` capture of every
# kernel of the step.  Output lands in gpurun_out/ (merged back by gpurun); profiles/summarize.py turns it into the
# tracked summaries.  Nothing printed by a run under ncu is used as a bench number.
TAG=${1:-r02z}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests_gpu.log 2>&1
tail -n 3 gpurun_out/${TAG}_tests_gpu.log
for c in c1 c2 c3; do
  python bench.py --config $c > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
  cut -c1-200 gpurun_out/${TAG}_bench_$c.json
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cut -c1-200 gpurun_out/${TAG}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
# prof_step.py: 3 warm steps of 17 matching launches each, then the stage calls (13) and one whole step (17)
ncu --set full --clock-control none --import-source on \
    -k regex:"k_lambda|k_delta_p|k_vorticity|k_plan|k_onesweep|k_predict|k_reorder|k_build_runs" \
    --launch-skip 51 --launch-count 30 -f -o gpurun_out/prof_${TAG} python profiles/prof_step.py > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/prof_${TAG}.ncu-rep
tail -n 3 gpurun_out/${TAG}_ncu_full.log
