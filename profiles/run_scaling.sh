#!/bin/sh
# Multi-GPU evidence on an N-GPU box (run from the repository root): bench lines of the weak-scaling splash scene (bench.py's
# default at N > 1), BASELINE configs[3] (64M-particle tank, strong scaling) and configs[4] (16M particles per GPU, splash)
# at every N in "$@" that the box has.  Output: gpurun_out/<tag>_<config>_n<N>.json (+ .err).
# CONFIGS selects the series (default: all three); TESTS=<pytest -k expression> also runs the real-rank parity tests first.
TAG=${TAG:-r02}
CONFIGS=${CONFIGS:-weak strong weak16}
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then
  timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -v -s --timeout 200 -k "$TESTS" > gpurun_out/${TAG}_tests_multi_gpu_$1.log 2>&1
  grep -E "MGPU_RESULT|passed|failed" gpurun_out/${TAG}_tests_multi_gpu_$1.log | cut -c1-200
fi
for n in "$@"; do
  for cfg in $CONFIGS; do
    out=gpurun_out/${TAG}_${cfg}_n${n}
    if [ "$n" = "1" ]; then
      [ "$cfg" = "weak" ] && continue           # N = 1 of the weak series is the headline bench line itself
      timeout 300 python bench.py --config $cfg --no-cpu-baseline > $out.json 2> $out.err
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
          bench.py --gpus $n --config $cfg > $out.json 2> $out.err
    fi
    echo "$cfg n=$n: $(cut -c190-250 $out.json | head -1)"
  done
done
