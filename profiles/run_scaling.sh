#!/bin/sh
# Multi-GPU evidence on an N-GPU box (run from the repository root): bench lines of the weak-scaling splash scene (bench.py's
# default at N > 1), BASELINE configs[3] (64M-particle tank, strong scaling) and configs[4] (16M particles per GPU, splash)
# at every N in "$@" that the box has.  Output: gpurun_out/<tag>_<config>_n<N>.json (+ .err).
TAG=${TAG:-r02}
mkdir -p gpurun_out
for n in "$@"; do
  for cfg in weak strong weak16; do
    out=gpurun_out/${TAG}_${cfg}_n${n}
    if [ "$n" = "1" ]; then
      [ "$cfg" = "weak" ] && continue           # N = 1 of the weak series is the headline bench line itself
      python bench.py --config $cfg --no-cpu-baseline > $out.json 2> $out.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
          bench.py --gpus $n --config $cfg > $out.json 2> $out.err
    fi
    echo "$cfg n=$n: $(cut -c190-250 $out.json | head -1)"
  done
done
