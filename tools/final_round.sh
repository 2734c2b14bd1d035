#!/bin/bash
# Everything the round's profiles/ directory is built from, on one GPU box:  bash tools/final_round.sh <tag>
set -u
T=${1:-rX}
O=gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3 > $O/${T}_pytest.log
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
python bench.py --impl reference > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
python bench.py --graph off --no-cpu-baseline > $O/${T}_bench_eager.json 2> /dev/null
for c in c2 c3 c4; do python bench.py --config $c > $O/${T}_$c.json 2> $O/${T}_$c.err; done
python tools/bench_train_step.py > $O/${T}_train_step.json 2> $O/${T}_train_step.err
python tools/bench_regressor.py > $O/${T}_regressor.json 2> /dev/null
python tools/bench_render.py > $O/${T}_render.json 2> /dev/null
bash tools/profile_round.sh $T
