#!/bin/bash
# Final single-GPU validation of the round: GPU tests, smoke, the three bench lines, relation microbenchmark.
set -u
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/r05_gpu_tests.log 2>&1; echo "pytest rc=$?"
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/r05_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r05_bench_n1.json 2> $O/r05_bench_n1.err; echo "bench rc=$?"
timeout 150 python bench.py --config div2k --steps 5 --warmup 3 > $O/r05_bench_div2k.json 2> $O/r05_bench_div2k.err; echo "div2k rc=$?"
timeout 150 python bench.py --config celebahq --steps 5 --warmup 3 > $O/r05_bench_celebahq.json 2> $O/r05_bench_celebahq.err; echo "celebahq rc=$?"
timeout 60 python tools/bench_relation.py 16 > $O/r05_relation_microbench.txt 2>&1; echo "relation rc=$?"
tail -3 $O/r05_gpu_tests.log; tail -1 $O/r05_smoke.log
for f in n1 div2k celebahq; do cut -c1-200 $O/r05_bench_$f.json; tail -2 $O/r05_bench_$f.err; done
