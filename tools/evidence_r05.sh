#!/bin/bash
# Round-end evidence run (one B200): GPU tests, smoke, the three bench lines, the ncu launch list of the bench command and
# ncu --set full captures of the dominant conv kernels and the relation-layer kernels.  Outputs under gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/r05_gpu_tests.log 2>&1; echo "pytest rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r05_smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r05_bench_n1.json 2> $O/r05_bench_n1.err; echo "bench rc=$?"
timeout 200 python bench.py --config celebahq --steps 5 --warmup 3 > $O/r05_bench_celebahq.json 2> $O/r05_bench_celebahq.err; echo "celebahq rc=$?"
timeout 120 python tools/bench_relation.py 16 > $O/r05_relation_microbench.txt 2>&1; echo "relation rc=$?"
ADM_NCU_RANGE=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file $O/r05_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r05_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tc_conv_halo|tc_wgrad_rows" -s 3 -c 3 \
  -o $O/r05_conv_full -f python tools/run_conv_once.py > $O/r05_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"rel_gn|lerp_axis" -s 6 -c 6 \
  -o $O/r05_rel_full -f python tools/run_relation_once.py 16 > $O/r05_ncu_rel.log 2>&1; echo "ncu rel rc=$?"
tail -3 $O/r05_gpu_tests.log; tail -1 $O/r05_smoke.log; cut -c1-300 $O/r05_bench_n1.json; cat $O/r05_relation_microbench.txt
