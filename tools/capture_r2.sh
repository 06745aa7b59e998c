#!/bin/bash
# Round-2 evidence of the final build, one gpurun call: every capture only after the same command ran clean without ncu.
# Outputs in gpurun_out/fin_*; copied / summarised into profiles/ by hand (profiles/README.md).
set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/fin_tests.log 2>&1; tail -2 $O/fin_tests.log
python bench.py > $O/fin_bench.json 2> $O/fin_bench.err; tail -c 400 $O/fin_bench.json
python bench.py --impl reference > $O/fin_ref.json 2> $O/fin_ref.err; tail -c 300 $O/fin_ref.json
A="--ctx 1 --steps 4 --no-cpu-baseline --no-sharded --no-os1"
python bench.py $A > $O/fin_ctx1.json 2> $O/fin_ctx1.err
S2M_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/fin_launches_mature.csv python bench.py $A > $O/fin_ncu1.log 2>&1
S2M_NCU_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"qgroup_kernel|knn_group_kernel|fit_kernel|solve_kernel" -s 7 -c 4 -o $O/fin_k4 python bench.py --ctx 1 --steps 3 --no-cpu-baseline --no-sharded --no-os1 > $O/fin_ncu2.log 2>&1
python bench_fx.py > $O/fin_fx.json 2> $O/fin_fx.err
python bench_frontend.py > $O/fin_frontend64.json 2> $O/fin_frontend64.err
python bench_frontend.py --batch 1 --host > $O/fin_frontend1.json 2> $O/fin_frontend1.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/fin_frontend_launches.csv python bench_frontend.py --frames 6 --warmup 4 > $O/fin_ncu3.log 2>&1
python tools/single_stream.py --sensor HDL64 > $O/fin_single_hdl.json 2> $O/fin_single_hdl.err
python tools/single_stream.py --sensor VLP16 > $O/fin_single_vlp.json 2> $O/fin_single_vlp.err
python bench_os1.py > $O/fin_os1.json 2> $O/fin_os1.err
S2M_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/fin_launches_os1.csv python bench_os1.py > $O/fin_ncu4.log 2>&1
timeout 600 python tools/stress_parity.py --seeds 30 --frames 10 > $O/fin_stress.txt 2>&1; tail -3 $O/fin_stress.txt
ls -la $O/fin_*
