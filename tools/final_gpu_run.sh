set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r1_final.log 2>&1; tail -3 gpurun_out/pytest_r1_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1_final.log 2>&1; tail -1 gpurun_out/smoke_r1_final.log
python bench.py > gpurun_out/bench_r1_final.log 2> gpurun_out/bench_r1_final.err; tail -c 600 gpurun_out/bench_r1_final.err; cut -c1-1200 gpurun_out/bench_r1_final.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.log 2>&1; tail -c 700 gpurun_out/bench_r1_reference.log
ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_c5.csv python bench.py --steps 2 --warmup 3 --no-c4 --no-cpu-baseline --no-layer > gpurun_out/ncu_launch_c5.log 2>&1
ncu --nvtx --nvtx-include "timed/" --set full --clock-control none --import-source on -c 6 -f -o gpurun_out/spmm_r1_c5 python bench.py --steps 2 --warmup 3 --no-c4 --no-cpu-baseline --no-layer > gpurun_out/ncu_full_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"graphnorm|layernorm|colstats_kernel" -s 24 -c 10 -f -o gpurun_out/norm_r1 python tools/time_graphnorm.py > gpurun_out/ncu_norm.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
