# The GPU pass run at the end of a round (one B200):  gpurun --timeout 1800 -- 'bash tools/final_gpu_run.sh'
# Writes everything under gpurun_out/; the summaries that are judged are copied into profiles/ by hand.
set -x
python -m pytest tests -q -m gpu > gpurun_out/final_pytest.log 2>&1; tail -2 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log | cut -c1-300
python bench.py > gpurun_out/final_bench.log 2> gpurun_out/final_bench.err; cut -c1-600 gpurun_out/final_bench.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_reference.log 2>&1; tail -c 400 gpurun_out/final_reference.log
python tools/profile_encoder.py c2 2>&1 | sed -n "/device kernels/,\$p" > gpurun_out/final_prof_c2.log
python tools/profile_encoder.py c4 2>&1 | sed -n "/device kernels/,\$p" > gpurun_out/final_prof_c4.log
python tools/bench_gemm.py > gpurun_out/final_gemm_bench.log 2>&1
# launch list of the default line's timed region (only after the same command exited 0 without ncu)
python bench.py --steps 2 --warmup 3 --no-c4 --no-small --no-cpu-baseline --no-layer > gpurun_out/plain_launch_c5.log 2>&1 &&
ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/final_launches_c5.csv python bench.py --steps 2 --warmup 3 --no-c4 --no-small --no-cpu-baseline --no-layer \
    > gpurun_out/ncu_launch_c5.log 2>&1
