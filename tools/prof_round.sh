set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t_v4.log
python bench.py > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_v4_ref.json 2> gpurun_out/bench_v4_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_v4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_v4.log 2>&1
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/step_counters_v4.csv python tools/one_step.py > gpurun_out/one_step.log 2>&1
ncu --set full --clock-control none --import-source on -c 14 -o gpurun_out/prof_commit_v4 -f python tools/profile_commit.py 20 14 1 > gpurun_out/ncu_commit_v4.log 2>&1
ls -la gpurun_out | tail -8
cat gpurun_out/t_v4.log
