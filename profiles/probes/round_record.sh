set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2m_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2m_smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2m_bench_reference.json 2> gpurun_out/r2m_bench_reference.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2m_ncu_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ekf_window_split_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/prof_window_r2 -f python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-solves > gpurun_out/r2m_ncu_full.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:lane2 --launch-skip 2 --launch-count 1 -o gpurun_out/prof_lane2_r2 -f python profiles/probes/lane_prof_run.py 1024 > gpurun_out/r2m_ncu_lane2.log 2>&1
