set -x
cd $GRAFT_REPO_ROOT
export PATH=$PATH:/usr/local/cuda/bin
(FBUS_LANE=1 FBUS_LANE_FPC=32 timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_synth_batch.py -m gpu -q -k "lane32 and (trace_rows or missing_detections)" 2>&1 | tail -8) > gpurun_out/r2q_race_lane32.log
(FBUS_LANE=1 timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_synth_batch.py -m gpu -q -k "lane-" 2>&1 | tail -8) > gpurun_out/r2q_race_lane.log
(timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_synth_batch.py tests/test_gpu_sensor_f32.py -m gpu -q -k "lane32 or fused_windows" 2>&1 | tail -8) > gpurun_out/r2q_mem_lane.log
(timeout 600 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_synth_batch.py -m gpu -q -k "lane32 and (trace_rows or missing_detections)" 2>&1 | tail -8) > gpurun_out/r2q_sync_lane.log
