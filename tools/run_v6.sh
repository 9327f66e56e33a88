set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v6.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/v6_default.json 2> gpurun_out/v6_default.err
BLIGHT_FILTER_BITS=0 $B > gpurun_out/v6_nofilter.json 2>&1
BLIGHT_FILTER_BITS=8 $B > gpurun_out/v6_f8.json 2>&1
BLIGHT_FILTER_BITS=16 $B > gpurun_out/v6_f16.json 2>&1
BLIGHT_FILTER_ANCHORS=1 $B > gpurun_out/v6_fanch.json 2>&1
BLIGHT_POS_ID=0 $B > gpurun_out/v6_nopid.json 2>&1
tail -3 gpurun_out/pytest_gpu_v6.log
