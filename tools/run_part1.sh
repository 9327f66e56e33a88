# single-GPU checks of the fused partition kernels + a world=1 torchrun of the multi-GPU driver
TAG=${1:-p1}
python -m pytest tests/test_gpu_partition.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -30 gpurun_out/${TAG}_pytest.log
BLIGHT_CHECK_GENOME=100000000 BLIGHT_CHECK_READS=4000000 BLIGHT_CHECK_M=7 BLIGHT_CHECK_N=5 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > gpurun_out/${TAG}_w1.json 2> gpurun_out/${TAG}_w1.err
tail -3 gpurun_out/${TAG}_w1.err; cat gpurun_out/${TAG}_w1.json
