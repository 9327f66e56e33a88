# N-GPU run of the multi-GPU driver: bash tools/run_part2.sh TAG NGPU [env assignments...]
TAG=$1; N=$2; shift 2
env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
echo "rc=$?"; tail -5 gpurun_out/${TAG}.err; cat gpurun_out/${TAG}.json
