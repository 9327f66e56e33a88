# usage: bash tools/run_ncu_short.sh TAG — the captures of tools/run_ncu.sh that change with the read / partition kernels:
# plain run first (must exit 0), launch list, one --set full capture of k_reads_sk per mode and of the two partition kernels
set -x
TAG=${1:-r02x}
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-file-query"
$B > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_reads_sk -s 3 -c 1 -f -o gpurun_out/${TAG}_sk_count $B --count-only > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_reads_sk -s 3 -c 1 -f -o gpurun_out/${TAG}_sk_ids $B --ids-only > gpurun_out/${TAG}_ncu3.log 2>&1
export RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 MASTER_ADDR=127.0.0.1 MASTER_PORT=29533 BLIGHT_CHECK_GENOME=100000000 BLIGHT_CHECK_READS=4000000 BLIGHT_CHECK_M=9 BLIGHT_CHECK_N=10 BLIGHT_CHECK_PLAIN=0 BLIGHT_CHECK_REPS=1 BLIGHT_CHECK_SUB=67108864
python tools/multigpu_check.py > gpurun_out/${TAG}_part_plain.json 2> gpurun_out/${TAG}_part_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_runs_lookup -s 1 -c 1 -f -o gpurun_out/${TAG}_part_lookup python tools/multigpu_check.py > gpurun_out/${TAG}_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dispatch_runs -s 1 -c 1 -f -o gpurun_out/${TAG}_part_dispatch python tools/multigpu_check.py > gpurun_out/${TAG}_ncu5.log 2>&1
ls -la gpurun_out | grep ${TAG}
