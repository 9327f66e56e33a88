# usage: bash tools/run_ncu.sh TAG   (on the GPU box, via gpurun) — every command first WITHOUT ncu (it must exit 0), then the
# launch list (--metrics gpu__time_duration.sum) and one --set full capture per kernel of interest; everything lands in gpurun_out/
set -x
TAG=${1:-r02x}
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-file-query"
$B > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_reads_sk -s 3 -c 1 -f -o gpurun_out/${TAG}_sk_count $B --count-only > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_reads_sk -s 3 -c 1 -f -o gpurun_out/${TAG}_sk_ids $B --ids-only > gpurun_out/${TAG}_ncu3.log 2>&1
# the upload pass (valid / pos_id / filter / exact positions)
ncu --set full --clock-control none --import-source on -k regex:k_window_answers -c 1 -f -o gpurun_out/${TAG}_window_answers $B --count-only > gpurun_out/${TAG}_ncu3b.log 2>&1
# the same read kernel on an index that fits the L2 (locality bound, profiles/r02_locality_bound.json)
$B --genome 2000000 --count-only > gpurun_out/${TAG}_small_plain.json 2> gpurun_out/${TAG}_small_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_reads_sk -s 3 -c 1 -f -o gpurun_out/${TAG}_sk_count_l2 $B --genome 2000000 --count-only > gpurun_out/${TAG}_ncu3c.log 2>&1
# the two kernels of the fused partition path, one rank looping back to itself (the kernels are the ones the NVLink path runs)
export RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 MASTER_ADDR=127.0.0.1 MASTER_PORT=29533 BLIGHT_CHECK_GENOME=100000000 BLIGHT_CHECK_READS=4000000 BLIGHT_CHECK_M=9 BLIGHT_CHECK_N=10 BLIGHT_CHECK_PLAIN=0 BLIGHT_CHECK_REPS=1 BLIGHT_CHECK_SUB=67108864
python tools/multigpu_check.py > gpurun_out/${TAG}_part_plain.json 2> gpurun_out/${TAG}_part_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_runs_lookup -s 1 -c 1 -f -o gpurun_out/${TAG}_part_lookup python tools/multigpu_check.py > gpurun_out/${TAG}_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dispatch_runs -s 1 -c 1 -f -o gpurun_out/${TAG}_part_dispatch python tools/multigpu_check.py > gpurun_out/${TAG}_ncu5.log 2>&1
# GPU index construction: launch list of one 100 M-k-mer build
BUILD_GENOMES=100000000 python tools/build_bench.py > gpurun_out/${TAG}_build_plain.jsonl 2> gpurun_out/${TAG}_build_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_build_launches.csv python tools/build_bench.py > gpurun_out/${TAG}_ncu6.log 2>&1
ls -la gpurun_out | tail -30
