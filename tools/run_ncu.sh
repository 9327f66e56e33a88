# usage: bash tools/run_ncu.sh TAG   (on the GPU box, via gpurun) — bench without ncu first, then the launch list and one
# --set full capture per read kernel mode; everything lands in gpurun_out/
set -x
TAG=${1:-r01x}
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_reads_sk -s 3 -c 1 -f -o gpurun_out/${TAG}_sk_count $B --count-only > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_reads_sk -s 3 -c 1 -f -o gpurun_out/${TAG}_sk_ids $B --ids-only > gpurun_out/${TAG}_ncu3.log 2>&1
ls -la gpurun_out
