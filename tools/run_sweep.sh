TAG=${1:-r01}
python tools/sweep.py > gpurun_out/${TAG}_sweep.jsonl 2> gpurun_out/${TAG}_sweep.err; echo "sweep rc=$?"; tail -3 gpurun_out/${TAG}_sweep.err; cut -c1-200 gpurun_out/${TAG}_sweep.jsonl
