set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s50_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/s50_tests.log
timeout 300 python tools/instinfo_bench.py 2048 > gpurun_out/s50_instinfo_bench.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/s50_instinfo_bench.log
timeout 300 python tools/wsi_bench.py 4608 30 4096 > gpurun_out/s50_wsi.log 2>&1; echo "wsi rc=$?"; tail -2 gpurun_out/s50_wsi.log
