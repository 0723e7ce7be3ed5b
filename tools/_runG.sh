mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/G_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/G_pytest.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 --launch-skip 4 --launch-count 8 -o gpurun_out/G_gemms python bench.py --steps 1 --warmup 1 > gpurun_out/G_ncu_gemms.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --config 2 > gpurun_out/G_bench2.json 2> gpurun_out/G_bench2.err; tail -c 600 gpurun_out/G_bench2.json
