mkdir -p gpurun_out
timeout 600 python tools/gemm_shapes.py --tag tail > gpurun_out/F_gemm_tail.log 2>&1; echo "gemm tail rc=$?"
FA_GEMM_TAIL=0 timeout 600 python tools/gemm_shapes.py --tag whole > gpurun_out/F_gemm_whole.log 2>&1; echo "gemm whole rc=$?"
tail -3 gpurun_out/F_gemm_tail.log
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/E_pytest.log 2>&1; echo "pytest rc=$?" 
tail -8 gpurun_out/E_pytest.log
FA_RZ_COMP=0 NOISE_PROBE_OUT=noise_rz0.json timeout 600 python tools/noise_probe.py > gpurun_out/E_noise0.log 2>&1
FA_RZ_COMP=1 NOISE_PROBE_OUT=noise_rz1.json timeout 600 python tools/noise_probe.py > gpurun_out/E_noise1.log 2>&1
tail -4 gpurun_out/E_noise0.log; tail -4 gpurun_out/E_noise1.log
timeout 600 python bench.py --config 2 > gpurun_out/E_bench2.json 2> gpurun_out/E_bench2.err; tail -c 1200 gpurun_out/E_bench2.json
FA_RZ_COMP=0 timeout 600 python bench.py --config 2 > gpurun_out/E_bench2_rz0.json 2> gpurun_out/E_bench2_rz0.err
FA_GEMM_TAIL=0 timeout 600 python bench.py --config 2 > gpurun_out/E_bench2_whole.json 2> gpurun_out/E_bench2_whole.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fbank_lfr -c 1 -o gpurun_out/E_fbank python bench.py --steps 1 --warmup 1 > gpurun_out/E_ncu_fbank.log 2>&1
