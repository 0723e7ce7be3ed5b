mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "ct_transformer" --timeout 600 > gpurun_out/H_punc.log 2>&1; echo "punc rc=$?"; tail -3 gpurun_out/H_punc.log
timeout 300 python tests/diag_punc.py > gpurun_out/H_punc_diag.log 2>&1; tail -6 gpurun_out/H_punc_diag.log
timeout 600 python tools/gemm_shapes.py --tag ring5 > gpurun_out/H_gemm.log 2>&1; echo "gemm rc=$?"; grep -c name gpurun_out/H_gemm.log
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/H_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/H_pytest.log
timeout 600 python bench.py --config 2 > gpurun_out/H_bench2.json 2> gpurun_out/H_bench2.err; tail -c 300 gpurun_out/H_bench2.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/H_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/H_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fbank_tab -c 1 -o gpurun_out/H_fbank python bench.py --steps 1 --warmup 1 > gpurun_out/H_ncu_fbank.log 2>&1
