O=gpurun_out/r2t; mkdir -p $O; cd /root/repo
timeout 600 python -m pytest tests/test_decoder_gpu.py tests/test_zz_model_route.py tests/test_pipeline_gpu.py -m gpu -q -x > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
timeout 300 python tools/determinism_probe.py 16 > $O/determinism.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 2 > $O/bench_default.json 2> $O/bench_default.err
tail -n 4 $O/tests.log; tail -n 17 $O/determinism.log | cut -c1-200; tail -c 300 $O/bench_default.json
