O=gpurun_out/r2x; mkdir -p $O; cd /root/repo
timeout 200 python -m pytest tests/test_feat_gpu.py tests/test_zz_model_route.py tests/test_streaming_gpu.py -m gpu -q -x > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
tail -n 30 $O/tests.log | cut -c1-300
if grep -q "rc=0" $O/tests.log; then
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --parity-utts 2 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
  tail -c 1800 $O/bench.json
fi
