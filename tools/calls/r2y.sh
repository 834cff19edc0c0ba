O=gpurun_out/r2y; mkdir -p $O; cd /root/repo
timeout 200 python -m pytest tests/test_feat_gpu.py tests/test_streaming_gpu.py -m gpu -q > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
tail -n 40 $O/tests.log | cut -c1-400
