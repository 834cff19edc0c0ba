O=gpurun_out/r2w; mkdir -p $O; cd /root/repo
timeout 280 python -m pytest tests/test_streaming_gpu.py tests/test_pipeline_gpu.py tests/test_zz_model_route.py -m gpu -q > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
tail -n 45 $O/tests.log | cut -c1-400; tail -n 3 $O/smoke.log
