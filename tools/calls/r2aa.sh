O=gpurun_out/r2aa; mkdir -p $O; cd /root/repo
timeout 200 python -m pytest tests/test_nnet_stream.py "tests/test_scale_gpu.py::test_full_width_looped_stream_vs_compiled_reference" -m gpu -q > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
tail -n 30 $O/tests.log | cut -c1-300
