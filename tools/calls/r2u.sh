O=gpurun_out/r2u; mkdir -p $O; cd /root/repo
timeout 240 python -m pytest tests/test_nnet_stream.py tests/test_decoder_bestpath_gpu.py tests/test_ivector_gpu.py -m gpu -q > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
tail -n 60 $O/tests.log | cut -c1-300
