O=gpurun_out/r2v; mkdir -p $O; cd /root/repo
timeout 280 python -m pytest tests/test_nnet_stream.py tests/test_ivector_gpu.py tests/test_feat_gpu.py tests/test_decoder_bestpath_gpu.py -m gpu -q > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
tail -n 50 $O/tests.log | cut -c1-400
