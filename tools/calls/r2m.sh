O=gpurun_out/r2m; mkdir -p $O; cd /root/repo
timeout 500 python -m pytest tests/test_decoder_gpu.py -m gpu -q -x > $O/tests_dec.log 2>&1; echo "rc=$?" >> $O/tests_dec.log
B2K_DEC_RS_CAPS=256,64,256 timeout 500 python -m pytest tests/test_decoder_gpu.py -m gpu -q -x > $O/tests_dec_smallcaps.log 2>&1; echo "rc=$?" >> $O/tests_dec_smallcaps.log
B2K_DEC_CID_SMEM=0 timeout 500 python -m pytest tests/test_decoder_gpu.py -m gpu -q -x > $O/tests_dec_nocid.log 2>&1; echo "rc=$?" >> $O/tests_dec_nocid.log
timeout 500 python -m pytest tests/test_nnet_gpu.py tests/test_ivector_gpu.py tests/test_pipeline_gpu.py tests/test_zz_model_route.py -m gpu -q -rf > $O/tests_nnet_iv_zz.log 2>&1; echo "rc=$?" >> $O/tests_nnet_iv_zz.log
for a in librispeech_1d:150 librispeech_cnn_tdnn_1a:90 mini_librispeech_1k:150; do timeout 200 python tools/nnet_precision_probe.py ${a%%:*} ${a##*:} $O/prec_${a%%:*}.npy 2>&1 | tail -n 1 >> $O/precision.log; done
B2K_NNET_FLUSH=0 timeout 200 python tools/nnet_precision_probe.py librispeech_1d 150 $O/prec_1d_noflush.npy 2>&1 | tail -n 1 >> $O/precision.log
timeout 600 python -m pytest tests/test_scale_gpu.py -m gpu -q > $O/tests_scale.log 2>&1; echo "rc=$?" >> $O/tests_scale.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 2 > $O/bench_default.json 2> $O/bench_default.err
B2K_DEC_PROF=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 0 > $O/bench_prof.json 2> $O/bench_prof.err
B2K_DEC_CID_SMEM=0 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 0 > $O/bench_nocid.json 2> $O/bench_nocid.err
B2K_NNET_FLUSH=0 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 0 > $O/bench_noflush.json 2> $O/bench_noflush.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-utts 0 > $O/ncu_bench.log 2>&1
timeout 1200 python bench.py --workload librispeech_tdnn_1d/hclg50M/batch512 --max-tpf 65536 --tok-per-frame 14000 --links-per-frame 28000 --steps 3 --warmup 2 --no-cpu-baseline > $O/bench_cfg3.json 2> $O/bench_cfg3.err
for f in tests_dec tests_dec_smallcaps tests_dec_nocid tests_nnet_iv_zz tests_scale; do tail -n 3 $O/$f.log; done; cat $O/precision.log; grep -h "finals here" $O/tests_nnet_iv_zz.log | head -3
