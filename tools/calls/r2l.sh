O=gpurun_out/r2l; mkdir -p $O; cd /root/repo; P=tools/_bin/tc5b
timeout 400 python -m pytest tests/test_decoder_gpu.py -m gpu -q -x > $O/tests_dec.log 2>&1; echo "rc=$?" >> $O/tests_dec.log
for acc in 0 1 2; do for dist in 0 1; do timeout 120 $P 2 8192 96 4608 $acc $dist >> $O/probe_acc.log 2>&1; done; done
for acc in 0 1; do timeout 120 $P 1 8192 768 288 $acc 1 >> $O/probe_acc.log 2>&1; timeout 120 $P 1 8192 1536 480 $acc 1 >> $O/probe_acc.log 2>&1; timeout 120 $P 2 8192 96 1536 $acc 1 >> $O/probe_acc.log 2>&1; done
timeout 300 python -m pytest tests/test_ivector_gpu.py tests/test_zz_model_route.py tests/test_pipeline_gpu.py -m gpu -q -rxXf > $O/tests_iv_zz.log 2>&1; echo "rc=$?" >> $O/tests_iv_zz.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 1 > $O/bench_default.json 2> $O/bench_default.err
B2K_DEC_PROF=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 0 > $O/bench_prof.json 2> $O/bench_prof.err
for caps in 2048,2048,4096 4096,1024,2048 2048,4096,2048; do B2K_DEC_THREADS=256 B2K_DEC_CTAS=3 B2K_DEC_RS_CAPS=$caps B2K_DEC_PROF=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 1 > $O/bench_t256c3_$caps.json 2> $O/bench_t256c3_$caps.err; done
B2K_DEC_THREADS=256 B2K_DEC_CTAS=2 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 1 > $O/bench_t256c2.json 2> $O/bench_t256c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-utts 0 > $O/ncu_bench.log 2>&1
timeout 1200 python bench.py --workload librispeech_tdnn_1d/hclg50M/batch512 --max-tpf 65536 --tok-per-frame 14000 --links-per-frame 28000 --steps 3 --warmup 2 --cpu-utts 2 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
tail -n 3 $O/tests_dec.log; cat $O/probe_acc.log; tail -n 5 $O/tests_iv_zz.log; tail -c 400 $O/bench_cfg3.err
