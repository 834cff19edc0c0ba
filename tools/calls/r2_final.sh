# final evidence of round 2 (one GPU): the driver's own bench commands, launch list, DRAM traffic of the decoder kernel,
# full ncu captures of the decoder kernel and of two GEMM launches, config-3 and sweep bench lines
O=gpurun_out/r2_final; mkdir -p $O; cd /root/repo
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-utts 0 > $O/ncu_launches.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dec_advance_v2 -c 4 --csv --log-file $O/decoder_dram.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-utts 0 > $O/ncu_dram.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dec_advance_v2 -s 1 -c 1 -o $O/dec_advance_v2_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-utts 0 > $O/ncu_dec_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nnet_gemm_ts -s 1 -c 6 -o $O/nnet_gemm_ts_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline --parity-utts 0 > $O/ncu_gemm_full.log 2>&1
timeout 1200 python bench.py --workload librispeech_tdnn_1d/hclg50M/batch512 --max-tpf 65536 --tok-per-frame 14000 --links-per-frame 28000 --steps 5 --warmup 3 --cpu-utts 2 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
timeout 900 python bench.py --workload decoder_sweep --steps 2 --warmup 1 > $O/bench_sweep.json 2> $O/bench_sweep.err
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests_gpu_all.log 2>&1; echo "rc=$?" >> $O/tests_gpu_all.log
B2K_BIG_TESTS=1 timeout 1500 python -m pytest tests/test_scale_gpu.py -m gpu -q -k 50m > $O/tests_50m.log 2>&1; echo "rc=$?" >> $O/tests_50m.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
tail -c 600 $O/bench_n1.json; echo; tail -c 400 $O/bench_reference_n1.json; echo; tail -n 3 $O/tests_gpu_all.log $O/tests_50m.log $O/smoke.log; ls -la $O
