O=gpurun_out/r2z; mkdir -p $O; cd /root/repo
timeout 300 python -m pytest tests -m gpu -q > $O/tests_full.log 2>&1; echo "rc=$?" >> $O/tests_full.log
tail -n 25 $O/tests_full.log | cut -c1-300
timeout 150 python bench.py --workload streaming --steps 2 --warmup 1 > $O/bench_streaming.json 2> $O/bench_streaming.err; echo "bench rc=$?"
tail -c 1500 $O/bench_streaming.json; tail -n 5 $O/bench_streaming.err | cut -c1-300
