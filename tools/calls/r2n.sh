O=gpurun_out/r2n; mkdir -p $O; cd /root/repo
timeout 500 python -m pytest tests/test_decoder_gpu.py tests/test_ivector_gpu.py -m gpu -q -x > $O/tests_dec.log 2>&1; echo "rc=$?" >> $O/tests_dec.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 2 > $O/bench_default.json 2> $O/bench_default.err
B2K_DEC_PROF=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 0 > $O/bench_prof.json 2> $O/bench_prof.err
B2K_FIN_SMEM=0 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 0 > $O/bench_nofinsmem.json 2> $O/bench_nofinsmem.err
for it in 2 3; do B2K_DEC_IT=$it timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 1 > $O/bench_it$it.json 2> $O/bench_it$it.err; done
timeout 600 python -m pytest tests/test_scale_gpu.py -m gpu -q -k decoder > $O/tests_scale_dec.log 2>&1; echo "rc=$?" >> $O/tests_scale_dec.log
tail -n 3 $O/tests_dec.log $O/tests_scale_dec.log; for f in default prof nofinsmem it2 it3; do python -c "
import json,sys
d=json.loads(open('$O/bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']), d['stage_ms'], d.get('parity_checked')); print(d.get('decoder_phase_share'))"; done
