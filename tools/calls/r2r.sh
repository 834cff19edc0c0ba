O=gpurun_out/r2r; mkdir -p $O; cd /root/repo
timeout 500 python -m pytest tests/test_decoder_gpu.py -m gpu -q -x > $O/tests_dec.log 2>&1; echo "rc=$?" >> $O/tests_dec.log
B2K_DEC_PARWALK=1 timeout 500 python -m pytest tests/test_decoder_gpu.py -m gpu -q -x > $O/tests_dec_par.log 2>&1; echo "rc=$?" >> $O/tests_dec_par.log
B2K_DEC_RS_CAPS=256,64,256 timeout 500 python -m pytest tests/test_decoder_gpu.py -m gpu -q -x > $O/tests_dec_smallcaps.log 2>&1; echo "rc=$?" >> $O/tests_dec_smallcaps.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 2 > $O/bench_default.json 2> $O/bench_default.err
B2K_DEC_PROF=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --parity-utts 0 > $O/bench_prof.json 2> $O/bench_prof.err
timeout 600 python -m pytest tests/test_scale_gpu.py -m gpu -q -k decoder > $O/tests_scale_dec.log 2>&1; echo "rc=$?" >> $O/tests_scale_dec.log
tail -n 3 $O/tests_dec.log $O/tests_dec_par.log $O/tests_dec_smallcaps.log $O/tests_scale_dec.log; for f in default prof; do python -c "
import json,sys
d=json.loads(open('$O/bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']), d['stage_ms'], d.get('parity_checked')); print(d.get('decoder_phase_share'), d.get('eps_replay_per_frame'), d.get('replay_routes'))"; tail -n 3 $O/bench_$f.err; done
